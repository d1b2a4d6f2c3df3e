// The reference's double-precision arithmetic on the device, operation for operation and never contracted into FMAs:
// the 8-point Arai pass of include/Dct.hpp:52-131 / 134-213 in a strided and in an all-register form.  Shared by the exact
// refinement / planes kernels (forward.cu) and the stage kernels (stages.cu).
#pragma once
#include "common.cuh"

namespace jpgenc {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// one 8-point pass, Dct.hpp:52-131 / 134-213; o has stride `os`
__device__ __forceinline__ void aan8_exact(const double* x, int xs, double* o, int os, const ExactConsts& e) {
    const double x0 = x[0], x1 = x[xs], x2 = x[2 * xs], x3 = x[3 * xs], x4 = x[4 * xs], x5 = x[5 * xs], x6 = x[6 * xs],
                 x7 = x[7 * xs];
    const double z0 = dadd(x0, x7), z1 = dadd(x1, x6), z2 = dadd(x2, x5), z3 = dadd(x3, x4);
    const double z4 = dadd(-x4, x3), z5 = dadd(-x5, x2), z6 = dadd(-x6, x1), z7 = dadd(-x7, x0);
    const double r0 = dadd(z0, z3), r1 = dadd(z1, z2), r2 = dsub(z1, z2), r3 = dsub(z0, z3);
    const double r4 = dsub(-z4, z5), r5 = dadd(z5, z6), r6 = dadd(z6, z7), r7 = z7;
    const double t0 = dadd(r0, r1), t1 = dsub(r0, r1);
    double t2 = dadd(r2, r3), t4 = r4, t5 = r5, t6 = r6;
    const double t3 = r3, t7 = r7;
    const double tmp = dmul(dadd(t4, t6), e.a5);
    t2 = dmul(t2, e.a1); t4 = dmul(t4, e.a2); t5 = dmul(t5, e.a3); t6 = dmul(t6, e.a4);
    const double u4 = dsub(-t4, tmp), u6 = dsub(t6, tmp);
    const double v2 = dadd(t2, t3), v3 = dsub(t3, t2), v5 = dadd(t5, t7), v7 = dsub(t7, t5);
    const double w4 = dadd(u4, v7), w5 = dadd(v5, u6), w6 = dadd(-u6, v5), w7 = dsub(v7, u4);
    o[0 * os] = dmul(t0, e.s[0]); o[4 * os] = dmul(t1, e.s[4]); o[2 * os] = dmul(v2, e.s[2]); o[6 * os] = dmul(v3, e.s[6]);
    o[5 * os] = dmul(w4, e.s[5]); o[1 * os] = dmul(w5, e.s[1]); o[7 * os] = dmul(w6, e.s[7]); o[3 * os] = dmul(w7, e.s[3]);
}

// one 8-point pass on registers, Dct.hpp:52-131 / 134-213
__device__ __forceinline__ void aan8_exact_regs(const double (&x)[8], double (&o)[8], const ExactConsts& e) {
    const double z0 = dadd(x[0], x[7]), z1 = dadd(x[1], x[6]), z2 = dadd(x[2], x[5]), z3 = dadd(x[3], x[4]);
    const double z4 = dadd(-x[4], x[3]), z5 = dadd(-x[5], x[2]), z6 = dadd(-x[6], x[1]), z7 = dadd(-x[7], x[0]);
    const double r0 = dadd(z0, z3), r1 = dadd(z1, z2), r2 = dsub(z1, z2), r3 = dsub(z0, z3);
    const double r4 = dsub(-z4, z5), r5 = dadd(z5, z6), r6 = dadd(z6, z7), r7 = z7;
    const double t0 = dadd(r0, r1), t1 = dsub(r0, r1);
    double t2 = dadd(r2, r3), t4 = r4, t5 = r5, t6 = r6;
    const double t3 = r3, t7 = r7;
    const double tmp = dmul(dadd(t4, t6), e.a5);
    t2 = dmul(t2, e.a1); t4 = dmul(t4, e.a2); t5 = dmul(t5, e.a3); t6 = dmul(t6, e.a4);
    const double u4 = dsub(-t4, tmp), u6 = dsub(t6, tmp);
    const double v2 = dadd(t2, t3), v3 = dsub(t3, t2), v5 = dadd(t5, t7), v7 = dsub(t7, t5);
    const double w4 = dadd(u4, v7), w5 = dadd(v5, u6), w6 = dadd(-u6, v5), w7 = dsub(v7, u4);
    o[0] = dmul(t0, e.s[0]); o[4] = dmul(t1, e.s[4]); o[2] = dmul(v2, e.s[2]); o[6] = dmul(v3, e.s[6]);
    o[5] = dmul(w4, e.s[5]); o[1] = dmul(w5, e.s[1]); o[7] = dmul(w6, e.s[7]); o[3] = dmul(w7, e.s[3]);
}

}  // namespace jpgenc
