// Stage methods of `Image` on the device for the modes Image::writeJPEG does not use:
//   Image::applySubsampling(mode) for every SubsamplingMode   (reference src/Image.cpp:198-319)
//   Image::applyDCT(mode) for DCTMode Simple / Matrix / Arai   (reference src/Image.cpp:540-595, include/Dct.hpp:47-276)
// on whole planes of doubles, in the reference's arithmetic (FP64, its operation order, nothing contracted), so the planes
// that come back are the reference's, bit for bit.  These are not on the encode hot path -- writeJPEG is S420_m + Arai,
// fused in K1 -- they exist so that a caller driving the stages one by one (the reference's tests do) finds every mode.
#include <cmath>
#include <cstring>

#include "exact.cuh"
#include "internal.hpp"

namespace jpgenc {

// mode: the enum order of include/Image.hpp -- 0 S444, 1 S422, 2 S411, 3 S420, 4 S420_m, 5 S420_lm
struct SubsampleGeometry { uint32_t hdiv, vdiv, taps; double w1; bool second_line; double div; };
static SubsampleGeometry subsample_geometry(int mode) {
    switch (mode) {
        case 1: return {2, 1, 2, 0.0, false, 1.0};             // x- x-            Mask{{1,0}, false}
        case 2: return {4, 1, 4, 0.0, false, 1.0};             // x- --            Mask{{1,0,0,0}, false}
        case 3: return {2, 2, 2, 0.0, false, 1.0};             // every 2nd row    Mask{{1,0}, true}
        case 4: return {2, 2, 2, 1.0, true, 4.0};              // mean of 2x2      Mask{{1,1}, false}, averaging
        case 5: return {2, 2, 2, 0.0, true, 2.0};              // vertical mean    Mask{{1,0}, false}, averaging
        default: return {1, 1, 1, 0.0, false, 1.0};
    }
}

// one thread per output sample.  The reference adds EVERY tap, the zero-weighted ones included (pix = 0; pix += row[m] *
// chan(y, x + m)), and for the averaging modes adds the second scanline's sum to the first and divides in place
// (src/Image.cpp:209-226): the same operations here.
__global__ void subsample_plane_kernel(const double* __restrict__ in, uint32_t w, double* __restrict__ out, uint32_t ow, uint32_t oh,
                                       SubsampleGeometry g) {
    const uint32_t ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= ow || oy >= oh) return;
    auto line = [&](uint32_t y) {
        const double* p = in + static_cast<size_t>(y) * w + static_cast<size_t>(ox) * g.taps;
        double pix = dadd(0.0, dmul(1.0, p[0]));
        if (g.taps >= 2) pix = dadd(pix, dmul(g.w1, p[1]));
        if (g.taps == 4) { pix = dadd(pix, dmul(0.0, p[2])); pix = dadd(pix, dmul(0.0, p[3])); }
        return pix;
    };
    double v = line(oy * g.vdiv);
    if (g.second_line) v = __ddiv_rn(dadd(v, line(oy * g.vdiv + 1)), g.div);
    out[static_cast<size_t>(oy) * ow + ox] = v;
}

struct DctBasis { double a[64]; };                              // A of Dct.hpp:217-236, row-major

// one 8x8 block per 64 threads (thread = one output sample for Simple / Matrix; threads 0..7 = the 1-D passes of Arai)
__global__ void __launch_bounds__(256) dct_plane_kernel(const double* __restrict__ in, uint32_t w, uint32_t blocks_x, uint32_t nblocks,
                                                        double* __restrict__ out, int mode, const __grid_constant__ DctBasis basis,
                                                        const __grid_constant__ ExactConsts e) {
    __shared__ double s_x[4][64], s_t[4][64];
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63, i = t >> 3, j = t & 7;
    const uint32_t blk = blockIdx.x * 4 + sub;
    const bool live = blk < nblocks;
    const uint32_t by = live ? blk / blocks_x : 0, bx = live ? blk - by * blocks_x : 0;
    const size_t at = (static_cast<size_t>(by) * 8 + i) * w + static_cast<size_t>(bx) * 8 + j;
    double* X = s_x[sub];
    double* T = s_t[sub];
    X[t] = live ? in[at] : 0.0;
    __syncthreads();
    double y = 0.0;
    if (mode == 0) {
        // dctDirect (Dct.hpp:238-262): Y(j2, i2) = sum over x, then y of X(y, x) * A(i2, x) * A(j2, y); this thread owns Y(i, j),
        // i.e. the reference's loop variables are i2 = j, j2 = i
        double sum = 0.0;
        for (int x = 0; x < 8; ++x)
            for (int yy = 0; yy < 8; ++yy) sum = dadd(sum, dmul(dmul(X[yy * 8 + x], basis.a[j * 8 + x]), basis.a[i * 8 + yy]));
        y = sum;
    } else if (mode == 1) {
        // dctMat (Dct.hpp:264-276): first = X * A^T, Y = A * first, both as plain ascending-k sums (ublas prod)
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s = dadd(s, dmul(X[i * 8 + k], basis.a[j * 8 + k]));
        T[t] = s;
        __syncthreads();
        s = 0.0;
        for (int k = 0; k < 8; ++k) s = dadd(s, dmul(basis.a[i * 8 + k], T[k * 8 + j]));
        y = s;
    } else {
        // dctArai (Dct.hpp:47-215): pass 1 turns column c of the block into row c of the temporary, pass 2 column c of the
        // temporary into row c of the result
        if (t < 8) aan8_exact(X + t, 8, T + t * 8, 1, e);
        __syncthreads();
        if (t < 8) aan8_exact(T + t, 8, X + t * 8, 1, e);
        __syncthreads();
        y = X[t];
    }
    if (live) out[at] = y;
}

}  // namespace jpgenc

using namespace jpgenc;
using namespace jpgenc::detail;

extern "C" {

int jpgenc_stage_subsample_dims(int mode, uint32_t width, uint32_t height, uint32_t* out_width, uint32_t* out_height) {
    if (mode < 0 || mode > 5 || !out_width || !out_height) return JPGENC_ERR_ARG;
    const SubsampleGeometry g = subsample_geometry(mode);
    *out_width = width / g.hdiv;
    *out_height = height / g.vdiv;
    return JPGENC_OK;
}

int jpgenc_stage_subsample(jpgenc_ctx* c, const double* plane, uint32_t width, uint32_t height, int mode, double* out) try {
    if (!c || !plane || !out || mode < 0 || mode > 5 || width == 0 || height == 0) return JPGENC_ERR_ARG;
    if (mode == 0) {                                             // S444: applySubsampling returns at once (src/Image.cpp:263-268)
        std::memcpy(out, plane, static_cast<size_t>(width) * height * sizeof(double));
        return JPGENC_OK;
    }
    const SubsampleGeometry g = subsample_geometry(mode);
    // the reference walks the plane two scanlines / `taps` columns at a time and would run past its end otherwise
    if (width % g.taps || (g.vdiv == 2 && height % 2)) return fail(c, JPGENC_ERR_ARG, "plane size is not a multiple of the subsampling pattern");
    JPGENC_CUDA(c, cudaSetDevice(c->device));
    const uint32_t ow = width / g.hdiv, oh = height / g.vdiv;
    const size_t n_in = static_cast<size_t>(width) * height, n_out = static_cast<size_t>(ow) * oh;
    double* d = nullptr;
    JPGENC_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&d), (n_in + n_out) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, plane, n_in * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        subsample_plane_kernel<<<dim3((ow + 127) / 128, oh), 128, 0, c->stream>>>(d, width, d + n_in, ow, oh, g);
        e = cudaGetLastError();
        c->launches += 1;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + n_in, n_out * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) { c->error = std::string("jpgenc_stage_subsample: ") + cudaGetErrorString(e); return JPGENC_ERR_CUDA; }
    return JPGENC_OK;
} JPGENC_CATCH(c)

int jpgenc_stage_dct(jpgenc_ctx* c, const double* plane, uint32_t width, uint32_t height, int mode, double* out) try {
    if (!c || !plane || !out || mode < 0 || mode > 2) return JPGENC_ERR_ARG;
    if (width == 0 || height == 0 || width % 8 || height % 8) return fail(c, JPGENC_ERR_ARG, "plane sides must be multiples of 8");
    JPGENC_CUDA(c, cudaSetDevice(c->device));
    DctBasis basis;
    {   // the reference's expressions (Dct.hpp:217-236); pi and root_two are Boost's double constants
        const double pi = 3.141592653589793238462643383279502884, root_two = 1.414213562373095048801688724209698078;
        const double scale = std::sqrt(2. / 8);
        for (unsigned k = 0; k < 8; ++k)
            for (unsigned n = 0; n < 8; ++n) {
                const double term = (2. * n + 1.) * ((k * pi) / 16.);
                basis.a[k * 8 + n] = (k == 0 ? 1. / root_two : 1.) * scale * std::cos(term);
            }
    }
    ExactConsts ex{};
    ex.a1 = c->dct_a[0]; ex.a2 = c->dct_a[1]; ex.a3 = c->dct_a[2]; ex.a4 = c->dct_a[3]; ex.a5 = c->dct_a[4];
    for (int i = 0; i < 8; ++i) ex.s[i] = c->dct_s[i];
    const size_t n = static_cast<size_t>(width) * height;
    const uint32_t blocks_x = width / 8, nblocks = blocks_x * (height / 8);
    double* d = nullptr;
    JPGENC_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&d), 2 * n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, plane, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        dct_plane_kernel<<<(nblocks + 3) / 4, 256, 0, c->stream>>>(d, width, blocks_x, nblocks, d + n, mode, basis, ex);
        e = cudaGetLastError();
        c->launches += 1;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) { c->error = std::string("jpgenc_stage_dct: ") + cudaGetErrorString(e); return JPGENC_ERR_CUDA; }
    return JPGENC_OK;
} JPGENC_CATCH(c)

}  // extern "C"
