// 8x8 DCT entry points with the reference's signatures (include/Dct.hpp:47,238,264,278).
//
// On the encode path (Image::writeJPEG) the Arai transform runs on the GPU, fused with colour conversion and
// quantisation (csrc/forward.cu).  These per-block host entry points remain for callers of the stage API and for the
// reference's own tests: dctArai is the exact double-precision AAN network (the arithmetic the GPU refinement kernel
// reproduces operation for operation), dctDirect / dctMat are the two non-default variants (checked to 1e-5 only).
#pragma once
#include <cassert>
#include <cmath>

#include "Image.hpp"

namespace jpgenc_detail {
struct AanConstants {
    double a1, a2, a3, a4, a5, s[8];
    AanConstants() {
        // the reference's expressions (Dct.hpp:21-43)
        const double pi = 3.141592653589793238462643383279502884, root_two = 1.414213562373095048801688724209698078;
        double c[8];
        for (int k = 1; k < 8; ++k) c[k] = std::cos(k * pi / 16);
        a1 = c[4]; a2 = c[2] - c[6]; a3 = c[4]; a4 = c[6] + c[2]; a5 = c[6];
        s[0] = 1 / (2 * root_two);
        for (int k = 1; k < 8; ++k) s[k] = 1 / (4 * c[k]);
    }
};
inline const AanConstants& aan() { static const AanConstants k; return k; }

// 8-point Arai/Agui/Nakajima network; in[] -> out[] (frequency order), every operation rounded separately in the
// reference's order so results are bit-identical to Dct.hpp:52-131
inline void aan_1d(const double in[8], double out[8]) {
    const AanConstants& k = aan();
    const double e0 = in[0] + in[7], e1 = in[1] + in[6], e2 = in[2] + in[5], e3 = in[3] + in[4];
    const double o0 = -in[4] + in[3], o1 = -in[5] + in[2], o2 = -in[6] + in[1], o3 = -in[7] + in[0];
    const double p0 = e0 + e3, p1 = e1 + e2, p2 = e1 - e2, p3 = e0 - e3;
    const double q0 = -o0 - o1, q1 = o1 + o2, q2 = o2 + o3;
    const double rot = (q0 + q2) * k.a5;
    const double m2 = (p2 + p3) * k.a1, m4 = q0 * k.a2, m5 = q1 * k.a3, m6 = q2 * k.a4;
    const double n4 = -m4 - rot, n6 = m6 - rot;
    const double h5 = m5 + o3, h7 = o3 - m5;
    out[0] = (p0 + p1) * k.s[0];
    out[4] = (p0 - p1) * k.s[4];
    out[2] = (m2 + p3) * k.s[2];
    out[6] = (p3 - m2) * k.s[6];
    out[5] = (n4 + h7) * k.s[5];
    out[1] = (h5 + n6) * k.s[1];
    out[7] = (-n6 + h5) * k.s[7];
    out[3] = (h7 - n4) * k.s[3];
}
inline void dct_basis(double a[8][8]) {
    const double pi = 3.141592653589793238462643383279502884, root_two = 1.414213562373095048801688724209698078;
    for (unsigned k = 0; k < 8; ++k)
        for (unsigned n = 0; n < 8; ++n)
            a[k][n] = (k == 0 ? 1. / root_two : 1.) * std::sqrt(2. / 8) * std::cos((2. * n + 1.) * ((k * pi) / 16.));
}
}  // namespace jpgenc_detail

inline void dctArai(const matrix_range<matrix<PixelDataType>>& x, matrix_range<matrix<PixelDataType>>& y) {
    assert(x.size1() == 8 && x.size2() == 8);
    double col[8], tmp[8][8];
    for (int j = 0; j < 8; ++j) {                       // columns of x -> rows of tmp
        for (int i = 0; i < 8; ++i) col[i] = x(i, j);
        jpgenc_detail::aan_1d(col, tmp[j]);
    }
    for (int j = 0; j < 8; ++j) {                       // columns of tmp -> rows of y
        double out[8];
        for (int i = 0; i < 8; ++i) col[i] = tmp[i][j];
        jpgenc_detail::aan_1d(col, out);
        for (int i = 0; i < 8; ++i) y(j, i) = out[i];
    }
}

inline void dctDirect(const matrix_range<matrix<PixelDataType>>& X, matrix_range<matrix<PixelDataType>>& Y) {
    assert(X.size1() == 8 && X.size2() == 8);
    double a[8][8];
    jpgenc_detail::dct_basis(a);
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            double acc = 0.0;
            for (int x = 0; x < 8; ++x)
                for (int y = 0; y < 8; ++y) acc += X(y, x) * a[u][x] * a[v][y];
            Y(v, u) = acc;
        }
}

inline void dctMat(const matrix_range<matrix<PixelDataType>>& X, matrix_range<matrix<PixelDataType>>& Y) {
    assert(X.size1() == 8 && X.size2() == 8 && Y.size1() == 8 && Y.size2() == 8);
    double a[8][8], xa[8][8];
    jpgenc_detail::dct_basis(a);
    for (int i = 0; i < 8; ++i)                         // X * A^T
        for (int j = 0; j < 8; ++j) {
            double acc = 0;
            for (int k = 0; k < 8; ++k) acc += X(i, k) * a[j][k];
            xa[i][j] = acc;
        }
    for (int i = 0; i < 8; ++i)                         // A * (X * A^T)
        for (int j = 0; j < 8; ++j) {
            double acc = 0;
            for (int k = 0; k < 8; ++k) acc += a[i][k] * xa[k][j];
            Y(i, j) = acc;
        }
}

inline matrix<PixelDataType> inverseDctMat(matrix<PixelDataType> X) {
    assert(X.size1() == 8 && X.size2() == 8);
    double a[8][8], ax[8][8];
    jpgenc_detail::dct_basis(a);
    for (int i = 0; i < 8; ++i)                         // A^T * X
        for (int j = 0; j < 8; ++j) {
            double acc = 0;
            for (int k = 0; k < 8; ++k) acc += a[k][i] * X(k, j);
            ax[i][j] = acc;
        }
    matrix<PixelDataType> out(8, 8);
    for (int i = 0; i < 8; ++i)                         // (A^T * X) * A
        for (int j = 0; j < 8; ++j) {
            double acc = 0;
            for (int k = 0; k < 8; ++k) acc += ax[i][k] * a[k][j];
            out(i, j) = acc;
        }
    return out;
}
