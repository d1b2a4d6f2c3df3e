// Per-block coding helpers with the reference's public interface (include/Coding.hpp:17-283): zigzag, quantisation,
// run-length coding of AC coefficients, size categories.  Host-side utilities for callers of the stage API and for
// tests; inside Image::writeJPEG this work runs on the GPU (csrc/forward.cu, csrc/stats.cu, csrc/entropy.cu).
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <utility>
#include <vector>

#include "BitstreamGeneric.hpp"
#include "matrix_types.hpp"

typedef double PixelDataType;
using mat = matrix<PixelDataType>;
typedef unsigned int uint;
typedef uint8_t Byte;

namespace jpgenc_detail {
// natural index (row*8+col) of the i-th coefficient of the zigzag scan
inline const uint8_t* zigzag_order() {
    static const uint8_t order[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                      41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                      30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return order;
}
}  // namespace jpgenc_detail

template <typename T>
matrix<T> from_vector(const std::vector<T>& v) {
    assert(v.size() == 64);
    matrix<T> m(8, 8);
    for (std::size_t i = 0; i < 64; ++i) m(i / 8, i % 8) = v[i];
    return m;
}

// 8x8 matrix -> its 64 values in zigzag order
template <typename T>
std::vector<T> zigzag(matrix<T> m) {
    assert(m.size1() == 8 && m.size2() == 8);
    std::vector<T> out(64);
    const uint8_t* order = jpgenc_detail::zigzag_order();
    for (int i = 0; i < 64; ++i) out[i] = m(order[i] / 8, order[i] % 8);
    return out;
}
// natural index of the i-th zigzag coefficient
inline int zigzag(int i) {
    assert(i >= 0 && i < 64);
    return jpgenc_detail::zigzag_order()[i];
}

// int(round(m / table)) element-wise, round half away from zero (reference Coding.hpp:84-97)
inline matrix<int> quantize(const mat& m, const mat& table) {
    assert(m.size1() == 8 && m.size2() == 8 && table.size1() == 8 && table.size2() == 8);
    matrix<int> q(8, 8);
    for (std::size_t r = 0; r < 8; ++r)
        for (std::size_t c = 0; c < 8; ++c) q(r, c) = static_cast<int>(std::round(m(r, c) / table(r, c)));
    return q;
}

struct RLE_PAIR {
    unsigned short num_zeros_before : 4;
    int value;
    RLE_PAIR() = default;
    RLE_PAIR(short zeros, int v) : num_zeros_before(zeros), value(v) { assert(zeros < 16); }
};
inline bool operator==(const RLE_PAIR& a, const RLE_PAIR& b) { return a.num_zeros_before == b.num_zeros_before && a.value == b.value; }

// zigzag-ordered values (first = DC) -> (run, value) pairs with (15,0) for 16 zeros and a final (0,0) when zeros trail
inline std::vector<RLE_PAIR> RLE_AC(const std::vector<int>& data) {
    assert(data.size() > 1);
    std::vector<RLE_PAIR> out(1, RLE_PAIR(0, data[0]));
    unsigned zeros = 0;
    for (std::size_t i = 1; i < data.size(); ++i) {
        if (data[i] == 0) { ++zeros; continue; }
        for (; zeros > 15; zeros -= 16) out.push_back(RLE_PAIR(15, 0));
        out.push_back(RLE_PAIR(static_cast<short>(zeros), data[i]));
        zeros = 0;
    }
    if (zeros) out.push_back(RLE_PAIR(0, 0));
    return out;
}
// natural-order 8x8 block: zigzag first, then as above
inline std::vector<RLE_PAIR> RLE_AC(const matrix<int>& block) {
    assert(block.size1() == 8 && block.size2() == 8);
    std::vector<int> zz(64);
    const uint8_t* order = jpgenc_detail::zigzag_order();
    for (int i = 0; i < 64; ++i) zz[i] = block(order[i] / 8, order[i] % 8);
    return RLE_AC(zz);
}

struct Category_Code {
    uint8_t symbol;
    Bitstream code;
    Category_Code(uint8_t s, Bitstream b) : symbol(s), code(b) {}
    ~Category_Code() {}
};
inline bool operator==(const Category_Code& a, const Category_Code& b) { return a.symbol == b.symbol && a.code == b.code; }

// size category (bits needed for |value|) and the magnitude bits: value itself when positive, 2^cat-1-|value| otherwise
inline void getCategoryAndCode(int value, short& category, Bitstream& code) {
    if (value == 0) { category = 0; code = Bitstream(); return; }
    const long mag = std::labs(static_cast<long>(value));
    short cat = 1;
    while ((1L << cat) <= mag) ++cat;
    assert(cat < 16);
    category = cat;
    code = Bitstream(static_cast<uint32_t>(value < 0 ? (1L << cat) - 1 - mag : value), cat);
}
inline std::pair<short, Bitstream> getCategoryAndCode(int value) {
    std::pair<short, Bitstream> r;
    getCategoryAndCode(value, r.first, r.second);
    return r;
}

// (run, value) pairs -> Huffman symbol (run<<4 | category) + magnitude bits
inline std::vector<Category_Code> encode_category(const std::vector<RLE_PAIR>& data) {
    std::vector<Category_Code> out;
    out.reserve(data.size());
    for (const RLE_PAIR& p : data) {
        short cat = 0;
        Bitstream bits;
        getCategoryAndCode(p.value, cat, bits);
        out.emplace_back(static_cast<uint8_t>((p.num_zeros_before << 4) | cat), bits);
    }
    return out;
}
