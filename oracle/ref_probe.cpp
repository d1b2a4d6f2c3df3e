// C-ABI probe over the UNMODIFIED reference encoder (Nuos/jpgEnc), used only to pin the oracle.
// TEST INFRASTRUCTURE: built by oracle/build_ref.sh into oracle/_ref/libjpgenc_ref.so from the
// reference's own sources where they lie (/root/reference); nothing here is shipped or timed
// as the product.  Compiled with -fno-access-control so the stage members of `Image`
// (Image.hpp:104-114) can be read back after each public stage method.
#include "Image.hpp"
#include "Dct.hpp"
#include "JpegSegments.hpp"

#include <chrono>
#include <cstring>
#include <sstream>

using boost::numeric::ublas::range;
using boost::numeric::ublas::matrix_range;

namespace {
matrix<int> annexk_luma() {
    return from_vector<int>({16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
                             14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                             18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                             49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99});
}
matrix<int> annexk_chroma() {
    return from_vector<int>({17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                             24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                             99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                             99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99});
}
template <class T, class M>
void copy_out(const M& m, T* dst) {
    if (!dst) return;
    for (std::size_t i = 0; i < m.size1(); ++i)
        for (std::size_t j = 0; j < m.size2(); ++j)
            dst[i * m.size2() + j] = static_cast<T>(m(i, j));
}
}

extern "C" {

// whole file, exactly what main.cpp does (main.cpp:16,29). Returns 0, or -1 on exception.
int ref_encode_file(const char* ppm_path, const char* jpg_path, double* load_ms, double* encode_ms) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        auto img = loadPPM(ppm_path);
        auto t1 = std::chrono::steady_clock::now();
        img.writeJPEG(jpg_path);
        auto t2 = std::chrono::steady_clock::now();
        if (load_ms) *load_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (encode_ms) *encode_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// padded geometry after loadPPM (Image.cpp:479-532)
int ref_padded_dims(const char* ppm_path, unsigned* w16, unsigned* h16, unsigned* real_w, unsigned* real_h) {
    try {
        auto img = loadPPM(ppm_path);
        *w16 = img.width; *h16 = img.height; *real_w = img.real_width; *real_h = img.real_height;
        return 0;
    } catch (const std::exception&) { return -1; }
}

// Stage dump: every pointer may be null.  Planes are row-major.
//   rgb    : 3 planes of h16*w16 doubles after loadPPM
//   ycc_y  : h16*w16, ycc_cb/ycc_cr : (h16/2)*(w16/2) doubles after colour conversion + S420_m
//   dct_*  : same shapes, after applyDCT(Arai)
//   q_*    : same shapes, int32, after applyQuantization and BEFORE DC differencing
int ref_stage_dump(const char* ppm_path, double* rgb, double* ycc_y, double* ycc_cb, double* ycc_cr,
                   double* dct_y, double* dct_cb, double* dct_cr, int* q_y, int* q_cb, int* q_cr) {
    try {
        auto img = loadPPM(ppm_path);
        const std::size_t n = std::size_t(img.width) * img.height;
        if (rgb) { copy_out(img.R, rgb); copy_out(img.G, rgb + n); copy_out(img.B, rgb + 2 * n); }
        img = img.convertToColorSpace(Image::YCbCr);
        img.applySubsampling(Image::S420_m);
        copy_out(img.Y, ycc_y); copy_out(img.Cb, ycc_cb); copy_out(img.Cr, ycc_cr);
        img.applyDCT(Image::Arai);
        copy_out(img.DctY, dct_y); copy_out(img.DctCb, dct_cb); copy_out(img.DctCr, dct_cr);
        img.applyQuantization(annexk_luma(), annexk_chroma());
        copy_out(img.QY, q_y); copy_out(img.QCb, q_cb); copy_out(img.QCr, q_cr);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// Image::writeJPEG on an Image assembled in memory (Image.hpp:36, 104-114): three planes of h16*w16 doubles, taken as R,G,B
// (ycbcr == 0) or as an image that already is YCbCr (ycbcr != 0: convertToColorSpace returns it unchanged, Image.cpp:112-115).
int ref_encode_planes(const double* p0, const double* p1, const double* p2, unsigned w16, unsigned h16, unsigned real_w,
                      unsigned real_h, int ycbcr, const char* jpg_path) {
    try {
        Image img(w16, h16, ycbcr ? Image::YCbCr : Image::RGB);
        img.real_width = real_w;
        img.real_height = real_h;
        const std::size_t n = std::size_t(w16) * h16;
        for (std::size_t i = 0; i < n; ++i) { img.R.data()[i] = p0[i]; img.G.data()[i] = p1[i]; img.B.data()[i] = p2[i]; }
        img.writeJPEG(jpg_path);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// Image::applySubsampling(mode) on an Image whose Cb plane is `in` (h x w doubles): the subsampled Cb plane comes back in `out`
// (sized by the caller: Image.cpp:237-319 gives the divisors); mode = the enum's order S444, S422, S411, S420, S420_m, S420_lm
int ref_subsample_plane(const double* in, unsigned w, unsigned h, int mode, double* out, unsigned* ow, unsigned* oh) {
    try {
        Image img(w, h, Image::YCbCr);
        for (std::size_t i = 0; i < std::size_t(w) * h; ++i) { img.Y.data()[i] = 0; img.Cb.data()[i] = in[i]; img.Cr.data()[i] = in[i]; }
        img.applySubsampling(static_cast<Image::SubsamplingMode>(mode));
        *ow = static_cast<unsigned>(img.Cb.size2()); *oh = static_cast<unsigned>(img.Cb.size1());
        copy_out(img.Cb, out);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// Image::applyDCT(mode) on an Image whose Y plane is `in` (h x w doubles, multiples of 8... the chroma planes are dummies)
int ref_dct_plane(const double* in, unsigned w, unsigned h, int mode, double* out) {
    try {
        Image img(w, h, Image::YCbCr);
        for (std::size_t i = 0; i < std::size_t(w) * h; ++i) { img.Y.data()[i] = in[i]; img.Cb.data()[i] = 0; img.Cr.data()[i] = 0; }
        img.applyDCT(static_cast<Image::DCTMode>(mode));
        copy_out(img.DctY, out);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// one 8x8 block through each DCT variant (Dct.hpp:47,238,264); mode 0=Simple 1=Matrix 2=Arai
void ref_dct_block(const double* in, double* out, int mode) {
    matrix<double> x(8, 8), y(8, 8);
    for (int i = 0; i < 64; ++i) x.data()[i] = in[i];
    const matrix_range<matrix<double>> xs(x, range(0, 8), range(0, 8));
    matrix_range<matrix<double>> ys(y, range(0, 8), range(0, 8));
    if (mode == 0) dctDirect(xs, ys); else if (mode == 1) dctMat(xs, ys); else dctArai(xs, ys);
    for (int i = 0; i < 64; ++i) out[i] = y.data()[i];
}

// quantize (Coding.hpp:84-97)
void ref_quantize_block(const double* in, const double* table, int* out) {
    matrix<double> m(8, 8), t(8, 8);
    for (int i = 0; i < 64; ++i) { m.data()[i] = in[i]; t.data()[i] = table[i]; }
    auto q = quantize(m, t);
    for (int i = 0; i < 64; ++i) out[i] = q.data()[i];
}

// RLE_AC(matrix) + encode_category (Coding.hpp:148-183,265-283): natural-order 8x8 ints ->
// symbols[], magnitude bit values[] and bit counts[]; returns the number of entries (<= 64).
int ref_block_symbols(const int* natural64, unsigned char* symbols, unsigned* mag_bits, unsigned char* mag_len) {
    matrix<int> m(8, 8);
    for (int i = 0; i < 64; ++i) m.data()[i] = natural64[i];
    auto cc = encode_category(RLE_AC(m));
    int k = 0;
    for (auto& e : cc) {
        symbols[k] = e.symbol;
        mag_len[k] = static_cast<unsigned char>(e.code.size());
        mag_bits[k] = e.code.size() ? (e.code.extract(static_cast<uint8_t>(e.code.size()), 0) >> (32 - e.code.size())) : 0u;
        ++k;
    }
    return k;
}

// generateHuffmanCode (Huffman.cpp:3-35): text -> per-symbol (code MSB-aligned, length) and the
// DHT order (symbols grouped by length 1..16, concatenated).  Returns number of distinct symbols.
int ref_generate_huffman(const int* text, int n, unsigned* code_msb /*[256]*/, unsigned char* length /*[256]*/,
                         unsigned char* counts16 /*[16]*/, unsigned char* dht_symbols /*[256]*/) {
    std::vector<int> t(text, text + n);
    auto r = generateHuffmanCode(t);
    std::memset(code_msb, 0, 256 * sizeof(unsigned));
    std::memset(length, 0, 256);
    for (auto& kv : r.first) { code_msb[kv.first & 255] = kv.second.code; length[kv.first & 255] = kv.second.length; }
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        counts16[len - 1] = static_cast<unsigned char>(r.second[len].size());
        for (int s : r.second[len]) dht_symbols[k++] = static_cast<unsigned char>(s);
    }
    return static_cast<int>(r.first.size());
}

// Bitstream semantics (BitstreamGeneric.hpp): push (value,nbits) MSB-mode pairs, optional fill(),
// stream out with FF->FF00 stuffing.  Returns number of bytes written to out (cap bytes max).
int ref_bitstream_pack(const unsigned* msb_aligned, const unsigned char* nbits, int n, int do_fill,
                       unsigned char* out, int cap, unsigned* size_bits) {
    Bitstream bs;
    for (int i = 0; i < n; ++i) bs.push_back(msb_aligned[i], nbits[i]);
    if (do_fill) bs.fill();
    if (size_bits) *size_bits = bs.size();
    std::ostringstream os;
    os << bs;
    const std::string s = os.str();
    const int m = static_cast<int>(s.size()) < cap ? static_cast<int>(s.size()) : cap;
    std::memcpy(out, s.data(), m);
    return static_cast<int>(s.size());
}

} // extern "C"
