// Host side of `Image` (see include/Image.hpp).  writeJPEG/encodeJPEG call the CUDA path through the C-ABI; the other
// stage methods are host code kept for API compatibility with the reference (src/Image.cpp).
#include "include/Image.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>
#include <stdexcept>

#include "../../include/jpgenc_b200.h"
#include "include/Dct.hpp"
#include "ppm_reader.hpp"

using boost::numeric::ublas::zero_matrix;

namespace {

// one GPU context per host thread, created on first use (a context is single-threaded by contract)
struct ThreadContext {
    jpgenc_ctx* ctx = nullptr;
    ~ThreadContext() { if (ctx) jpgenc_destroy(ctx); }
    jpgenc_ctx* get() {
        if (!ctx) {
            int device = 0;
            if (const char* env = std::getenv("JPGENC_DEVICE")) device = std::atoi(env);
            if (jpgenc_create(device, &ctx) != JPGENC_OK)
                throw std::runtime_error(std::string("jpgenc: ") + jpgenc_last_error(nullptr));
        }
        return ctx;
    }
};
thread_local ThreadContext g_gpu;

void check(jpgenc_ctx* c, int rc) {
    if (rc != JPGENC_OK) throw std::runtime_error(std::string("jpgenc: ") + jpgenc_last_error(c));
}

uint pad16(uint v) { return (v + 15u) & ~15u; }

}  // namespace

int fast_atoi(const char* str) {
    int v = 0;
    for (; *str; ++str) v = v * 10 + (*str - '0');
    return v;
}

Image::Image(uint w, uint h, ColorSpace color)
    : width(w), height(h), real_width(w), real_height(h), subsample_width(w), subsample_height(h),
      R(one), G(two), B(three), Y(one), Cb(two), Cr(three), color_space_type(color), one(h, w), two(h, w), three(h, w) {}

Image::Image(const Image& o)
    : width(o.width), height(o.height), real_width(o.real_width), real_height(o.real_height),
      subsample_width(o.subsample_width), subsample_height(o.subsample_height),
      R(one), G(two), B(three), Y(one), Cb(two), Cr(three), color_space_type(o.color_space_type),
      one(o.one), two(o.two), three(o.three), samples_(o.samples_), maxval_(o.maxval_) {}

Image::Image(Image&& o)
    : width(o.width), height(o.height), real_width(o.real_width), real_height(o.real_height),
      subsample_width(o.subsample_width), subsample_height(o.subsample_height),
      R(one), G(two), B(three), Y(one), Cb(two), Cr(three), color_space_type(o.color_space_type),
      one(std::move(o.one)), two(std::move(o.two)), three(std::move(o.three)), samples_(std::move(o.samples_)),
      maxval_(o.maxval_) {}

Image::~Image() {}

Image& Image::operator=(const Image& o) {
    if (this != &o) {
        one = o.one; two = o.two; three = o.three;
        width = o.width; height = o.height; real_width = o.real_width; real_height = o.real_height;
        subsample_width = o.subsample_width; subsample_height = o.subsample_height;
        color_space_type = o.color_space_type; samples_ = o.samples_; maxval_ = o.maxval_;
    }
    return *this;
}

Image& Image::operator=(Image&& o) {
    if (this != &o) {
        one = std::move(o.one); two = std::move(o.two); three = std::move(o.three);
        width = o.width; height = o.height; real_width = o.real_width; real_height = o.real_height;
        subsample_width = o.subsample_width; subsample_height = o.subsample_height;
        color_space_type = o.color_space_type; samples_ = std::move(o.samples_); maxval_ = o.maxval_;
    }
    return *this;
}

// ---------------------------------------------------------------------------------------------------------------
// loading
// ---------------------------------------------------------------------------------------------------------------
void encodePPMFile(std::string in, std::string out) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    jpgenc_ctx* c = g_gpu.get();
    const int rc = jpgenc_encode_ppm_file(c, in.c_str(), out.c_str());
    if (rc == JPGENC_ERR_IO || rc == JPGENC_ERR_FORMAT) {
        const std::string what = jpgenc_last_error(c);                  // the reference's texts (src/Image.cpp:428,450)
        throw std::runtime_error(what.find("Failed to open input") == 0 ? "Failed to open \"" + in + "\"" : what);
    }
    check(c, rc);
    const auto t1 = std::chrono::high_resolution_clock::now();
    std::cout << "PPM loading took 0 ms (streamed with the encode)\n";
    std::cout << "Encoding duration: " << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() << " ms" << std::endl;
}

Image loadPPM(std::string path) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<uint8_t> file, p3;
    if (jpgenc::slurp_file(path, &file) != JPGENC_OK) throw std::runtime_error("Failed to open \"" + path + "\"");
    jpgenc::PpmHeader h;
    if (jpgenc::parse_ppm_header(file.data(), file.size(), &h) != JPGENC_OK)
        throw std::runtime_error("Only P3 and P6 format is supported!");
    const uint8_t* px = nullptr;
    if (jpgenc::ppm_samples(file.data(), file.size(), h, &p3, &px) != JPGENC_OK)
        throw std::runtime_error("Only P3 and P6 format is supported!");

    Image img(h.width, h.height, Image::RGB);
    img.samples_.assign(px, px + static_cast<std::size_t>(h.width) * h.height * 3);
    img.maxval_ = h.maxval;
    const double scale = 255. / h.maxval;                     // white is maxval in the file, 255 in the planes
    const std::size_t n = static_cast<std::size_t>(h.width) * h.height;
    for (std::size_t i = 0; i < n; ++i) {
        img.R.data()[i] = px[3 * i] * scale;
        img.G.data()[i] = px[3 * i + 1] * scale;
        img.B.data()[i] = px[3 * i + 2] * scale;
    }
    // planes are padded to whole 16x16 MCUs by repeating the last column / row (the GPU path clamps instead)
    const uint w16 = pad16(h.width), h16 = pad16(h.height);
    if (w16 != h.width || h16 != h.height) {
        img.width = w16; img.height = h16; img.subsample_width = w16; img.subsample_height = h16;
        for (matrix<PixelDataType>* plane : {&img.one, &img.two, &img.three}) {
            plane->resize(h16, w16, true);
            for (uint y = 0; y < h16; ++y)
                for (uint x = 0; x < w16; ++x)
                    if (y >= h.height || x >= h.width)
                        (*plane)(y, x) = (*plane)(std::min(y, h.height - 1), std::min(x, h.width - 1));
        }
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    std::cout << "PPM loading took " << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() << " ms\n";
    return img;
}

// ---------------------------------------------------------------------------------------------------------------
// the hot path: GPU
// ---------------------------------------------------------------------------------------------------------------
// The 8-bit samples the planes hold, when they hold such: every real pixel k * (255 / maxval) for an integer k in
// [0, maxval], the padding a replication of the last column / row (what loadPPM produces, src/Image.cpp:465-530).  `cached`
// (loadPPM's copy of the file's samples) is only trusted after it has been compared with the planes: a caller may have
// edited R/G/B since (the reference encodes the planes, src/Image.cpp:831-846).
static bool planes_as_samples(const Image& img, const matrix<PixelDataType>* plane[3], uint maxval, std::vector<Byte>* cached) {
    const uint rw = img.real_width, rh = img.real_height;
    if (img.width != pad16(rw) || img.height != pad16(rh) || maxval == 0 || maxval > 255) return false;
    for (int k = 0; k < 3; ++k)
        if (plane[k]->size1() != img.height || plane[k]->size2() != img.width) return false;
    const double scale = 255. / maxval;
    const bool have = cached->size() == static_cast<std::size_t>(rw) * rh * 3;
    if (!have) cached->assign(static_cast<std::size_t>(rw) * rh * 3, 0);
    for (int k = 0; k < 3; ++k) {
        const matrix<PixelDataType>& p = *plane[k];
        for (uint y = 0; y < img.height; ++y) {
            const uint sy = std::min(y, rh - 1);
            for (uint x = 0; x < img.width; ++x) {
                const uint sx = std::min(x, rw - 1);
                const double v = p(y, x);
                Byte& b = (*cached)[(static_cast<std::size_t>(sy) * rw + sx) * 3 + k];
                if (y < rh && x < rw && !have) {
                    const double q = std::floor(v / scale + 0.5);
                    if (!(q >= 0 && q <= maxval)) return false;
                    b = static_cast<Byte>(q);
                }
                if (b * scale != v) return false;              // also checks the padding against the pixel it must replicate
            }
        }
    }
    return true;
}

std::vector<Byte> Image::encodeJPEG() {
    jpgenc_ctx* c = g_gpu.get();
    uint64_t need = 0;
    const matrix<PixelDataType>* plane[3] = {&one, &two, &three};
    bool as_samples = false;
    if (color_space_type == RGB) {
        as_samples = planes_as_samples(*this, plane, maxval_, &samples_);
        if (!as_samples && !samples_.empty()) {                // the cache is stale (edited planes): one more try without it
            samples_.clear();
            as_samples = planes_as_samples(*this, plane, maxval_, &samples_);
        }
    }
    if (as_samples) {
        // upload (in bands, overlapped with the first kernel) + encode; the scan stays on the device, its size comes back
        check(c, jpgenc_encode_rgb(c, samples_.data(), real_width, real_height, maxval_, nullptr, 0, &need));
    } else {
        // anything else the planes may hold -- edited or real-valued samples, an image that already is YCbCr (the reference
        // converts only an RGB image, src/Image.cpp:112-115, 839): the planes themselves go to the device
        samples_.clear();
        if (one.size1() != height || one.size2() != width || two.size1() != height || two.size2() != width || three.size1() != height ||
            three.size2() != width)
            throw std::runtime_error("writeJPEG: the three planes must have the image's (padded) size");
        check(c, jpgenc_encode_planes(c, &one.data()[0], &two.data()[0], &three.data()[0], width, height, real_width, real_height,
                                      color_space_type == YCbCr ? 1 : 0, nullptr, 0, &need));
    }
    std::vector<Byte> out(need);
    check(c, jpgenc_assemble_last(c, out.data(), out.size(), &need));   // headers + D2H of the scan + EOI
    return out;
}

void Image::writeJPEG(std::string file) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::cout << "Processing image size: " << real_width << "x" << real_height << std::endl;
    const std::vector<Byte> jpeg = encodeJPEG();
    std::ofstream out(file, std::ios::binary);
    out.write(reinterpret_cast<const char*>(jpeg.data()), static_cast<std::streamsize>(jpeg.size()));
    const auto t1 = std::chrono::high_resolution_clock::now();
    std::cout << "Encoding duration: " << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() << " ms" << std::endl;
}

// ---------------------------------------------------------------------------------------------------------------
// stage methods (host; API compatibility)
// ---------------------------------------------------------------------------------------------------------------
Image Image::convertToColorSpace(ColorSpace target) const {
    if (target == color_space_type) return *this;
    Image out(*this);
    const std::size_t n = static_cast<std::size_t>(width) * height;
    if (target == YCbCr) {
        // float constants, double arithmetic, level shift folded in (reference src/Image.cpp:131-143)
        const float ky[3] = {.299f, .587f, .114f}, kb[3] = {-.1687f, -.3312f, .5f}, kr[3] = {.5f, -.4186f, -.0813f};
        const float off[3] = {.0f, 128.f, 128.f};
        for (std::size_t i = 0; i < n; ++i) {
            const double r = one.data()[i], g = two.data()[i], b = three.data()[i];
            out.one.data()[i] = off[0] + (ky[0] * r + ky[1] * g + ky[2] * b) - 128;
            out.two.data()[i] = off[1] + (kb[0] * r + kb[1] * g + kb[2] * b) - 128;
            out.three.data()[i] = off[2] + (kr[0] * r + kr[1] * g + kr[2] * b) - 128;
        }
    } else {
        const float kr[3] = {1.f, .0f, 1.402f}, kg[3] = {1.f, -.344f, -.714f}, kb[3] = {1.f, 1.772f, .0f};
        for (std::size_t i = 0; i < n; ++i) {
            const double y = one.data()[i] + 128, cb = two.data()[i] + 128, cr = three.data()[i] + 128;
            out.one.data()[i] = kr[0] * y + kr[1] * cb + kr[2] * cr;
            out.two.data()[i] = kg[0] * y + kg[1] * cb + kg[2] * cr;
            out.three.data()[i] = kb[0] * y + kb[1] * cb + kb[2] * cr;
        }
    }
    out.color_space_type = target;
    return out;
}

static bool g_stages_on_device = false;
void Image::stagesOnDevice(bool on) { g_stages_on_device = on; }

void Image::applySubsampling(SubsamplingMode mode) {
    if (mode == S444) return;
    if (g_stages_on_device) {
        // enum order == the C-ABI's mode numbers (S444, S422, S411, S420, S420_m, S420_lm)
        jpgenc_ctx* c = g_gpu.get();
        uint32_t ow = 0, oh = 0;
        for (matrix<PixelDataType>* plane : {&three, &two}) {             // Cr first, then Cb, as the reference does
            const uint32_t w = static_cast<uint32_t>(plane->size2()), h = static_cast<uint32_t>(plane->size1());
            check(c, jpgenc_stage_subsample_dims(static_cast<int>(mode), w, h, &ow, &oh));
            matrix<PixelDataType> small(oh, ow);
            check(c, jpgenc_stage_subsample(c, &plane->data()[0], w, h, static_cast<int>(mode), &small.data()[0]));
            *plane = small;
        }
        subsample_width = ow;
        subsample_height = oh;
        return;
    }
    // horizontal step, vertical step, taps per row, whether the second row is averaged in, divisor
    int hstep = 2, vstep = 2, taps = 1;
    bool second_row = false;
    double divisor = 1;
    switch (mode) {
        case S422: hstep = 2; vstep = 1; break;
        case S411: hstep = 4; vstep = 1; break;
        case S420: break;
        case S420_m: taps = 2; second_row = true; divisor = 4; break;
        case S420_lm: second_row = true; divisor = 2; break;
        default: break;
    }
    subsample_width = width / hstep;
    subsample_height = height / vstep;
    for (matrix<PixelDataType>* plane : {&three, &two}) {             // Cr first, then Cb, as the reference does
        matrix<PixelDataType> small(plane->size1() / vstep, plane->size2() / hstep);
        for (std::size_t y = 0; y < small.size1(); ++y)
            for (std::size_t x = 0; x < small.size2(); ++x) {
                const std::size_t sy = y * vstep, sx = x * hstep;
                PixelDataType top = 0;
                for (int t = 0; t < taps; ++t) top += 1 * (*plane)(sy, sx + t);
                if (second_row) {
                    PixelDataType bottom = 0;
                    for (int t = 0; t < taps; ++t) bottom += 1 * (*plane)(sy + 1, sx + t);
                    small(y, x) = (top + bottom) / divisor;
                } else {
                    small(y, x) = top;
                }
            }
        *plane = small;
    }
}

void Image::applyDCT(DCTMode mode) {
    std::function<void(const matrix_range<matrix<PixelDataType>>&, matrix_range<matrix<PixelDataType>>&)> fn;
    switch (mode) {
        case Simple: fn = dctDirect; break;
        case Matrix: fn = dctMat; break;
        default: fn = dctArai; break;
    }
    const matrix<PixelDataType>* src[3] = {&one, &two, &three};
    matrix<PixelDataType>* dst[3] = {&DctY, &DctCb, &DctCr};
    if (g_stages_on_device) {
        jpgenc_ctx* c = g_gpu.get();
        for (int p = 0; p < 3; ++p) {
            *dst[p] = matrix<PixelDataType>(src[p]->size1(), src[p]->size2());
            check(c, jpgenc_stage_dct(c, &src[p]->data()[0], static_cast<uint32_t>(src[p]->size2()), static_cast<uint32_t>(src[p]->size1()),
                                      mode == Simple ? 0 : mode == Matrix ? 1 : 2, &dst[p]->data()[0]));
        }
        return;
    }
    for (int p = 0; p < 3; ++p) {
        *dst[p] = matrix<PixelDataType>(src[p]->size1(), src[p]->size2());
        for (std::size_t y = 0; y + 8 <= src[p]->size1(); y += 8)
            for (std::size_t x = 0; x + 8 <= src[p]->size2(); x += 8) {
                const matrix_range<matrix<PixelDataType>> in(*src[p], range(y, y + 8), range(x, x + 8));
                matrix_range<matrix<PixelDataType>> out(*dst[p], range(y, y + 8), range(x, x + 8));
                fn(in, out);
            }
    }
}

void Image::applyQuantization(const matrix<Byte>& qy, const matrix<Byte>& qc) {
    const matrix<PixelDataType>* src[3] = {&DctY, &DctCb, &DctCr};
    matrix<int>* dst[3] = {&QY, &QCb, &QCr};
    for (int p = 0; p < 3; ++p) {
        const mat table(p == 0 ? qy : qc);
        *dst[p] = matrix<int>(src[p]->size1(), src[p]->size2());
        for (std::size_t y = 0; y + 8 <= src[p]->size1(); y += 8)
            for (std::size_t x = 0; x + 8 <= src[p]->size2(); x += 8) {
                const mat block(matrix_range<matrix<PixelDataType>>(*src[p], range(y, y + 8), range(x, x + 8)));
                matrix_range<matrix<int>>(*dst[p], range(y, y + 8), range(x, x + 8)).assign(quantize(block, table));
            }
    }
}

void Image::applyDCdifferenceCoding() {
    int prev = 0;                                               // luma predicts along the MCU order Y00 Y01 Y10 Y11
    for (std::size_t y = 0; y + 16 <= QY.size1(); y += 16)
        for (std::size_t x = 0; x + 16 <= QY.size2(); x += 16)
            for (int k = 0; k < 4; ++k) {
                int& dc = QY(y + 8 * (k >> 1), x + 8 * (k & 1));
                const int cur = dc;
                dc = cur - prev;
                prev = cur;
            }
    for (matrix<int>* plane : {&QCb, &QCr}) {                   // chroma predicts in raster order
        prev = 0;
        for (std::size_t y = 0; y + 8 <= plane->size1(); y += 8)
            for (std::size_t x = 0; x + 8 <= plane->size2(); x += 8) {
                const int cur = (*plane)(y, x);
                (*plane)(y, x) = cur - prev;
                prev = cur;
            }
    }
}

void Image::doZigZagSorting() {}

void Image::doRLEandCategoryCoding() {
    const matrix<int>* src[3] = {&QY, &QCb, &QCr};
    matrix<std::vector<Category_Code>>* dst[3] = {&CategoryCodeY, &CategoryCodeCb, &CategoryCodeCr};
    for (int p = 0; p < 3; ++p) {
        *dst[p] = matrix<std::vector<Category_Code>>(src[p]->size1() / 8, src[p]->size2() / 8);
        for (std::size_t by = 0; by < dst[p]->size1(); ++by)
            for (std::size_t bx = 0; bx < dst[p]->size2(); ++bx) {
                const matrix<int> block(matrix_range<matrix<int>>(*src[p], range(by * 8, by * 8 + 8), range(bx * 8, bx * 8 + 8)));
                (*dst[p])(by, bx) = encode_category(RLE_AC(block));
            }
    }
}

void Image::doHuffmanEncoding(SymbolCodeMap& Y_DC, SymbolCodeMap& Y_AC, SymbolCodeMap& C_DC, SymbolCodeMap& C_AC) {
    matrix<std::vector<Category_Code>>* src[3] = {&CategoryCodeY, &CategoryCodeCb, &CategoryCodeCr};
    matrix<Bitstream>* dst[3] = {&BitstreamY, &BitstreamCb, &BitstreamCr};
    for (int p = 0; p < 3; ++p) {
        SymbolCodeMap& dc = p == 0 ? Y_DC : C_DC;
        SymbolCodeMap& ac = p == 0 ? Y_AC : C_AC;
        *dst[p] = matrix<Bitstream>(src[p]->size1(), src[p]->size2());
        for (std::size_t by = 0; by < src[p]->size1(); ++by)
            for (std::size_t bx = 0; bx < src[p]->size2(); ++bx) {
                Bitstream bits;
                std::vector<Category_Code>& entries = (*src[p])(by, bx);
                for (std::size_t i = 0; i < entries.size(); ++i) {
                    const Code& code = (i == 0 ? dc : ac)[entries[i].symbol];
                    bits.push_back(code.code, code.length);
                    bits << entries[i].code;
                }
                (*dst[p])(by, bx) = bits;
            }
    }
}
