// Test driver for the C++ host mirror (jpgenc_b200/host): exercises the reference's public surface as a caller would.
// Compiled by tests/test_host_mirror.py with -fno-access-control (the stage results are private members, as in the
// reference).  Modes:
//   stages <in.ppm> <out.bin>   loadPPM -> convertToColorSpace -> applySubsampling(S420_m) -> applyDCT(Arai) ->
//                               applyQuantization -> applyDCdifferenceCoding/doRLEandCategoryCoding/doHuffmanEncoding;
//                               dumps Y, Cb (doubles), DctY (doubles), QY, QCb, QCr (int32, before DC differencing)
//                               and the per-stage host scan (stuffed) for comparison with the golden vectors   [CPU]
//   stages_device <in.ppm> <out.bin>   the same with Image::stagesOnDevice(true): applySubsampling / applyDCT on the GPU [GPU]
//   huffman                     generateHuffmanCode(text) == jpgenc_build_huffman(histogram, first positions) [CPU]
//   segments                    header bytes through the Segment:: classes == jpgenc_write_headers           [CPU]
//   encode <in.ppm> <out.jpg>   loadPPM + Image::writeJPEG                                                    [GPU]
//   encode_ycc / encode_edited8 / encode_edited <in.ppm> <out.jpg>   writeJPEG on an image that already is YCbCr / whose
//                               planes were edited after loading (src/Image.cpp:831-846 encodes the planes)     [GPU]
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>

#include "Dct.hpp"
#include "Image.hpp"
#include "JpegSegments.hpp"
#include "../../include/jpgenc_b200.h"

static const Byte kLuma[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                               14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                               18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                               49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const Byte kChroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

template <class T>
static void dump(std::ofstream& f, const matrix<T>& m) {
    const uint32_t dims[2] = {static_cast<uint32_t>(m.size1()), static_cast<uint32_t>(m.size2())};
    f.write(reinterpret_cast<const char*>(dims), sizeof dims);
    for (std::size_t r = 0; r < m.size1(); ++r)
        for (std::size_t c = 0; c < m.size2(); ++c) {
            const T v = m(r, c);
            f.write(reinterpret_cast<const char*>(&v), sizeof v);
        }
}

static matrix<Byte> table(const Byte* q) {
    matrix<Byte> m(8, 8);
    for (int i = 0; i < 64; ++i) m(i / 8, i % 8) = q[i];
    return m;
}

static int stages(const char* in, const char* out) {
    Image img = loadPPM(in);
    img = img.convertToColorSpace(Image::YCbCr);
    img.applySubsampling(Image::S420_m);
    std::ofstream f(out, std::ios::binary);
    dump(f, img.Y);
    dump(f, img.Cb);
    img.applyDCT(Image::Arai);
    dump(f, img.DctY);
    img.applyQuantization(table(kLuma), table(kChroma));
    dump(f, img.QY);
    dump(f, img.QCb);
    dump(f, img.QCr);
    // the rest of the stage API, driven the way Image::writeJPEG drives it in the reference (src/Image.cpp:878-971)
    img.applyDCdifferenceCoding();
    img.doRLEandCategoryCoding();
    std::vector<int> text[4];                                   // Y_DC, Y_AC, C_DC, C_AC
    const matrix<std::vector<Category_Code>>* planes[3] = {&img.CategoryCodeY, &img.CategoryCodeCb, &img.CategoryCodeCr};
    for (int p = 0; p < 3; ++p)
        for (std::size_t r = 0; r < planes[p]->size1(); ++r)
            for (std::size_t c = 0; c < planes[p]->size2(); ++c) {
                const std::vector<Category_Code>& e = (*planes[p])(r, c);
                for (std::size_t i = 0; i < e.size(); ++i) text[(p ? 2 : 0) + (i ? 1 : 0)].push_back(e[i].symbol);
            }
    SymbolCodeMap maps[4];
    for (int t = 0; t < 4; ++t) maps[t] = generateHuffmanCode(text[t]).first;
    img.doHuffmanEncoding(maps[0], maps[1], maps[2], maps[3]);
    Bitstream scan;
    for (std::size_t r = 0; r < img.BitstreamCb.size1(); ++r)
        for (std::size_t c = 0; c < img.BitstreamCb.size2(); ++c) {
            scan << img.BitstreamY(2 * r, 2 * c) << img.BitstreamY(2 * r, 2 * c + 1) << img.BitstreamY(2 * r + 1, 2 * c)
                 << img.BitstreamY(2 * r + 1, 2 * c + 1) << img.BitstreamCb(r, c) << img.BitstreamCr(r, c);
        }
    scan.fill();
    std::ostringstream s;
    s << scan;
    const std::string bytes = s.str();
    const uint32_t dims[2] = {1, static_cast<uint32_t>(bytes.size())};
    f.write(reinterpret_cast<const char*>(dims), sizeof dims);
    f.write(bytes.data(), static_cast<std::streamsize>(bytes.size()));
    return 0;
}

static int huffman() {
    std::mt19937 rng(12345);
    int bad = 0;
    for (int round = 0; round < 400; ++round) {
        const int distinct = 1 + rng() % 60, n = distinct + rng() % 3000;
        std::vector<int> alphabet(distinct), text(n);
        for (int& a : alphabet) a = rng() % 256;
        for (int& t : text) t = alphabet[(rng() % distinct) * (rng() % distinct) / distinct];   // skewed
        uint32_t count[256] = {0};
        uint64_t first[256];
        std::memset(first, 0xFF, sizeof first);
        for (std::size_t i = 0; i < text.size(); ++i)
            if (count[text[i]]++ == 0) first[text[i]] = i;
        jpgenc_huff_table t;
        if (jpgenc_build_huffman(count, first, &t) != JPGENC_OK) { ++bad; continue; }
        const auto pair = generateHuffmanCode(text);
        int k = 0;
        for (int len = 1; len <= 16; ++len) {
            if (pair.second[len].size() != t.counts[len - 1]) { ++bad; break; }
            for (int s : pair.second[len])
                if (t.symbols[k++] != s) { ++bad; break; }
        }
        for (const auto& kv : pair.first)
            if (kv.second.code != t.code_msb[kv.first] || kv.second.length != t.length[kv.first]) { ++bad; break; }
        const std::vector<int> back = huffmanDecode(huffmanEncode(text, pair.first), pair.first);
        if (back != text) ++bad;
    }
    std::cout << "huffman mismatches " << bad << std::endl;
    return bad != 0;
}

static int segments() {
    // the header the Segment:: classes write == the header the C-ABI writes, for tables built from a toy text
    std::vector<int> texts[4] = {{0, 1, 2, 2, 3, 3, 3}, {0, 0, 0, 1, 17, 0xF0, 33, 2, 2}, {0, 1, 1}, {0, 0, 17, 1}};
    jpgenc_huff_table tabs[4];
    std::pair<SymbolCodeMap, SymbolsPerLength> gen[4];
    for (int t = 0; t < 4; ++t) {
        uint32_t count[256] = {0};
        uint64_t first[256];
        std::memset(first, 0xFF, sizeof first);
        for (std::size_t i = 0; i < texts[t].size(); ++i)
            if (count[texts[t][i]]++ == 0) first[texts[t][i]] = i;
        jpgenc_build_huffman(count, first, &tabs[t]);
        gen[t] = generateHuffmanCode(texts[t]);
    }
    const uint32_t w = 1234, h = 777;
    std::vector<uint8_t> want(jpgenc_write_headers(w, h, kLuma, kChroma, tabs, nullptr));
    jpgenc_write_headers(w, h, kLuma, kChroma, tabs, want.data());

    using namespace Segment;
    std::ostringstream out;
    sDQT dqt_y, dqt_c;
    std::vector<Byte> zz_y = zigzag<Byte>(table(kLuma)), zz_c = zigzag<Byte>(table(kChroma));
    dqt_y.pushQuantizationTable(zz_y, ComponentSetup::Zero);
    dqt_c.pushQuantizationTable(zz_c, ComponentSetup::One);
    sSOF0 sof;
    sof.setImageSizeX(static_cast<short>(w)).setImageSizeY(static_cast<short>(h))
        .setupY(ComponentSetup::NoSubSampling, ComponentSetup::Zero)
        .setupCb(ComponentSetup::Half, ComponentSetup::One)
        .setupCr(ComponentSetup::Half, ComponentSetup::One);
    sDHT dht[4];
    dht[0].pushCodeData(gen[0].second, sDHT::DC, sDHT::First);
    dht[1].pushCodeData(gen[1].second, sDHT::AC, sDHT::First);
    dht[2].pushCodeData(gen[2].second, sDHT::DC, sDHT::Second);
    dht[3].pushCodeData(gen[3].second, sDHT::AC, sDHT::Second);
    sSOS sos;
    sos.setupY(sDHT::First, sDHT::First).setupCb(sDHT::Second, sDHT::Second).setupCr(sDHT::Second, sDHT::Second);
    out << sSOI() << sAPP0() << dqt_y << dqt_c << sof << dht[0] << dht[1] << dht[2] << dht[3] << sos;
    const std::string got = out.str();
    const bool same = got.size() == want.size() && std::memcmp(got.data(), want.data(), want.size()) == 0;
    std::cout << "segments " << got.size() << " bytes, identical " << same << "; sizeof(sAPP0)=" << sizeof(sAPP0)
              << " sizeof(sSOF0)=" << sizeof(sSOF0) << std::endl;
    return !(same && sizeof(sAPP0) == 18 && sizeof(sSOF0) == 19);
}

int main(int argc, char** argv) {
    try {
        const std::string mode = argc > 1 ? argv[1] : "";
        if (mode == "stages" && argc == 4) return stages(argv[2], argv[3]);
        if (mode == "stages_device" && argc == 4) {           // the same stage sequence with applySubsampling / applyDCT on the GPU
            Image::stagesOnDevice(true);
            return stages(argv[2], argv[3]);
        }
        if (mode == "huffman") return huffman();
        if (mode == "segments") return segments();
        if (mode == "encode" && argc == 4) {
            Image img = loadPPM(argv[2]);
            img.writeJPEG(argv[3]);
            return 0;
        }
        if (mode == "encode_ycc" && argc == 4) {             // a caller converts first: writeJPEG must not convert again
            Image img = loadPPM(argv[2]).convertToColorSpace(Image::YCbCr);
            img.writeJPEG(argv[3]);
            return 0;
        }
        if (mode == "encode_edited8" && argc == 4) {         // planes edited after loading, still 8-bit samples
            Image img = loadPPM(argv[2]);
            img.R(1, 2) = 255; img.G(0, 0) = 0; img.B(2, 1) = 17;
            img.writeJPEG(argv[3]);
            return 0;
        }
        if (mode == "encode_edited" && argc == 4) {          // planes edited after loading: a real-valued sample, a padding sample
            Image img = loadPPM(argv[2]);
            img.R(1, 2) += 0.37;
            img.B(img.height - 1, img.width - 1) = 99.5;
            img.writeJPEG(argv[3]);
            return 0;
        }
        std::cerr << "usage: mirror_probe stages|huffman|segments|encode|encode_ycc|encode_edited8|encode_edited ..." << std::endl;
        return 2;
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what() << std::endl;
        return 3;
    }
}
