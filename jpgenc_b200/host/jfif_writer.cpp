// JFIF marker segments of the encode path, as a flat byte writer.
//
// Emits the same bytes as the reference's `jpeg << sSOI() << sAPP0() << sDQT()... << sSOS()` chain
// (src/Image.cpp:933-954; include/JpegSegments.hpp:55-377).  The reference serialises packed structs; here the
// segments are written field by field into a caller-supplied buffer so the header can be laid down directly in
// front of the scan bytes that come back from the GPU.
#include <cstdint>
#include <cstring>
#include <initializer_list>

#include "../../include/jpgenc_b200.h"

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

class ByteSink {
public:
    explicit ByteSink(uint8_t* dst) : dst_(dst), n_(0) {}
    void u8(unsigned v) { if (dst_) dst_[n_] = static_cast<uint8_t>(v); ++n_; }
    void be16(unsigned v) { u8(v >> 8); u8(v); }
    void marker(unsigned code, unsigned payload_len) { be16(0xFF00u | code); be16(payload_len); }
    size_t size() const { return n_; }
private:
    uint8_t* dst_;
    size_t n_;
};

}  // namespace

extern "C" size_t jpgenc_write_headers(uint32_t real_w, uint32_t real_h, const uint8_t qy[64], const uint8_t qc[64],
                                       const jpgenc_huff_table tables[4], uint8_t* dst) {
    ByteSink o(dst);
    o.be16(0xFFD8);                                            // SOI
    o.marker(0xE0, 16);                                        // APP0 "JFIF", rev 1.1, no units, density 1x1, no thumbnail
    for (char ch : {'J', 'F', 'I', 'F', '\0'}) o.u8(static_cast<unsigned char>(ch));
    o.u8(1); o.u8(1); o.u8(0); o.be16(1); o.be16(1); o.u8(0); o.u8(0);
    for (int id = 0; id < 2; ++id) {                           // one DQT segment per table, entries in zigzag order
        const uint8_t* q = id ? qc : qy;
        o.marker(0xDB, 2 + 65);
        o.u8(id);
        for (int i = 0; i < 64; ++i) o.u8(q[kZigzag[i]]);
    }
    o.marker(0xC0, 8 + 3 * 3);                                 // SOF0: 8-bit, height, width, 3 components
    o.u8(8);
    o.be16(real_h & 0xFFFFu);
    o.be16(real_w & 0xFFFFu);
    o.u8(3);
    o.u8(1); o.u8(0x22); o.u8(0);                              // Y : 2x2 sampling, quant table 0
    o.u8(2); o.u8(0x11); o.u8(1);                              // Cb: 1x1, table 1
    o.u8(3); o.u8(0x11); o.u8(1);                              // Cr: 1x1, table 1
    const uint8_t ht_info[4] = {0x00, 0x10, 0x01, 0x11};       // (class << 4) | destination for Y_DC, Y_AC, C_DC, C_AC
    for (int t = 0; t < 4; ++t) {
        unsigned nsym = 0;
        for (int i = 0; i < 16; ++i) nsym += tables[t].counts[i];
        o.marker(0xC4, 2 + 17 + nsym);
        o.u8(ht_info[t]);
        for (int i = 0; i < 16; ++i) o.u8(tables[t].counts[i]);
        for (unsigned i = 0; i < nsym; ++i) o.u8(tables[t].symbols[i]);
    }
    o.marker(0xDA, 12);                                        // SOS: Y -> tables 0/0, Cb and Cr -> 1/1, Ss=0 Se=63 Ah/Al=0
    o.u8(3);
    o.u8(1); o.u8(0x00);
    o.u8(2); o.u8(0x11);
    o.u8(3); o.u8(0x11);
    o.u8(0x00); o.u8(0x3F); o.u8(0x00);
    return o.size();
}
