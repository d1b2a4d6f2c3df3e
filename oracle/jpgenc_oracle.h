/*
 * jpgenc_oracle.h — CPU restatement of the Nuos/jpgEnc encode path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library, and
 * only as the checker.  The product (jpgenc_b200/) never links, imports or executes it.
 *
 * Parity status: PINNED.  Every stage below is checked (tests/test_oracle_vs_reference.py, run in the
 * build container where /root/reference exists) against the reference encoder itself, built by
 * oracle/build_ref.sh into oracle/_ref/, and against the known-answer vectors of the reference's own
 * unit tests (tests/test_oracle_known_answers.py); whole files are byte-identical to the reference on
 * all 21 fixture PPMs and on the synthetic 512x512 / 1920x1080 / 3840x2160 pins of SURVEY.md 8(c).
 *
 * All `file:line` citations are into the reference tree (/root/reference).
 */
#ifndef JPGENC_ORACLE_H
#define JPGENC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- PPM (Image.cpp:326-473) ------------------------------------------------------------- */
typedef struct {
    int      magic;      /* 3 or 6 */
    uint32_t width, height, maxval;
    size_t   payload;    /* offset of the first sample (byte after the single whitespace that ends maxval) */
} jo_ppm_header;

/* returns 0, -1 = not P3/P6 (reference throws, Image.cpp:449-450), -2 = truncated */
int jo_ppm_parse(const uint8_t* file, size_t n, jo_ppm_header* h);
/* decode samples to interleaved u8 RGB (raw sample values, NOT scaled); P3 or P6 */
int jo_ppm_samples(const uint8_t* file, size_t n, const jo_ppm_header* h, uint8_t* rgb);

/* ---- per-block numeric kernels -------------------------------------------------------------- */
/* Dct.hpp:47-215 — Arai/AAN 8x8, double, reference operation order; in/out row-major (v,u) */
void jo_dct_arai(const double in[64], double out[64]);
/* Dct.hpp:238-262 / 264-276 — the two non-default variants (test parity to 1e-5 only) */
void jo_dct_direct(const double in[64], double out[64]);
void jo_dct_matrix(const double in[64], double out[64]);
/* Coding.hpp:84-97 */
void jo_quantize(const double in[64], const uint8_t table[64], int32_t out[64]);
/* Coding.hpp:57-81 — natural index of the i-th zigzag coefficient */
int jo_zigzag_index(int i);
/* Coding.hpp:197-230 — size category and magnitude bits */
void jo_category(int value, int* category, uint32_t* bits);
/* Coding.hpp:148-183 + 265-283 — natural-order block -> symbol list ([0] is the DC entry) */
int jo_block_symbols(const int32_t natural[64], uint8_t sym[64], uint32_t bits[64], uint8_t nbits[64]);

/* ---- Huffman (Huffman.cpp:3-66, Huffman.hpp:114-174) ------------------------------------------- */
typedef struct {
    uint32_t code_msb[256];   /* Code::code — MSB-aligned (Huffman.hpp:26-34) */
    uint8_t  length[256];     /* 0 = symbol absent */
    uint8_t  counts[16];      /* DHT: number of codes of length 1..16 */
    uint8_t  symbols[256];    /* DHT: symbols in the reference's SymbolsPerLength order */
    int      nsymbols;
} jo_huff_table;

/* from a symbol text (exactly generateHuffmanCode) */
void jo_huffman_from_text(const int32_t* text, size_t n, jo_huff_table* t);
/* from counts + first-occurrence order (what a histogram kernel can supply).  order[] lists the
 * distinct symbols in order of first appearance in the text. */
void jo_huffman_from_hist(const uint32_t count[256], const uint8_t* order, int ndistinct, jo_huff_table* t);

/* ---- bit container (BitstreamGeneric.hpp) ------------------------------------------------------- */
typedef struct {
    uint8_t* bytes;
    size_t   cap;
    uint64_t nbits;
} jo_bits;
void   jo_bits_init(jo_bits* b);
void   jo_bits_free(jo_bits* b);
void   jo_bits_push_msb(jo_bits* b, uint32_t msb_aligned, int n);   /* push_back(data,n)      :182-195 */
void   jo_bits_push_lsb(jo_bits* b, uint32_t value, int n);          /* push_back_LSB_mode     :197-210 */
void   jo_bits_fill(jo_bits* b);                                     /* fill()                 :242-248 */
size_t jo_bits_stuffed_size(const jo_bits* b);                       /* operator<<(ostream&)   :213-224 */
size_t jo_bits_write_stuffed(const jo_bits* b, uint8_t* dst);

/* ---- pipeline stages on whole images ---------------------------------------------------------- */
/* Geometry: W16/H16 = dims padded up to multiples of 16 (Image.cpp:479-489); mcu_w = W16/16 ... */
static inline uint32_t jo_pad16(uint32_t v) { return (v + 15u) & ~15u; }

/* Forward path for MCU rows [mcu_y0, mcu_y1): RGB u8 -> (x 255/maxval) -> pad by edge replication ->
 * YCbCr (Image.cpp:131-143) -> S420_m (Image.cpp:198-235) -> dctArai -> quantize.
 * Output: MCU-ordered blocks (Y00,Y01,Y10,Y11,Cb,Cr per MCU, MCUs raster), each 64 int16 in ZIGZAG
 * order, DC NOT differenced.  out must hold (mcu_y1-mcu_y0)*mcu_w*6*64 int16. */
void jo_forward_mcu_rows(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint32_t maxval,
                         const uint8_t qy[64], const uint8_t qc[64],
                         uint32_t mcu_y0, uint32_t mcu_y1, int16_t* out);

/* Same arithmetic, but dumping the intermediate planes the reference holds (any pointer may be NULL):
 * ycc: Y (H16*W16) / Cb,Cr (H16/2*W16/2) doubles; dct_*: same shapes; q_*: int32 natural order planes. */
void jo_forward_planes(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint32_t maxval,
                       const uint8_t qy[64], const uint8_t qc[64],
                       double* y, double* cb, double* cr, double* dct_y, double* dct_cb, double* dct_cr,
                       int32_t* q_y, int32_t* q_cb, int32_t* q_cr);

/* Stage methods on one plane: Image::subsample for every SubsamplingMode (Image.cpp:198-319; mode = enum order S444, S422, S411,
 * S420, S420_m, S420_lm) and Image::applyDCT (Image.cpp:540-595; mode 0 Simple, 1 Matrix, 2 Arai) */
void jo_subsample_dims(int mode, uint32_t w, uint32_t h, uint32_t* ow, uint32_t* oh);
void jo_subsample_plane(const double* in, uint32_t w, uint32_t h, int mode, double* out);
void jo_dct_plane(const double* in, uint32_t w, uint32_t h, int mode, double* out);
/* the 8x8 DCT basis A of Dct.hpp:217-236 (row-major) */
void jo_dct_basis(double a[64]);

/* The same from three H16*W16 planes of doubles (an Image assembled or edited in memory): R,G,B, or with ycbcr != 0 planes that
 * already are level-shifted Y,Cb,Cr (Image.cpp:112-115: not converted again). */
void jo_forward_from_planes(const double* p0, const double* p1, const double* p2, uint32_t W16, uint32_t H16, int ycbcr,
                            const uint8_t qy[64], const uint8_t qc[64], int16_t* out);

/* planar natural-order int32 (the reference's QY/QCb/QCr) <-> MCU-ordered zigzag int16 */
void jo_planes_to_mcu(const int32_t* q_y, const int32_t* q_cb, const int32_t* q_cr,
                      uint32_t mcu_w, uint32_t mcu_h, int16_t* out);

/* Entropy statistics in the reference's TEXT order (Image.cpp:888-906): table 0=Y_DC 1=Y_AC 2=C_DC 3=C_AC.
 * first_pos = order-preserving key of the first occurrence in that table's text, UINT64_MAX if absent:
 * (block index in text order)*256 + k, k = 0 for DC, 2p for a ZRL before zigzag position p, 2p+1 for the symbol of
 * position p, 129 for EOB. */
void jo_symbol_stats(const int16_t* mcu_blocks, uint32_t mcu_w, uint32_t mcu_h,
                     uint32_t count[4][256], uint64_t first_pos[4][256]);

/* Full entropy coding of MCU-ordered coefficients (DC differencing Image.cpp:638-678, RLE/category,
 * table build, Huffman encode Image.cpp:737-829, interleave + fill Image.cpp:957-970).
 * tables[4] receives the tables; bits receives the un-stuffed, 1-padded scan. */
void jo_entropy_encode(const int16_t* mcu_blocks, uint32_t mcu_w, uint32_t mcu_h,
                       jo_huff_table tables[4], jo_bits* bits);

/* Header bytes SOI..SOS (Image.cpp:933-954, JpegSegments.hpp); returns length (dst may be NULL) */
size_t jo_write_headers(uint32_t real_w, uint32_t real_h, const uint8_t qy[64], const uint8_t qc[64],
                        const jo_huff_table tables[4], uint8_t* dst);

/* Whole file from a PPM in memory == main.cpp.  *out is malloc'd; caller frees with jo_free.
 * returns 0 or the jo_ppm_parse error. */
int  jo_encode_ppm(const uint8_t* file, size_t n, uint8_t** out, size_t* out_n);
/* Whole file from raw RGB (maxval 255) */
int  jo_encode_rgb(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint8_t** out, size_t* out_n);
/* Whole file from planes of doubles == Image::writeJPEG on such an Image */
int  jo_encode_planes(const double* p0, const double* p1, const double* p2, uint32_t W16, uint32_t H16, uint32_t real_w,
                      uint32_t real_h, int ycbcr, uint8_t** out, size_t* out_n);
void jo_free(void* p);

extern const uint8_t jo_qtable_luma[64];     /* Image.cpp:850-859 */
extern const uint8_t jo_qtable_chroma[64];   /* Image.cpp:860-869 */

#ifdef __cplusplus
}
#endif
#endif
