// `Image` and loadPPM with the reference's public interface (include/Image.hpp:28-115).
//
// Image::writeJPEG is the hot path and runs on the B200 through the C-ABI in include/jpgenc_b200.h: the 8-bit samples
// loadPPM read are uploaded as they are, colour conversion / subsampling / DCT / quantisation / entropy coding / byte
// stuffing happen on the device, and the host only builds the four Huffman tables and writes the header.  There is no
// CPU fallback for it: without a GPU it throws std::runtime_error.
// The individual stage methods (convertToColorSpace, applySubsampling, applyDCT, ...) keep the reference's contract
// (they transform the planes held by the object, in pipeline order) as plain host code for callers and tests that
// drive the stages one by one; writeJPEG does not use them.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "Coding.hpp"
#include "Huffman.hpp"
#include "matrix_types.hpp"

typedef unsigned int uint;
typedef uint8_t Byte;
typedef double PixelDataType;

class Image;

// P3 or P6; throws std::runtime_error when the file cannot be opened or is not a PPM (reference src/Image.cpp:427-450)
Image loadPPM(std::string path);
// loadPPM(in).writeJPEG(out) without materialising the image on the host: a binary PPM is streamed from the file to the
// GPU band by band (jpgenc_encode_ppm_file).  Same file bytes, same exceptions; what the bundled command line uses.
void encodePPMFile(std::string in, std::string out);
int fast_atoi(const char* str);

class Image {
public:
    enum ColorSpace { RGB, YCbCr };
    enum SubsamplingMode { S444, S422, S411, S420, S420_m, S420_lm };
    enum DCTMode { Simple, Matrix, Arai };

    explicit Image(uint w, uint h, ColorSpace color);
    Image(const Image& other);
    Image(Image&& other);
    ~Image();
    Image& operator=(const Image& other);
    Image& operator=(Image&& other);

    Image convertToColorSpace(ColorSpace target_space) const;
    void applySubsampling(SubsamplingMode mode);
    void applyDCT(DCTMode mode);
    void applyQuantization(const matrix<Byte>& q_table_y, const matrix<Byte>& q_table_c);
    void applyDCdifferenceCoding();
    void doZigZagSorting();
    void doRLEandCategoryCoding();
    void doHuffmanEncoding(SymbolCodeMap& Y_DC, SymbolCodeMap& Y_AC, SymbolCodeMap& C_DC, SymbolCodeMap& C_AC);

    // The stage methods applySubsampling / applyDCT are host code by default (API compatibility; they are not on the encode
    // path).  stagesOnDevice(true) sends them -- every SubsamplingMode, every DCTMode -- through jpgenc_stage_subsample /
    // jpgenc_stage_dct instead (same doubles, bit for bit; throws std::runtime_error without a GPU).
    static void stagesOnDevice(bool on);

    // whole encode on the GPU; `file` receives a baseline JFIF file byte-identical to the reference's
    void writeJPEG(std::string file);
    // same, into memory (what writeJPEG writes)
    std::vector<Byte> encodeJPEG();

    uint width, height;                    // padded to multiples of 16
    uint real_width, real_height;          // as in the file
    uint subsample_width, subsample_height;
    matrix<PixelDataType>&R, &G, &B;
    matrix<PixelDataType>&Y, &Cb, &Cr;

private:
    friend Image loadPPM(std::string path);
    ColorSpace color_space_type;
    matrix<PixelDataType> one, two, three;
    matrix<PixelDataType> DctY, DctCb, DctCr;
    matrix<int> QY, QCb, QCr;
    matrix<std::vector<Category_Code>> CategoryCodeY, CategoryCodeCb, CategoryCodeCr;
    matrix<Bitstream> BitstreamY, BitstreamCb, BitstreamCr;
    // what the GPU path consumes: the file's raw 8-bit samples (un-scaled, un-padded) and its maxval
    std::vector<Byte> samples_;
    uint maxval_ = 255;
};
