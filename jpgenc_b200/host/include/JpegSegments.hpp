// JFIF marker segments with the reference's public interface (include/JpegSegments.hpp:10-379): sSOI, sAPP0, sDQT,
// sSOF0, sDHT, sSOS, sEOI with the same setters and operator<<.  The reference streams packed structs; here every
// segment serialises itself field by field into a byte vector (no reliance on struct layout), and Image::writeJPEG
// itself does not go through these classes at all: it lays the header down with jpgenc_write_headers directly in
// front of the scan that comes back from the GPU.  sizeof() of the fixed-size segments is preserved because the
// reference's tests check it (ImageTest.cpp:302,319).
#pragma once
#include <array>
#include <cassert>
#include <cstdint>
#include <initializer_list>
#include <ostream>
#include <vector>

#include "Image.hpp"

namespace Segment {
using std::vector;

template <std::size_t N>
using Bytes = std::array<Byte, N>;

template <std::size_t N>
void set(Bytes<N>& dst, std::initializer_list<Byte> src) {
    assert(src.size() <= N);
    std::size_t i = 0;
    for (Byte b : src) dst[i++] = b;
}
inline Byte getHi(short v) { return static_cast<Byte>((v >> 8) & 0xFF); }
inline Byte getLo(short v) { return static_cast<Byte>(v & 0xFF); }

namespace ComponentSetup {
enum ID : Byte { Y = 1, Cb, Cr };
enum Subsampling : Byte { NoSubSampling = 0x22, Half = 0x11 };
enum QuantizationTableID { Zero = 0, One, Two, Three };
}  // namespace ComponentSetup

namespace detail {
inline void put(std::ostream& out, const Byte* p, std::size_t n) { out.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n)); }
template <std::size_t N>
void put(std::ostream& out, const Bytes<N>& b) { put(out, b.data(), N); }
}  // namespace detail

struct sSOI {
    const Bytes<2> marker{{0xFF, 0xD8}};
    friend std::ostream& operator<<(std::ostream& out, const sSOI& s) { detail::put(out, s.marker); return out; }
};

struct sEOI {
    const Bytes<2> marker{{0xFF, 0xD9}};
    friend std::ostream& operator<<(std::ostream& out, const sEOI& s) { detail::put(out, s.marker); return out; }
};

struct sAPP0 {
    const Bytes<2> marker{{0xFF, 0xE0}};
    Bytes<2> len{{0, 16}};
    const Bytes<5> type{{'J', 'F', 'I', 'F', 0}};
    const Bytes<2> rev{{1, 1}};
    const Bytes<1> pixelsize{{0}};
    Bytes<2> x_density{{0, 1}};
    Bytes<2> y_density{{0, 1}};
    const Bytes<2> thumbnail_size{{0, 0}};
    sAPP0& setLen(short v) { set(len, {getHi(v), getLo(v)}); return *this; }
    sAPP0& setXdensity(short v) { set(x_density, {getHi(v), getLo(v)}); return *this; }
    sAPP0& setYdensity(short v) { set(y_density, {getHi(v), getLo(v)}); return *this; }
    friend std::ostream& operator<<(std::ostream& out, const sAPP0& s) {
        detail::put(out, s.marker); detail::put(out, s.len); detail::put(out, s.type); detail::put(out, s.rev);
        detail::put(out, s.pixelsize); detail::put(out, s.x_density); detail::put(out, s.y_density);
        detail::put(out, s.thumbnail_size);
        return out;
    }
};

struct sSOF0 {
    static const Byte num_components = 3;
    const Bytes<2> marker{{0xFF, 0xC0}};
    const Bytes<2> len{{0, 8 + num_components * 3}};
    const Bytes<1> precision{{8}};
    Bytes<2> image_size_y{{0, 0}};
    Bytes<2> image_size_x{{0, 0}};
    const Bytes<1> component_count{{num_components}};
    Bytes<num_components * 3> component_setup{{ComponentSetup::Y, ComponentSetup::NoSubSampling, 0,
                                               ComponentSetup::Cb, ComponentSetup::Half, 1,
                                               ComponentSetup::Cr, ComponentSetup::Half, 2}};
    sSOF0() {}
    sSOF0(int size_x, int size_y) {
        setImageSizeX(static_cast<short>(size_x));
        setImageSizeY(static_cast<short>(size_y));
    }
    sSOF0(int size_x, int size_y, std::initializer_list<Byte> comp_setup) : sSOF0(size_x, size_y) { set(component_setup, comp_setup); }
    sSOF0& setImageSizeX(short v) { set(image_size_x, {getHi(v), getLo(v)}); return *this; }
    sSOF0& setImageSizeY(short v) { set(image_size_y, {getHi(v), getLo(v)}); return *this; }
    sSOF0& setupY(ComponentSetup::Subsampling s, ComponentSetup::QuantizationTableID q) { component_setup[1] = s; component_setup[2] = static_cast<Byte>(q); return *this; }
    sSOF0& setupCb(ComponentSetup::Subsampling s, ComponentSetup::QuantizationTableID q) { component_setup[4] = s; component_setup[5] = static_cast<Byte>(q); return *this; }
    sSOF0& setupCr(ComponentSetup::Subsampling s, ComponentSetup::QuantizationTableID q) { component_setup[7] = s; component_setup[8] = static_cast<Byte>(q); return *this; }
    sSOF0& setComponentSetup(std::initializer_list<Byte> c) { set(component_setup, c); return *this; }
    friend std::ostream& operator<<(std::ostream& out, const sSOF0& s) {
        detail::put(out, s.marker); detail::put(out, s.len); detail::put(out, s.precision); detail::put(out, s.image_size_y);
        detail::put(out, s.image_size_x); detail::put(out, s.component_count); detail::put(out, s.component_setup);
        return out;
    }
};

struct sDHT {
    const Bytes<2> marker{{0xFF, 0xC4}};
    Bytes<2> len{{0, 0}};
    struct sHT {
        Bytes<1> HT_info;
        Bytes<16> code_lengths;
        std::vector<Byte> symbols;
    };
    std::vector<sHT> HTs;
    enum Class : Byte { DC = 0, AC };
    enum Destination : Byte { First = 0, Second };

    // codelength_symbols[n] = symbols whose code has n bits (17 entries, entry 0 unused)
    sDHT& pushCodeData(vector<vector<int>>& codelength_symbols, Class cls, Destination dest) {
        assert(codelength_symbols.size() == 17);
        sHT ht;
        ht.HT_info[0] = static_cast<Byte>((cls << 4) | dest);
        for (std::size_t n = 1; n < codelength_symbols.size(); ++n) {
            assert(codelength_symbols[n].size() < 256);
            ht.code_lengths[n - 1] = static_cast<Byte>(codelength_symbols[n].size());
            for (int s : codelength_symbols[n]) ht.symbols.push_back(static_cast<Byte>(s));
        }
        HTs.push_back(ht);
        return recalc();
    }
    sDHT& clear() { HTs.clear(); return recalc(); }
    friend std::ostream& operator<<(std::ostream& out, const sDHT& s) {
        detail::put(out, s.marker); detail::put(out, s.len);
        for (const sHT& ht : s.HTs) {
            detail::put(out, ht.HT_info); detail::put(out, ht.code_lengths);
            detail::put(out, ht.symbols.data(), ht.symbols.size());
        }
        return out;
    }
private:
    sDHT& recalc() {
        int n = 2;
        for (const sHT& ht : HTs) n += 17 + static_cast<int>(ht.symbols.size());
        assert(n < 65536);
        set(len, {getHi(static_cast<short>(n)), getLo(static_cast<short>(n))});
        return *this;
    }
};

struct sDQT {
    const Bytes<2> marker{{0xFF, 0xDB}};
    Bytes<2> len{{0, 0}};
    struct sQT {
        Bytes<1> QT_info;
        std::array<Byte, 64> coefficients;
    };
    std::vector<sQT> QTs;
    // coefficients already in zigzag order
    sDQT& pushQuantizationTable(vector<Byte>& coefficients, ComponentSetup::QuantizationTableID dest) {
        assert(coefficients.size() == 64);
        sQT qt;
        qt.QT_info[0] = static_cast<Byte>(dest);
        for (std::size_t i = 0; i < 64; ++i) qt.coefficients[i] = coefficients[i];
        QTs.push_back(qt);
        return recalc();
    }
    sDQT& clear() { QTs.clear(); return recalc(); }
    friend std::ostream& operator<<(std::ostream& out, const sDQT& s) {
        detail::put(out, s.marker); detail::put(out, s.len);
        for (const sQT& qt : s.QTs) { detail::put(out, qt.QT_info); detail::put(out, qt.coefficients.data(), 64); }
        return out;
    }
private:
    sDQT& recalc() {
        const std::size_t n = 2 + QTs.size() * 65;
        assert(n < 65536);
        set(len, {getHi(static_cast<short>(n)), getLo(static_cast<short>(n))});
        return *this;
    }
};

struct sSOS {
    const Bytes<2> marker{{0xFF, 0xDA}};
    Bytes<2> len{{0, 6 + 2 * 3}};
    Bytes<1> num_components{{3}};
    Bytes<6> component_setup{{ComponentSetup::Y, 0x00, ComponentSetup::Cb, 0x00, ComponentSetup::Cr, 0x00}};
    Bytes<3> blubb{{0x00, 0x3F, 0x00}};
    sSOS& setupY(sDHT::Destination dc, sDHT::Destination ac) { component_setup[1] = static_cast<Byte>((dc << 4) | ac); return *this; }
    sSOS& setupCb(sDHT::Destination dc, sDHT::Destination ac) { component_setup[3] = static_cast<Byte>((dc << 4) | ac); return *this; }
    sSOS& setupCr(sDHT::Destination dc, sDHT::Destination ac) { component_setup[5] = static_cast<Byte>((dc << 4) | ac); return *this; }
    friend std::ostream& operator<<(std::ostream& out, const sSOS& s) {
        detail::put(out, s.marker); detail::put(out, s.len); detail::put(out, s.num_components);
        detail::put(out, s.component_setup); detail::put(out, s.blubb);
        return out;
    }
};

}  // namespace Segment
