// Bit container with the public interface of the reference's Bitstream_Generic (include/BitstreamGeneric.hpp:12-314):
// bits are appended MSB-first into BlockType words, `fill()` completes the open word with 1s, streaming out inserts
// a 0x00 after every 0xFF block.  Written for the B200 build: bits are kept in whole words and appended word-wise
// (the reference appends one bit at a time); on the GPU path the final scan never passes through this class at all
// (K3/K4 in csrc/entropy.cu produce it), it remains for table codes, tests and callers of the stage API.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <istream>
#include <ostream>
#include <vector>

template <typename BlockType>
class Bitstream_Generic {
public:
    static const std::size_t block_size = sizeof(BlockType) * 8;
    typedef std::vector<BlockType> ContainerType;

    // proxy for one bit (operator[])
    class BitView {
        friend class Bitstream_Generic<BlockType>;
        BitView(ContainerType& words, std::size_t word, unsigned bit) : words_(words), word_(word), bit_(bit) {}
        ContainerType& words_;
        const std::size_t word_;
        const unsigned bit_;
    public:
        operator bool() const { return ((words_[word_] >> bit_) & BlockType(1)) != 0; }
        void operator=(bool v) {
            const BlockType m = BlockType(1) << bit_;
            words_[word_] = v ? BlockType(words_[word_] | m) : BlockType(words_[word_] & ~m);
        }
    };

    Bitstream_Generic() : nbits_(0) {}
    Bitstream_Generic(std::initializer_list<bool> bits) : nbits_(0) { *this << bits; }
    // the low `number_of_bits` bits of data, most significant of them first (reference ctor :119-124)
    Bitstream_Generic(uint32_t data, int number_of_bits) : nbits_(0) { push_back_LSB_mode(data, number_of_bits); }

    Bitstream_Generic& operator<<(bool bit) { append(bit ? 1u : 0u, 1); return *this; }
    Bitstream_Generic& operator<<(std::initializer_list<bool> bits) {
        for (bool b : bits) append(b ? 1u : 0u, 1);
        return *this;
    }
    Bitstream_Generic& operator<<(Bitstream_Generic<BlockType>& other) {
        // whole words of `other`, then its open tail
        const std::size_t full = other.nbits_ / block_size, tail = other.nbits_ % block_size;
        for (std::size_t i = 0; i < full; ++i) append64(other.words_[i], block_size);
        if (tail) append64(other.words_[full] >> (block_size - tail), tail);
        return *this;
    }
    Bitstream_Generic& push_back(bool bit) { return *this << bit; }
    // top `number_of_bits` bits of the 32-bit word (reference :182-195)
    Bitstream_Generic& push_back(uint32_t data, int number_of_bits) {
        if (number_of_bits > 0) append(number_of_bits >= 32 ? data : data >> (32 - number_of_bits), number_of_bits);
        return *this;
    }
    // low `number_of_bits` bits (reference :197-210)
    Bitstream_Generic& push_back_LSB_mode(uint32_t data, int number_of_bits) {
        if (number_of_bits > 0) append(data, number_of_bits);
        return *this;
    }

    template <typename B>
    friend std::ostream& operator<<(std::ostream& out, const Bitstream_Generic<B>& bs);
    template <typename B>
    friend std::istream& operator>>(std::istream& in, Bitstream_Generic<B>& bs);
    template <typename B>
    friend bool operator==(const Bitstream_Generic<B>& a, const Bitstream_Generic<B>& b);

    BitView operator[](unsigned int pos) {
        assert(pos < nbits_);
        return BitView(words_, pos / block_size, static_cast<unsigned>(block_size - 1 - pos % block_size));
    }

    // `number_of_bits` bits starting at from_position, returned left-aligned in T (reference :264-305)
    template <typename T>
    T extractT(uint8_t number_of_bits, std::size_t from_position) {
        assert(number_of_bits <= sizeof(T) * 8);
        assert(from_position + number_of_bits - 1 < nbits_);
        T acc = 0;
        for (unsigned i = 0; i < number_of_bits; ++i) {
            const std::size_t p = from_position + i;
            const BlockType bit = (words_[p / block_size] >> (block_size - 1 - p % block_size)) & BlockType(1);
            acc = static_cast<T>((acc << 1) | static_cast<T>(bit));
        }
        return number_of_bits ? static_cast<T>(acc << (sizeof(T) * 8 - number_of_bits)) : T(0);
    }
    uint32_t extract(uint8_t number_of_bits, std::size_t from_position) { return extractT<uint32_t>(number_of_bits, from_position); }

    unsigned int size() const { return static_cast<unsigned int>(nbits_); }

    // complete the open word with 1s; on an EMPTY stream the reference emits one whole word of 1s (:242-248 with the
    // constructor's bit_idx == block_size), which is reproduced here
    void fill() {
        const std::size_t used = nbits_ % block_size;
        if (nbits_ == 0) append64(~uint64_t(0) >> (64 - block_size), block_size);
        else if (used) append64(~uint64_t(0) >> (64 - (block_size - used)), block_size - used);
    }

private:
    void append(uint32_t value, int n) { append64(value, static_cast<std::size_t>(n)); }
    // low n bits of value (n <= 64), MSB of them first
    void append64(uint64_t value, std::size_t n) {
        while (n) {
            const std::size_t used = nbits_ % block_size;
            if (used == 0) words_.push_back(0);
            const std::size_t room = block_size - used, take = n < room ? n : room;
            const uint64_t piece = (value >> (n - take)) & (take == 64 ? ~uint64_t(0) : ((uint64_t(1) << take) - 1));
            words_.back() = static_cast<BlockType>(words_.back() | static_cast<BlockType>(piece << (room - take)));
            nbits_ += take;
            n -= take;
        }
    }
    ContainerType words_;
    std::size_t nbits_;
};

template <typename BlockType>
std::ostream& operator<<(std::ostream& out, const Bitstream_Generic<BlockType>& bs) {
    for (const BlockType& w : bs.words_) {
        out.write(reinterpret_cast<const char*>(&w), sizeof(w));
        if (w == 0xFF) out.put(0x00);                 // JPEG byte stuffing (reference :213-224)
    }
    return out;
}

template <typename BlockType>
std::istream& operator>>(std::istream& in, Bitstream_Generic<BlockType>& bs) {
    BlockType w = 0;
    while (in.read(reinterpret_cast<char*>(&w), sizeof(w))) {   // a trailing partial word is dropped, as in the reference
        bs.words_.push_back(w);
        bs.nbits_ += Bitstream_Generic<BlockType>::block_size;
    }
    return in;
}

template <typename BlockType>
bool operator==(const Bitstream_Generic<BlockType>& a, const Bitstream_Generic<BlockType>& b) {
    return a.nbits_ == b.nbits_ && a.words_ == b.words_;
}

using Bitstream8 = Bitstream_Generic<uint8_t>;
using Bitstream16 = Bitstream_Generic<uint16_t>;
using Bitstream32 = Bitstream_Generic<uint32_t>;
using Bitstream64 = Bitstream_Generic<uint64_t>;
using Bitstream = Bitstream8;
typedef std::initializer_list<bool> Bits;
