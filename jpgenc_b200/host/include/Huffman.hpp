// Huffman table construction / coding with the reference's public interface (include/Huffman.hpp:21-174,
// src/Huffman.cpp).  The table build is the B200 build's own (host/huffman_build.cpp): it starts from a symbol
// histogram + first-appearance order -- what the GPU statistics kernel delivers -- and reproduces the reference's
// tables exactly (including the order of symbols inside a code length, which depends on libstdc++'s unordered_map and
// heap orders).  generateHuffmanCode(text) is a thin adaptor over it for callers that still hold a symbol text.
#pragma once
#include <cstdint>
#include <iterator>
#include <unordered_map>
#include <utility>
#include <vector>

#include "BitstreamGeneric.hpp"

// names the reference's header exports into the global namespace (Huffman.hpp:15-18); callers rely on them
using std::pair;
using std::unordered_map;
using std::vector;

struct Code {
    using CodeType = uint32_t;
    static const std::size_t max_code_length = sizeof(CodeType) * 8;
    Code() : code(0), length(0) {}
    // right-aligned `code` of `length` bits -> stored MSB-aligned (reference Huffman.hpp:26-34)
    Code(CodeType code_, uint8_t length_) : code(length_ ? code_ << (max_code_length - length_) : 0), length(length_) {}
    explicit Code(Bitstream bits) : length(static_cast<uint8_t>(bits.size())) { code = length ? bits.extract(length, 0) : 0; }
    CodeType code;      // first code bit is the MSB
    uint8_t length;
};

using SymbolCodeMap = std::unordered_map<int, Code>;
using SymbolsPerLength = std::vector<std::vector<int>>;   // [len] = symbols with that code length, DHT order

struct Symbol {
    int symbol, frequency;
    Symbol() : symbol(0), frequency(0) {}
    Symbol(int s, int f) : symbol(s), frequency(f) {}
};
inline bool operator<(const Symbol& a, const Symbol& b) { return a.symbol < b.symbol; }

// one node of a package-merge level: a weight and the (sorted) symbols it covers
struct Package {
    int weight;
    std::vector<Symbol> symbols;
    Package(const Symbol& s) : weight(s.frequency), symbols(1, s) {}
    Package(const Package& a, const Package& b);
};

// text -> (symbol -> code, symbols grouped by code length); table lengths limited to 15 (+1 for the "no all-ones
// code" fix-up).  Reference: src/Huffman.cpp:3-35.
std::pair<SymbolCodeMap, SymbolsPerLength> generateHuffmanCode(std::vector<int> text);
// length-limited code lengths by package-merge; result has length_limit+2 entries (reference Huffman.hpp:114-174)
SymbolsPerLength package_merge(std::vector<Symbol> symbols, int length_limit);
void preventOnlyOnesCode(SymbolsPerLength& symbols);
SymbolCodeMap generateCodes(const SymbolsPerLength& symbols);

Bitstream huffmanEncode(std::vector<int> text, SymbolCodeMap code_map);
std::vector<int> huffmanDecode(Bitstream bitstream, SymbolCodeMap code_map);

struct DecodeEntry {
    DecodeEntry(uint32_t code, uint8_t code_length, int symbol);
    uint32_t code;          // MSB-aligned, unused low bits set to 1
    uint8_t code_length;
    int symbol;
    bool operator<(const DecodeEntry& o) const { return code < o.code; }
};
