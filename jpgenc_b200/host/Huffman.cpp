#include "include/Huffman.hpp"

#include <algorithm>
#include <cassert>
#include <cstring>

#include "../../include/jpgenc_b200.h"

Package::Package(const Package& a, const Package& b) : weight(a.weight + b.weight) {
    symbols.reserve(a.symbols.size() + b.symbols.size());
    std::merge(a.symbols.begin(), a.symbols.end(), b.symbols.begin(), b.symbols.end(), std::back_inserter(symbols));
}

namespace {

// heap item of one package-merge level: only the weight takes part in ordering
struct Entry {
    long long weight;
    int node;
};
struct LightestOnTop {
    bool operator()(const Entry& a, const Entry& b) const { return a.weight > b.weight; }
};
struct Node {
    int left, right, symbol;
};

void gather(const std::vector<Node>& nodes, int root, std::vector<int>& out) {
    std::vector<int> todo(1, root);
    while (!todo.empty()) {
        const Node n = nodes[todo.back()];
        todo.pop_back();
        if (n.symbol >= 0) out.push_back(n.symbol);
        else { todo.push_back(n.left); todo.push_back(n.right); }
    }
}

}  // namespace

SymbolsPerLength package_merge(std::vector<Symbol> symbols, int length_limit) {
    std::vector<Node> nodes;
    std::vector<Entry> leaves;                      // the heap every level starts from, built in input order
    for (const Symbol& s : symbols) {
        nodes.push_back({-1, -1, s.symbol});
        leaves.push_back({s.frequency, static_cast<int>(nodes.size()) - 1});
        std::push_heap(leaves.begin(), leaves.end(), LightestOnTop());
    }
    std::vector<Entry> level = leaves;
    for (int depth = 0; depth < length_limit; ++depth) {
        std::vector<Entry> above;
        if (depth + 1 < length_limit) above = leaves;
        while (level.size() > 1) {
            Entry pair[2];
            for (Entry& e : pair) {
                e = level.front();
                std::pop_heap(level.begin(), level.end(), LightestOnTop());
                level.pop_back();
            }
            nodes.push_back({pair[0].node, pair[1].node, -1});
            above.push_back({pair[0].weight + pair[1].weight, static_cast<int>(nodes.size()) - 1});
            std::push_heap(above.begin(), above.end(), LightestOnTop());
        }
        level.swap(above);
    }
    std::unordered_map<int, int> length_of;         // filled in the order symbols come out of the last level
    std::vector<int> members;
    while (!level.empty()) {
        const Entry top = level.front();
        std::pop_heap(level.begin(), level.end(), LightestOnTop());
        level.pop_back();
        members.clear();
        gather(nodes, top.node, members);
        std::sort(members.begin(), members.end());
        for (int s : members) ++length_of[s];
    }
    SymbolsPerLength by_length(length_limit + 2);
    for (const auto& kv : length_of) by_length[kv.second].push_back(kv.first);
    return by_length;
}

void preventOnlyOnesCode(SymbolsPerLength& symbols) {
    assert(symbols.back().empty());
    std::size_t deepest = symbols.size() - 1;
    while (deepest > 0 && symbols[deepest].empty()) --deepest;
    const int moved = symbols[deepest].back();
    symbols[deepest].pop_back();
    symbols[deepest + 1].push_back(moved);
}

SymbolCodeMap generateCodes(const SymbolsPerLength& symbols) {
    SymbolCodeMap map;
    uint32_t next = 0;
    for (std::size_t len = 1; len < symbols.size(); ++len) {
        for (int s : symbols[len]) map[s] = Code(next++, static_cast<uint8_t>(len));
        next <<= 1;
    }
    return map;
}

std::pair<SymbolCodeMap, SymbolsPerLength> generateHuffmanCode(std::vector<int> text) {
    assert(!text.empty());
    // Any int may be a symbol here (the JPEG path only produces 0..255 and goes through jpgenc_build_huffman, which
    // this function agrees with).  Symbols enter package-merge in the iteration order of a map filled in order of
    // first appearance -- the library-defined order the reference's tables depend on.
    std::unordered_map<int, int> frequency;
    for (int s : text) ++frequency[s];
    if (frequency.size() == 1) {                      // a lone symbol gets the one-bit code "0"
        SymbolsPerLength per_length(17);
        per_length[1].push_back(text[0]);
        SymbolCodeMap map;
        map[text[0]] = Code(0, 1);
        return std::make_pair(map, per_length);
    }
    std::vector<Symbol> symbols;
    symbols.reserve(frequency.size());
    for (const auto& kv : frequency) symbols.emplace_back(kv.first, kv.second);
    SymbolsPerLength per_length = package_merge(symbols, 15);
    preventOnlyOnesCode(per_length);
    return std::make_pair(generateCodes(per_length), per_length);
}

Bitstream huffmanEncode(std::vector<int> text, SymbolCodeMap code_map) {
    Bitstream out;
    for (int s : text) {
        const Code& c = code_map[s];
        out.push_back(c.code, c.length);
    }
    return out;
}

DecodeEntry::DecodeEntry(uint32_t code_, uint8_t len, int sym)
    : code(len >= 32 ? code_ : code_ | ((1u << (32 - len)) - 1)), code_length(len), symbol(sym) {}

std::vector<int> huffmanDecode(Bitstream bitstream, SymbolCodeMap code_map) {
    // canonical prefix decode: extend the current code bit by bit until it names a symbol
    std::vector<DecodeEntry> table;
    for (const auto& kv : code_map) table.emplace_back(kv.second.code, kv.second.length, kv.first);
    std::sort(table.begin(), table.end());
    std::vector<int> text;
    unsigned pos = 0;
    const unsigned n = bitstream.size();
    while (pos < n) {
        bool hit = false;
        for (const DecodeEntry& e : table) {
            if (pos + e.code_length > n) continue;
            const uint32_t got = bitstream.extract(e.code_length, pos);
            const uint32_t want = e.code_length >= 32 ? e.code : e.code & ~((1u << (32 - e.code_length)) - 1);
            if (got == want) {
                text.push_back(e.symbol);
                pos += e.code_length;
                hit = true;
                break;
            }
        }
        if (!hit) break;            // trailing padding bits
    }
    return text;
}
