// Host-side Huffman table construction for the encode path.
//
// Produces exactly what the reference's generateHuffmanCode(text) produces (src/Huffman.cpp:3-66,
// include/Huffman.hpp:114-174) but starts from the symbol histogram plus first-occurrence keys that the K2
// kernel delivers instead of the multi-million-entry symbol text.
//
// Why this is not "just a Huffman builder": the reference's code lengths AND the order of symbols inside a
// length depend on two library-defined orders:
//   * symbols enter package-merge in std::unordered_map<int,int> ITERATION order, where the map was filled in
//     order of first appearance in the text;
//   * ties between equal weights are broken by std::priority_queue's heap layout.
// Both are reproduced here by construction: the same libstdc++ containers/algorithms are driven through the
// same sequence of insertions, pushes and pops.  Packages are arena nodes (weight + two children) instead of
// the reference's merged symbol vectors; leaves are counted when the final level is drained.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/jpgenc_b200.h"

namespace {

struct HeapItem {
    int64_t weight;
    int node;
};
struct HeavierFirstOut {   // priority_queue "less": the lightest package surfaces (Huffman.hpp:117-119)
    bool operator()(const HeapItem& a, const HeapItem& b) const { return a.weight > b.weight; }
};
struct Node {
    int left, right, symbol;   // symbol >= 0 for a leaf
};

class Level {
public:
    void push(HeapItem it) {
        items_.push_back(it);
        std::push_heap(items_.begin(), items_.end(), HeavierFirstOut());
    }
    HeapItem pop() {
        HeapItem top = items_.front();
        std::pop_heap(items_.begin(), items_.end(), HeavierFirstOut());
        items_.pop_back();
        return top;
    }
    size_t size() const { return items_.size(); }
    void clear() { items_.clear(); }
private:
    std::vector<HeapItem> items_;
};

void collect_leaves(const std::vector<Node>& nodes, int root, std::vector<int>& out) {
    std::vector<int> stack{root};
    while (!stack.empty()) {
        const int k = stack.back();
        stack.pop_back();
        if (nodes[k].symbol >= 0) { out.push_back(nodes[k].symbol); continue; }
        stack.push_back(nodes[k].left);
        stack.push_back(nodes[k].right);
    }
}

// per_length[len] = symbols with that code length, in the order the DHT segment lists them
void assign_codes(const std::vector<std::vector<int>>& per_length, jpgenc_huff_table* t) {
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {                      // generateCodes, src/Huffman.cpp:50-66
        t->counts[len - 1] = static_cast<uint8_t>(per_length[len].size());
        for (int s : per_length[len]) {
            t->code_msb[s] = code << (32 - len);
            t->length[s] = static_cast<uint8_t>(len);
            t->symbols[k++] = static_cast<uint8_t>(s);
            ++code;
        }
        code <<= 1;
    }
}

}  // namespace

extern "C" int jpgenc_build_huffman(const uint32_t count[256], const uint64_t first_pos[256], jpgenc_huff_table* out) {
    if (!count || !first_pos || !out) return JPGENC_ERR_ARG;
    std::memset(out, 0, sizeof *out);

    // distinct symbols in order of first appearance == the order the reference's counting loop creates map entries
    std::vector<int> appearance;
    for (int s = 0; s < 256; ++s)
        if (count[s]) appearance.push_back(s);
    if (appearance.empty()) return JPGENC_ERR_ARG;             // reference asserts text.size() > 0
    std::sort(appearance.begin(), appearance.end(), [&](int a, int b) { return first_pos[a] < first_pos[b]; });
    std::unordered_map<int, int> freq;
    for (int s : appearance) freq[s] = static_cast<int>(count[s]);
    out->nsymbols = static_cast<int32_t>(freq.size());

    std::vector<std::vector<int>> per_length(18);
    if (freq.size() == 1) {                                    // src/Huffman.cpp:17-25: the lone symbol gets code "0"
        per_length[1].push_back(appearance[0]);
        assign_codes(per_length, out);
        return JPGENC_OK;
    }

    constexpr int kLimit = 15;                                 // package_merge(symbol_frequency, 15)
    std::vector<Node> nodes;
    nodes.reserve(freq.size() * (kLimit + 2));
    Level blueprint;
    for (const auto& kv : freq) {                              // map iteration order feeds the first heap
        nodes.push_back({-1, -1, kv.first});
        blueprint.push({static_cast<int64_t>(kv.second), static_cast<int>(nodes.size()) - 1});
    }
    Level current = blueprint;
    for (int lvl = 0; lvl < kLimit; ++lvl) {
        Level next;
        if (lvl + 1 < kLimit) next = blueprint;                // every level but the last starts as a copy of the leaves
        while (current.size() > 1) {
            const HeapItem a = current.pop();
            const HeapItem b = current.pop();
            nodes.push_back({a.node, b.node, -1});
            next.push({a.weight + b.weight, static_cast<int>(nodes.size()) - 1});
        }
        current = std::move(next);
    }
    // drain the last level: a symbol's code length is the number of times it occurs in the surviving packages
    std::unordered_map<int, int> code_lengths;
    std::vector<int> leaves;
    while (current.size()) {
        const HeapItem p = current.pop();
        leaves.clear();
        collect_leaves(nodes, p.node, leaves);
        std::sort(leaves.begin(), leaves.end());               // a package lists its symbols in ascending order
        for (int s : leaves) ++code_lengths[s];
    }
    for (const auto& kv : code_lengths) per_length[kv.second].push_back(kv.first);

    // preventOnlyOnesCode (src/Huffman.cpp:37-48): the last symbol of the deepest level moves one level down
    int deepest = 16;
    while (deepest > 0 && per_length[deepest].empty()) --deepest;
    const int moved = per_length[deepest].back();
    per_length[deepest].pop_back();
    per_length[deepest + 1].push_back(moved);

    assign_codes(per_length, out);
    return JPGENC_OK;
}
