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

// ---------------------------------------------------------------------------------------------------------------------
// The build the encode path calls.  Same containers for the two library-defined orders (std::unordered_map for the
// symbol and code-length maps, std::push_heap / std::pop_heap for the package queues), but no allocation per level:
// a queue item is ONE 64-bit word (weight << 13 | arena index, compared on the weight alone, so the heap algorithms
// make exactly the moves they make on the reference's items; pop_heap is written out, see PackedLevel::pop), the queues are fixed arrays and a level starts as a
// memcpy of the heapified leaves.  For the 28-symbol luma AC table of a 3840x2160 frame this is the host time between
// K2 and K3 of a single image -- 21 us with one vector per level, see DESIGN.md 3.
namespace {

constexpr int kNodeBits = 13;                                  // 256 leaves + 15 levels of < 256 packages < 8192 arena nodes
constexpr uint64_t kNodeMask = (1ull << kNodeBits) - 1;
struct PackedHeavierFirstOut {
    bool operator()(uint64_t a, uint64_t b) const { return (a >> kNodeBits) > (b >> kNodeBits); }
};
struct PackedNode { int16_t left, right, symbol; };
struct SymbolSet { uint64_t w[4]; };

struct PackedLevel {
    uint64_t v[2 * 256 + 8];
    int n = 0;
    void push(uint64_t weight, int node) {
        v[n++] = weight << kNodeBits | static_cast<uint64_t>(node);
        std::push_heap(v, v + n, PackedHeavierFirstOut());
    }
    // top() + std::pop_heap + pop_back.  pop_heap is libstdc++'s __adjust_heap written out (the hole at the root walks down
    // to the bottom, always to the lighter child, the right one on a tie; the former last item is then sifted up from
    // there) with the child choice as arithmetic instead of a branch: the choice is a coin flip for the predictor.
    uint64_t pop() {
        const uint64_t top = v[0];
        const int len = n - 1;
        if (len > 0) {
            const uint64_t val = v[len];
            int hole = 0, child = 0;
            const int last_parent = (len - 1) / 2;
            while (child < last_parent) {
                child = 2 * (child + 1);
                child -= (v[child] >> kNodeBits) > (v[child - 1] >> kNodeBits);
                v[hole] = v[child];
                hole = child;
            }
            if ((len & 1) == 0 && child == (len - 2) / 2) {
                child = 2 * (child + 1);
                v[hole] = v[child - 1];
                hole = child - 1;
            }
            int parent = (hole - 1) / 2;
            while (hole > 0 && (v[parent] >> kNodeBits) > (val >> kNodeBits)) {
                v[hole] = v[parent];
                hole = parent;
                parent = (hole - 1) / 2;
            }
            v[hole] = val;
        }
        n = len;
        return top;
    }
    void start_as(const PackedLevel& o) { std::memcpy(v, o.v, sizeof(uint64_t) * o.n); n = o.n; }
};

}  // namespace

static int build_huffman_packed(const uint32_t count[256], const uint64_t first_pos[256], jpgenc_huff_table* out);

// (no exception crosses the C boundary: the maps allocate)
extern "C" int jpgenc_build_huffman(const uint32_t count[256], const uint64_t first_pos[256], jpgenc_huff_table* out) try {
    if (!count || !first_pos || !out) return JPGENC_ERR_ARG;
    for (int s = 0; s < 256; ++s)
        if (count[s] >> 31) return jpgenc_build_huffman_containers(count, first_pos, out);   // int overflow of the reference's counter: its arithmetic, not ours
    return build_huffman_packed(count, first_pos, out);
} catch (...) {
    return JPGENC_ERR_NOMEM;
}

static int build_huffman_packed(const uint32_t count[256], const uint64_t first_pos[256], jpgenc_huff_table* out) {
    std::memset(out, 0, sizeof *out);
    int appearance[256], n = 0;
    for (int s = 0; s < 256; ++s)
        if (count[s]) appearance[n++] = s;
    if (n == 0) return JPGENC_ERR_ARG;
    std::sort(appearance, appearance + n, [&](int a, int b) { return first_pos[a] < first_pos[b]; });
    std::unordered_map<int, int> freq;
    for (int i = 0; i < n; ++i) freq[appearance[i]] = static_cast<int>(count[appearance[i]]);
    out->nsymbols = n;

    uint8_t per_length[18][256];
    int per_length_n[18] = {};
    auto emit = [&] {                                          // generateCodes, src/Huffman.cpp:50-66
        uint32_t code = 0;
        int k = 0;
        for (int len = 1; len <= 16; ++len) {
            out->counts[len - 1] = static_cast<uint8_t>(per_length_n[len]);
            for (int i = 0; i < per_length_n[len]; ++i) {
                const int s = per_length[len][i];
                out->code_msb[s] = code << (32 - len);
                out->length[s] = static_cast<uint8_t>(len);
                out->symbols[k++] = static_cast<uint8_t>(s);
                ++code;
            }
            code <<= 1;
        }
    };
    if (n == 1) {
        per_length[1][per_length_n[1]++] = static_cast<uint8_t>(appearance[0]);
        emit();
        return JPGENC_OK;
    }

    constexpr int kLimit = 15;
    constexpr int kArena = 256 * (kLimit + 1) + 8;
    PackedNode nodes[kArena];
    SymbolSet sets[kArena];                                    // the symbols among a node's leaves
    int nn = 0;
    PackedLevel blueprint, a, b;
    for (const auto& kv : freq) {
        nodes[nn] = {-1, -1, static_cast<int16_t>(kv.first)};
        sets[nn] = SymbolSet{};
        sets[nn].w[kv.first >> 6] = 1ull << (kv.first & 63);
        blueprint.push(static_cast<uint64_t>(kv.second), nn++);
    }
    PackedLevel *current = &a, *next = &b;
    current->start_as(blueprint);
    for (int lvl = 0; lvl < kLimit; ++lvl) {
        if (lvl + 1 < kLimit) next->start_as(blueprint); else next->n = 0;
        while (current->n > 1) {
            const uint64_t x = current->pop();
            const uint64_t y = current->pop();
            const int xn = static_cast<int>(x & kNodeMask), yn = static_cast<int>(y & kNodeMask);
            nodes[nn] = {static_cast<int16_t>(xn), static_cast<int16_t>(yn), -1};
            for (int w = 0; w < 4; ++w) sets[nn].w[w] = sets[xn].w[w] | sets[yn].w[w];
            next->push((x >> kNodeBits) + (y >> kNodeBits), nn++);
        }
        std::swap(current, next);
    }
    // Drain the last level.  The reference walks every surviving package, sorts its symbols and bumps a map entry per
    // occurrence; what that fixes is (a) the ORDER in which symbols enter the lengths map -- package by package in pop
    // order, ascending inside a package: read off the packages' symbol sets -- and (b) how often each symbol occurs in
    // all of them: one pass down the arena, a package handing its multiplicity to its two children (children always
    // have lower indices than their package).
    std::unordered_map<int, int> code_lengths;
    uint16_t times[kArena];
    std::memset(times, 0, sizeof(uint16_t) * nn);
    uint64_t entered[4] = {0, 0, 0, 0};
    while (current->n) {
        const int root = static_cast<int>(current->pop() & kNodeMask);
        ++times[root];
        for (int w = 0; w < 4; ++w) {
            uint64_t fresh = sets[root].w[w] & ~entered[w];
            entered[w] |= fresh;
            for (; fresh; fresh &= fresh - 1) code_lengths.emplace(w * 64 + __builtin_ctzll(fresh), 0);
        }
    }
    for (int k = nn - 1; k >= n; --k) {
        times[nodes[k].left] += times[k];
        times[nodes[k].right] += times[k];
    }
    uint16_t length_of[256];
    for (int k = 0; k < n; ++k) length_of[nodes[k].symbol] = times[k];
    for (const auto& kv : code_lengths) {
        const int len = length_of[kv.first];
        per_length[len][per_length_n[len]++] = static_cast<uint8_t>(kv.first);
    }

    int deepest = 16;                                          // preventOnlyOnesCode, src/Huffman.cpp:37-48
    while (deepest > 0 && per_length_n[deepest] == 0) --deepest;
    const int moved = per_length[deepest][--per_length_n[deepest]];
    per_length[deepest + 1][per_length_n[deepest + 1]++] = static_cast<uint8_t>(moved);
    emit();
    return JPGENC_OK;
}

// The straightforward statement: one container per thing the reference has one for.  Kept as the anchor the other
// builds (packed above, array restatement on the device) are tested against; itself pinned against the compiled
// reference in tests/test_oracle_vs_reference.py.
extern "C" int jpgenc_build_huffman_containers(const uint32_t count[256], const uint64_t first_pos[256], jpgenc_huff_table* out) try {
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
} catch (...) {
    return JPGENC_ERR_NOMEM;
}
