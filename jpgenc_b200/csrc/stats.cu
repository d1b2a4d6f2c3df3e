// K2 — DC differencing + run-length/category symbol generation.  One pass over the coefficients produces
//   * the 4 x 256 symbol histogram and, per symbol, the position of its first occurrence in the reference's symbol
//     text (what the Huffman table build needs), and
//   * the symbol stream itself, as flat 32-bit items in scan order (blockwalk.cuh), which is all K3 reads.
//
// Replaces Image::applyDCdifferenceCoding, Image::doRLEandCategoryCoding and the symbol-text collection
// (reference src/Image.cpp:638-735, 888-906; include/Coding.hpp:148-283).  The first-occurrence key is needed because
// generateHuffmanCode (src/Huffman.cpp:6-14) feeds package_merge in unordered_map iteration order, which depends on the
// order symbols first appear.  Text order (src/Image.cpp:892-906): Y blocks raster over the block grid,
// chroma = all Cb blocks then all Cr blocks; DC and AC symbols go to separate texts.
//
// Work distribution.  Blocks hold very different numbers of non-zero coefficients, so "one thread walks one block"
// diverges badly.  Instead a tile of 384 blocks is flattened: every thread sizes its block with a popcount of the
// block's 64-bit non-zero mask, a CTA scan gives each block its first item index, the threads expand their masks into
// 16-bit (block, position) descriptors, and then consecutive lanes process consecutive ITEMS: the run length of an
// item comes from the mask with one CLZ (distance to the previous set bit), so no lane ever loops over zeros and the
// item stores are coalesced.
#include "blockwalk.cuh"

namespace jpgenc {

constexpr int kDescCap = 2048;                           // descriptors per round (typical tiles need ~1300)

struct StatsParams {
    const int16_t* coef;
    uint32_t nblocks;
    uint32_t mcu_w;
    uint32_t n_mcu;
    uint32_t* g_hist;                 // [4][256]
    unsigned long long* g_first;      // [4][256]
    uint32_t* items;                  // flat item stream
    unsigned long long* cursor;       // next free item slot (tiles claim ranges in completion order)
    unsigned long long* tile_off;     // [tiles] first item of the tile
    uint32_t* tile_cnt;               // [tiles] items of the tile
};

// One histogram update for the whole warp: lanes with the same bin are counted by their lowest lane, so the hot
// symbols (EOB, +-1 coefficients, the common DC category) cost one shared-memory atomic instead of up to 32.
__device__ __forceinline__ void warp_count(uint32_t* s_hist, unsigned long long* s_first, bool has, int idx,
                                           unsigned long long key, uint32_t weight = 1) {
    const int lane = threadIdx.x & 31;
    const unsigned grp = __match_any_sync(0xffffffffu, has ? idx : (0x10000 | lane));
    if (has) {
        if (weight == 1) {
            if (lane == __ffs(grp) - 1) atomicAdd(&s_hist[idx], static_cast<uint32_t>(__popc(grp)));
        } else {
            atomicAdd(&s_hist[idx], weight);
        }
        if (key < s_first[idx]) atomicMin(&s_first[idx], key);
    }
}

constexpr int kStatsSmem = kTileSmemBytes                // tile + masks + dc
                           + kTileBlocks * 2             // dc difference
                           + kTileBlocks * 8             // text key of the block
                           + kDescCap * 2                // descriptors
                           + 4096 + 8192                 // histogram, first-occurrence keys
                           + 36 * 4;                     // scan scratch

__global__ void __launch_bounds__(kTileBlocks) symbol_stats_kernel(const __grid_constant__ StatsParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const TileView tv = tile_view(smem);
    uint8_t* at = smem + kTileSmemBytes;
    int16_t* s_diff = reinterpret_cast<int16_t*>(at);                         at += kTileBlocks * 2;
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(at);    at += kTileBlocks * 8;
    uint16_t* s_desc = reinterpret_cast<uint16_t*>(at);                       at += kDescCap * 2;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(at);                       at += 4096;
    unsigned long long* s_first = reinterpret_cast<unsigned long long*>(at);  at += 8192;
    uint32_t* s_scan = reinterpret_cast<uint32_t*>(at);
    __shared__ unsigned long long s_off;

    const int tid = threadIdx.x;
    const uint32_t first = blockIdx.x * kTileBlocks;
    const int nb = static_cast<int>(min(static_cast<uint32_t>(kTileBlocks), p.nblocks - first));

    // the global minima seen so far bound what this tile can still contribute: after the first tiles almost no
    // key is smaller, so the shared-memory atomicMin below is rarely executed
    for (int i = tid; i < 1024; i += kTileBlocks) { s_hist[i] = 0; s_first[i] = __ldcg(&p.g_first[i]); }
    stage_tile(tv, p.coef + static_cast<size_t>(first) * kCoefPerBlock, nb, tid, kTileBlocks);
    __syncthreads();

    // ---- per block: item count, DC difference, text key ----
    uint32_t lo = 0, hi = 0, count = 0;
    if (tid < nb) {
        load_mask(tv, tid, lo, hi);
        s_diff[tid] = static_cast<int16_t>(tv.dc[tid] - dc_predictor(tv, p.coef, first, tid));
        count = 1u + __popc(lo) + __popc(hi) + ((hi >> 31) ? 0u : 1u);        // DC + non-zero ACs + EOB unless coefficient 63 != 0
        const uint32_t g = first + tid, mcu = g / kBlocksPerMcu, k = g - mcu * kBlocksPerMcu;
        const uint32_t my = mcu / p.mcu_w, mx = mcu - my * p.mcu_w;
        s_key[tid] = 256ull * (k < 4 ? static_cast<unsigned long long>(my * 2 + (k >> 1)) * (2ull * p.mcu_w) + mx * 2 + (k & 1)
                                     : static_cast<unsigned long long>(k - 4) * p.n_mcu + mcu);
    }
    uint32_t total;
    const uint32_t base = block_exclusive_scan(count, s_scan, &total);
    if (tid == 0) {
        s_off = atomicAdd(p.cursor, static_cast<unsigned long long>(total));
        p.tile_cnt[blockIdx.x] = total;
    }
    __syncthreads();
    const unsigned long long off = s_off;
    if (tid == 0) p.tile_off[blockIdx.x] = off;
    uint32_t* __restrict__ out = p.items + off;

    for (uint32_t win = 0; win < total; win += kDescCap) {
        // ---- expand: descriptors (block << 7 | position) of the items in [win, win + kDescCap); position 0 = DC, 64 = EOB
        if (tid < nb && base < win + kDescCap && base + count > win) {
            uint32_t g = base;
            if (g >= win) s_desc[g - win] = static_cast<uint16_t>(tid << 7);
            ++g;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                uint32_t m = half ? hi : lo;
#pragma unroll 1
                while (m) {
                    const int pos = half * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    if (g - win < kDescCap) s_desc[g - win] = static_cast<uint16_t>((tid << 7) | pos);     // g < win wraps to a huge value
                    ++g;
                }
            }
            if (!(hi >> 31) && g - win < kDescCap) s_desc[g - win] = static_cast<uint16_t>((tid << 7) | 64);
        }
        __syncthreads();
        // ---- flat: consecutive lanes take consecutive items ----
        const uint32_t n = min(static_cast<uint32_t>(kDescCap), total - win);
        for (uint32_t j0 = 0; j0 < n; j0 += kTileBlocks) {              // uniform trip count: warp_count is a warp-wide operation
            const uint32_t j = j0 + tid;
            const bool live = j < n;
            int idx = 0, value = 0, nzrl = 0, table = 0, symbol = 0;
            unsigned long long key = 0;
            if (live) {
                const uint32_t d = s_desc[j], b = d >> 7, pos = d & 127;
                const int k = static_cast<int>(b % kBlocksPerMcu);          // tiles start on an MCU boundary
                const int tdc = k < 4 ? 0 : 2;
                key = s_key[b];
                if (pos == 0) {
                    value = s_diff[b];
                    symbol = category_of(value);
                    table = tdc;
                } else if (pos == 64) {
                    table = tdc + 1;
                    key += 129;
                } else {
                    uint32_t mlo, mhi;
                    load_mask(tv, b, mlo, mhi);
                    const unsigned long long below = ((static_cast<unsigned long long>(mhi) << 32) | mlo) & ((1ull << pos) - 1ull);
                    const int prev = below ? 63 - __clzll(static_cast<long long>(below)) : 0;
                    int run = static_cast<int>(pos) - prev - 1;
                    nzrl = run >> 4;
                    run &= 15;
                    value = slot_coef(tv, b, pos);
                    symbol = (run << 4) | category_of(value);
                    table = tdc + 1;
                    key += 2 * pos + 1;
                }
                idx = table * 256 + symbol;
                out[win + j] = make_item(table, symbol, nzrl, value);
            }
            warp_count(s_hist, s_first, live, idx, key);
            if (__any_sync(0xffffffffu, nzrl != 0)) {                   // ZRLs (runs of 16 zeros) are rare
                if (nzrl) {
                    const int zi = table * 256 + 0xF0;
                    atomicAdd(&s_hist[zi], static_cast<uint32_t>(nzrl));
                    const unsigned long long kz = key - 1;               // 2*pos: just before the symbol of position pos
                    if (kz < s_first[zi]) atomicMin(&s_first[zi], kz);
                }
            }
        }
        __syncthreads();
    }

    for (int i = tid; i < 1024; i += kTileBlocks) {
        const uint32_t n = s_hist[i];
        if (n) {
            atomicAdd(&p.g_hist[i], n);
            const unsigned long long k = s_first[i];
            if (k < __ldcg(&p.g_first[i])) atomicMin(&p.g_first[i], k);
        }
    }
}

int launch_symbol_stats(jpgenc_ctx* c) {
    const uint64_t n_mcu = static_cast<uint64_t>(c->mcu_w) * c->mcu_h, nblocks = n_mcu * kBlocksPerMcu;
    const unsigned grid = static_cast<unsigned>((nblocks + kTileBlocks - 1) / kTileBlocks);
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_hist, 0, 4 * 256 * sizeof(uint32_t), c->stream));
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_first, 0xFF, 4 * 256 * sizeof(unsigned long long), c->stream));
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_item_cursor, 0, sizeof(unsigned long long), c->stream));
    StatsParams p{};
    p.coef = c->d_coef;
    p.nblocks = static_cast<uint32_t>(nblocks);
    p.mcu_w = c->mcu_w;
    p.n_mcu = static_cast<uint32_t>(n_mcu);
    p.g_hist = c->d_hist;
    p.g_first = c->d_first;
    p.items = c->d_items;
    p.cursor = c->d_item_cursor;
    p.tile_off = c->d_tile_off;
    p.tile_cnt = c->d_tile_cnt;
    JPGENC_CUDA(c, cudaFuncSetAttribute(symbol_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatsSmem));
    symbol_stats_kernel<<<grid, kTileBlocks, kStatsSmem, c->stream>>>(p);
    JPGENC_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return JPGENC_OK;
}

}  // namespace jpgenc
