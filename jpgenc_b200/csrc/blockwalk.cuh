// Shared pieces of K2/K3/K4: staging MCU-ordered coefficient tiles in shared memory, the symbol-item format, CTA scan.
#pragma once
#include "common.cuh"

namespace jpgenc {

constexpr int kTileBlocks = 384;                       // 64 MCUs; one thread per block

// 8-bit non-zero flags of one 16-byte chunk (bit j = coefficient j of the chunk != 0).  Branch-free: VIMNMX.U16x2
// turns every halfword into a 0/1 flag.
__device__ __forceinline__ uint32_t chunk_flags(const uint4& q) {
    // the four 0/1 pairs never overlap, so the shifts and ORs are three multiply-adds
    const uint32_t f = (__vminu2(q.w, 0x00010001u) * 4u + __vminu2(q.z, 0x00010001u)) * 16u +
                       (__vminu2(q.y, 0x00010001u) * 4u + __vminu2(q.x, 0x00010001u));
    return (f | (f >> 15)) & 0xFFu;                 // even bits: low halfwords, odd bits: high halfwords
}

// Per-tile summary of the coefficient blocks a CTA works on, in shared memory:
//   flags  one byte per 16-byte chunk = chunk_flags(); 8 bytes per block = the block's 64-bit non-zero mask over the
//          zigzag positions (bit 0, the DC position, is cleared: it is not an AC coefficient)
//   dc     the block's DC coefficient
// The coefficients themselves are not kept: the few non-zero ones are re-read through L1/L2 when their items are built.
struct TileView {
    uint8_t* flags;
    int16_t* dc;
};
constexpr int kTileSmemBytes = kTileBlocks * 8 + kTileBlocks * 2;
constexpr int kItemsPerBlockMax = 64;                  // DC + 63 AC, or DC + at most 62 AC + EOB
constexpr int kSlabItems = kTileBlocks * kItemsPerBlockMax;   // item slots reserved per tile (worst case; typical use: a few %)

__device__ __forceinline__ TileView tile_view(uint8_t* smem) {
    return TileView{smem, reinterpret_cast<int16_t*>(smem + kTileBlocks * 8)};
}

// Coalesced pass over `nb` blocks of global memory: every thread takes 16-byte chunks (four loads in flight) and
// reduces each to its non-zero flags while the data is in registers.
__device__ __forceinline__ void scan_chunk(const TileView& tv, int j, const uint4& q) {
    const uint32_t f = chunk_flags(q);
    const bool head = (j & 7) == 0;
    tv.flags[j] = static_cast<uint8_t>(head ? (f & 0xFEu) : f);
    if (head) tv.dc[j >> 3] = static_cast<int16_t>(q.x & 0xFFFFu);
}

__device__ __forceinline__ void scan_tile(const TileView& tv, const int16_t* __restrict__ gsrc, int nb, int tid, int nthreads) {
    const uint4* g = reinterpret_cast<const uint4*>(gsrc);
    const int chunks = nb * 8;
    constexpr int kInFlight = 4;
    if (nb == kTileBlocks) {                        // full tile: no bounds checks
#pragma unroll
        for (int j0 = 0; j0 < kTileBlocks * 8; j0 += kInFlight * kTileBlocks) {
            uint4 q[kInFlight];
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) q[u] = __ldg(g + j0 + u * kTileBlocks + tid);
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) scan_chunk(tv, j0 + u * kTileBlocks + tid, q[u]);
        }
        return;
    }
    for (int j0 = 0; j0 < chunks; j0 += kInFlight * nthreads) {
        uint4 q[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int j = j0 + u * nthreads + tid;
            q[u] = j < chunks ? __ldg(g + j) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int j = j0 + u * nthreads + tid;
            if (j < chunks) scan_chunk(tv, j, q[u]);
        }
    }
}

__device__ __forceinline__ void load_mask(const TileView& tv, int slot, uint32_t& lo, uint32_t& hi) {
    const uint2 m = *reinterpret_cast<const uint2*>(tv.flags + slot * 8);
    lo = m.x;
    hi = m.y;
}

// DC predictor of block `t` of a tile that starts at global block `first` (a multiple of 6):
// Y follows MCU order (src/Image.cpp:640-659), Cb and Cr their own raster order (src/Image.cpp:661-677).
__device__ __forceinline__ int dc_predictor(const TileView& tv, const int16_t* __restrict__ coef, uint64_t first, int t) {
    const int k = t % kBlocksPerMcu, lm = t / kBlocksPerMcu;
    if (k >= 1 && k <= 3) return tv.dc[t - 1];
    const int back = (k == 0) ? 3 : 6;          // Y00 <- previous MCU's Y11 ; Cb/Cr <- previous MCU's Cb/Cr
    if (lm > 0) return tv.dc[t - back];
    if (first == 0) return 0;
    return coef[(first + t - back) * kCoefPerBlock];
}

__device__ __forceinline__ int category_of(int v) { return 32 - __clz(abs(v)); }   // 0 for v == 0

// ---- symbol items ------------------------------------------------------------------------------------------
// K2 turns every block into the reference's symbol sequence (include/Coding.hpp:148-183, 265-283) exactly once and
// leaves it in HBM as 32-bit items in scan order (MCU order, zigzag order inside a block): tile t (384 blocks) fills the
// first tile_cnt[t] slots of its own fixed slab of kSlabItems slots, so no tile waits for another.  K3 only maps items
// to codes.  One item = one Huffman symbol with its magnitude bits, plus the ZRL symbols (0xF0) that
// precede it:
//   [7:0]   symbol: DC category, or (run << 4) | category, or 0x00 = EOB
//   [9:8]   table id (0 Y_DC, 1 Y_AC, 2 C_DC, 3 C_AC)
//   [11:10] number of ZRL symbols emitted before this symbol (0..3)
//   [26:12] magnitude bits (category = symbol & 15 of them; value for v > 0, v - 1 truncated for v < 0, Coding.hpp:206-212)
//   [31:27] zero (K3 overwrites items in shared memory with (bit count << 27) | bits, which never has these bits zero)
__device__ __forceinline__ uint32_t make_item(int table, int symbol, int nzrl, int value) {
    const uint32_t cat = symbol & 15;
    const uint32_t mag = static_cast<uint32_t>(value < 0 ? value - 1 : value) & ((1u << cat) - 1u);
    return static_cast<uint32_t>(symbol) | (static_cast<uint32_t>(table) << 8) | (static_cast<uint32_t>(nzrl) << 10) | (mag << 12);
}

// block-wide exclusive scan of one value per thread (blockDim.x <= 1024); `scratch` holds 33 words
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* scratch, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < nwarps ? scratch[lane] : 0, t = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += n;
        }
        scratch[lane] = t - s;                       // exclusive warp offsets
        if (lane == 31) scratch[32] = t;             // grand total
    }
    __syncthreads();
    const uint32_t r = scratch[warp] + inc - v;
    *total = scratch[32];
    return r;
}

}  // namespace jpgenc
