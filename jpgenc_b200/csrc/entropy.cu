// K3 — Huffman encode + MCU interleave + 1-padding;  K4 — FF -> FF 00 byte stuffing.
//
// Replaces Image::doHuffmanEncoding, the per-MCU Bitstream concatenation, Bitstream::fill() and the stuffing
// operator<< (reference src/Image.cpp:737-829, 957-971; include/BitstreamGeneric.hpp:126-146, 182-195, 213-224, 242-248).
//
// K3 reads the symbol items K2 left in HBM (blockwalk.cuh), never the coefficients.  A tile is the item range of one
// K2 tile (384 blocks, 64 MCUs).  Its items are dealt to the threads in equal consecutive shares, so every thread does
// the same amount of work whatever the blocks look like: each thread sizes its share (code length + magnitude bits per
// item), the CTA scans the sizes, a decoupled look-back over per-tile status words turns the CTA total into the
// tile's global bit offset, and the codes are assembled MSB-first in a shared-memory bit buffer and written out as
// whole 32-bit words (only the two words a tile shares with its neighbours use atomicOr).  Tiles take their index from
// an atomic ticket, so a tile only ever waits for tiles that already run.
// K4 has the same structure at byte granularity: count FF bytes, scan, look back, compact.
#include "blockwalk.cuh"

namespace jpgenc {

constexpr int kBitBufWords = 8192;                      // 32 KB of bits per tile before falling back to global atomics

// MSB-first bit writer into 32-bit words ("word bit 31 is the earliest bit"); words are byte-swapped when they go
// to memory so that memory byte order is stream order.
template <bool kGlobal>
struct BitWriter {
    uint32_t* words;        // shared bit buffer (word 0 = the tile's first, partially foreign, word) or global raw scan
    uint32_t word;          // current word index
    uint32_t fill;          // bits already used in the current word
    uint32_t acc;

    __device__ __forceinline__ void start(uint32_t* base, unsigned long long bitpos) {
        words = base;
        word = static_cast<uint32_t>(bitpos >> 5);
        fill = static_cast<uint32_t>(bitpos & 31);
        acc = 0;
    }
    __device__ __forceinline__ void flush_word() {
        if (acc) atomicOr(&words[word], kGlobal ? __byte_perm(acc, 0, 0x0123) : acc);
    }
    __device__ __forceinline__ void put(uint32_t code, uint32_t n) {      // n <= 31, code < 2^n
        const unsigned long long v = static_cast<unsigned long long>(code) << (64 - fill - n);
        acc |= static_cast<uint32_t>(v >> 32);
        fill += n;
        if (fill >= 32) {
            flush_word();
            ++word;
            acc = static_cast<uint32_t>(v);
            fill -= 32;
        }
    }
    __device__ __forceinline__ void finish() { if (fill) flush_word(); }
};

struct EntropyParams {
    const uint32_t* items;              // K2's symbol stream
    const unsigned long long* tile_off; // [tiles] first item of tile
    const uint32_t* tile_cnt;           // [tiles] items of tile
    uint32_t nranges;                   // number of K2 tiles
    uint32_t ntiles;                    // number of K3 tiles = ceil(nranges / kPackGroup)
    const DeviceTables* tables;
    unsigned long long* status;     // look-back words, zeroed before launch
    uint32_t* ticket;               // zeroed before launch
    uint32_t* raw;                  // zeroed before launch, (total_bits+7)/8 bytes rounded up to words (+ slack)
    unsigned long long* total_out;  // [0] = total bits written (before padding)
};

// code bits of one item: `nz` ZRL codes first, then the symbol's code with the magnitude bits appended
// (src/Image.cpp:757-766: table[symbol] top `length` bits, then the category's bits)
template <class Put>
__device__ __forceinline__ void item_codes(uint32_t item, const uint32_t* s_tab, Put&& put) {
    const uint32_t* tab = s_tab + ((item >> 8) & 3u) * 256u;
    const uint32_t nz = (item >> 10) & 3u;
    if (nz) {
        const uint32_t z = tab[0xF0];
        for (uint32_t i = 0; i < nz; ++i) put(z & 0xFFFFu, z >> 16);
    }
    const uint32_t e = tab[item & 0xFFu], cat = item & 15u;
    put(((e & 0xFFFFu) << cat) | (item >> 12), (e >> 16) + cat);
}

__device__ __forceinline__ uint32_t item_bits(uint32_t item, const uint32_t* s_tab) {
    const uint32_t* tab = s_tab + ((item >> 8) & 3u) * 256u;
    return (tab[item & 0xFFu] >> 16) + (item & 15u) + ((item >> 10) & 3u) * (tab[0xF0] >> 16);
}

constexpr int kPackThreads = 256;
constexpr int kPackGroup = 4;                          // K2 tiles (item ranges) per K3 tile

// The items of a K3 tile are kPackGroup separate ranges of the item stream (K2 tiles claim their ranges in completion
// order).  for_items walks the slice [i0, i1) of their concatenation.
template <class F>
__device__ __forceinline__ void for_items(const uint32_t* __restrict__ items, const unsigned long long* s_off, const uint32_t* s_cum,
                                          uint32_t i0, uint32_t i1, F&& f) {
    int g = 0;
    while (g + 1 < kPackGroup && s_cum[g + 1] <= i0) ++g;
    for (uint32_t i = i0; i < i1; ++g) {
        const uint32_t end = min(i1, s_cum[g + 1]);
        const uint32_t* __restrict__ ptr = items + s_off[g] + (i - s_cum[g]);
        const uint32_t n = end - i;
        uint32_t j = 0;
        for (; j + 4 <= n; j += 4) {                    // four independent loads in flight
            const uint32_t a = __ldg(ptr + j), b = __ldg(ptr + j + 1), c = __ldg(ptr + j + 2), d = __ldg(ptr + j + 3);
            f(a); f(b); f(c); f(d);
        }
        for (; j < n; ++j) f(__ldg(ptr + j));
        i = end;
    }
}

__global__ void __launch_bounds__(kPackThreads) huffman_pack_kernel(const __grid_constant__ EntropyParams p) {
    __shared__ uint32_t s_tab[1024];                  // [4][256] (length << 16) | code
    __shared__ uint32_t s_bits[kBitBufWords + 1];
    __shared__ uint32_t s_scan[36];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_off[kPackGroup];
    __shared__ uint32_t s_cum[kPackGroup + 1];
    const int tid = threadIdx.x;

    if (tid == 0) {
        const uint32_t t = atomicAdd(p.ticket, 1u);
        s_tile = t;
        uint32_t cum = 0;
        for (int g = 0; g < kPackGroup; ++g) {
            const uint32_t k2 = t * kPackGroup + g;
            s_cum[g] = cum;
            s_off[g] = k2 < p.nranges ? p.tile_off[k2] : 0ull;
            cum += k2 < p.nranges ? p.tile_cnt[k2] : 0u;
        }
        s_cum[kPackGroup] = cum;
    }
    for (int i = tid; i < 1024; i += kPackThreads) s_tab[i] = (&p.tables->entry[0][0])[i];
    __syncthreads();
    const uint32_t tile_idx = s_tile;
    const uint32_t n = s_cum[kPackGroup];
    const uint32_t* __restrict__ items = p.items;
    // blocked arrangement: thread t owns items [t*per, t*per + per) so that it can merge their bits in registers
    const uint32_t per = (n + kPackThreads - 1) / kPackThreads;
    const uint32_t i0 = min(n, tid * per), i1 = min(n, i0 + per);

    uint32_t my_bits = 0;
    for_items(items, s_off, s_cum, i0, i1, [&](uint32_t item) { my_bits += item_bits(item, s_tab); });
    uint32_t tile_bits;
    const uint32_t local = block_exclusive_scan(my_bits, s_scan, &tile_bits);
    const bool last_tile = tile_idx + 1 == p.ntiles;
    const bool in_smem = tile_bits + 32 <= kBitBufWords * 32u;

    if (in_smem) {
        const uint32_t zero_words = ((tile_bits + 31) >> 5) + 1;
        for (uint32_t i = tid; i < zero_words; i += kPackThreads) s_bits[i] = 0;
        __syncthreads();
        // The codes are assembled at tile-relative bit positions, so this does not wait for the look-back: the last
        // warp resolves the tile's global bit offset while the others are still packing.
        if (tid >= kPackThreads - 32) {
            const unsigned long long b = lookback_exclusive(p.status, tile_idx, tile_bits);
            if (tid == kPackThreads - 32) s_base = b;
        }
        if (i0 < i1) {
            BitWriter<false> bw;
            bw.start(s_bits, local);
            for_items(items, s_off, s_cum, i0, i1, [&](uint32_t item) {
                item_codes(item, s_tab, [&](uint32_t code, uint32_t len) { bw.put(code, len); });
            });
            bw.finish();
        }
        __syncthreads();
        const unsigned long long base = s_base;
        const uint32_t lead = static_cast<uint32_t>(base & 31);             // bits of the first word owned by earlier tiles
        const uint32_t nwords = (lead + tile_bits + 31) >> 5;
        uint32_t* g = p.raw + (base >> 5);
        for (uint32_t i = tid; i < nwords; i += kPackThreads) {
            // global word i = tile-relative bits [32i - lead, 32i - lead + 32)
            const uint32_t v = __byte_perm(__funnelshift_r(s_bits[i], i ? s_bits[i - 1] : 0u, lead), 0, 0x0123);
            if (i == 0 || i == nwords - 1) { if (v) atomicOr(&g[i], v); }
            else g[i] = v;
        }
    } else {   // very dense tile: wait for the offset, then write straight to the (zeroed) global words
        if (tid < 32) {
            const unsigned long long b = lookback_exclusive(p.status, tile_idx, tile_bits);
            if (tid == 0) s_base = b;
        }
        __syncthreads();
        if (i0 < i1) {
            BitWriter<true> bw;
            bw.start(p.raw, s_base + local);
            for_items(items, s_off, s_cum, i0, i1, [&](uint32_t item) {
                item_codes(item, s_tab, [&](uint32_t code, uint32_t len) { bw.put(code, len); });
            });
            bw.finish();
        }
    }
    if (last_tile && tid == 0) {
        // 1-padding of Bitstream::fill() (BitstreamGeneric.hpp:242-248): the open byte is completed with ones
        const unsigned long long end = s_base + tile_bits;
        const uint32_t pad = static_cast<uint32_t>((8 - (end & 7)) & 7);
        if (pad) {
            BitWriter<true> bw;
            bw.start(p.raw, end);
            bw.put((1u << pad) - 1, pad);
            bw.finish();
        }
        p.total_out[0] = end;
    }
}

// ---- K4 ---------------------------------------------------------------------------------------------------
// FF -> FF 00 (BitstreamGeneric.hpp:213-224).  A tile is 16 KB of the raw scan.  Every thread counts the FF bytes of its
// 16 input bytes (SIMD byte compare), a CTA scan + look-back give the output position, the tile's output is assembled
// in shared memory (word stores; byte stores only at a thread's unaligned ends or where an FF actually occurs) and
// leaves as coalesced 32-bit stores, funnel-shifted to the alignment of the global position.
constexpr int kStuffThreads = 1024;
constexpr int kStuffBytesPerThread = 16;
constexpr int kStuffTile = kStuffThreads * kStuffBytesPerThread;

__global__ void __launch_bounds__(kStuffThreads) stuff_kernel(const uint8_t* __restrict__ raw, uint64_t nbytes,
                                                              uint8_t* __restrict__ out, unsigned long long* status,
                                                              uint32_t* ticket, unsigned long long* total_ff) {
    __shared__ alignas(16) uint8_t s_out[2 * kStuffTile + 16];
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_at = static_cast<uint64_t>(tile) * kStuffTile;
    const uint64_t at = tile_at + static_cast<uint64_t>(tid) * kStuffBytesPerThread;
    uint4 q = make_uint4(0, 0, 0, 0);
    int n = 0;
    if (at < nbytes) {
        n = static_cast<int>(umin64(kStuffBytesPerThread, nbytes - at));
        q = *reinterpret_cast<const uint4*>(raw + at);         // raw is padded to a multiple of 16 bytes and zero-filled
    }
    uint32_t wv[4] = {q.x, q.y, q.z, q.w};
    if (n < 16) {                                              // bytes past the end of the scan are not FF candidates
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j >= n) wv[j >> 2] &= ~(0xFFu << (8 * (j & 3)));
    }
    const uint32_t ff = (__popc(__vcmpeq4(wv[0], 0xFFFFFFFFu)) + __popc(__vcmpeq4(wv[1], 0xFFFFFFFFu)) +
                         __popc(__vcmpeq4(wv[2], 0xFFFFFFFFu)) + __popc(__vcmpeq4(wv[3], 0xFFFFFFFFu))) >> 3;
    uint32_t tile_ff;
    const uint32_t local = block_exclusive_scan(ff, s_scan, &tile_ff);
    if (tid >= kStuffThreads - 32) {                           // the last warp resolves the global position meanwhile
        const unsigned long long b = lookback_exclusive(status, tile, tile_ff);
        if (tid == kStuffThreads - 32) s_base = b;
    }
    // ---- assemble the tile's output at tile-relative positions ----
    uint32_t o = tid * kStuffBytesPerThread + local;
    if (ff == 0 && n == 16) {
        const uint32_t lead = (4u - (o & 3u)) & 3u;            // bytes up to the next word boundary
        if (lead == 0) {
            *reinterpret_cast<uint32_t*>(s_out + o) = wv[0];
            *reinterpret_cast<uint32_t*>(s_out + o + 4) = wv[1];
            *reinterpret_cast<uint32_t*>(s_out + o + 8) = wv[2];
            *reinterpret_cast<uint32_t*>(s_out + o + 12) = wv[3];
        } else {
            for (uint32_t j = 0; j < lead; ++j) s_out[o + j] = static_cast<uint8_t>(wv[0] >> (8 * j));
            const uint32_t sh = 8 * lead;
            uint32_t* w = reinterpret_cast<uint32_t*>(s_out + o + lead);
            w[0] = __funnelshift_r(wv[0], wv[1], sh);
            w[1] = __funnelshift_r(wv[1], wv[2], sh);
            w[2] = __funnelshift_r(wv[2], wv[3], sh);
            for (uint32_t j = 0; j < 4u - lead; ++j) s_out[o + lead + 12 + j] = static_cast<uint8_t>(wv[3] >> (sh + 8 * j));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < n) {
                const uint8_t b = static_cast<uint8_t>((wv[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                s_out[o++] = b;
                if (b == 0xFF) s_out[o++] = 0;
            }
        }
    }
    __syncthreads();
    // ---- copy out: global word w holds tile-relative bytes [4w - sh, 4w - sh + 4) ----
    const uint32_t in_tile = static_cast<uint32_t>(umin64(kStuffTile, nbytes - tile_at));
    const uint32_t len = in_tile + tile_ff;
    uint8_t* g = out + tile_at + s_base;
    const uint32_t sh = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g) & 3u);
    const uint32_t head = min(len, (4u - sh) & 3u);            // bytes before the first aligned global word
    const uint32_t nwords = (len - head) >> 2;
    if (tid < head) g[tid] = s_out[tid];
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_out);
    uint32_t* gw = reinterpret_cast<uint32_t*>(g + head);
    for (uint32_t w = tid; w < nwords; w += kStuffThreads)
        gw[w] = head ? __funnelshift_r(sw[w], sw[w + 1], 8 * head) : sw[w];
    const uint32_t tail0 = head + 4 * nwords;
    if (tid < len - tail0) g[tail0 + tid] = s_out[tail0 + tid];
    if (tid == 0 && tile_at + kStuffTile >= nbytes) total_ff[0] = s_base + tile_ff;
}

int launch_entropy(jpgenc_ctx* c, uint64_t total_bits) {
    const uint64_t n_mcu = static_cast<uint64_t>(c->mcu_w) * c->mcu_h, nblocks = n_mcu * kBlocksPerMcu;
    const uint64_t nbytes = (total_bits + 7) / 8;
    const uint32_t ranges = static_cast<uint32_t>((nblocks + kTileBlocks - 1) / kTileBlocks);
    const uint32_t tiles3 = (ranges + kPackGroup - 1) / kPackGroup;
    const uint32_t tiles4 = static_cast<uint32_t>((nbytes + kStuffTile - 1) / kStuffTile);
    // status words: [0,1] totals (bits written by K3, FF bytes stuffed by K4), then K3's tiles, then K4's
    unsigned long long* totals = c->d_lookback;
    unsigned long long* st3 = c->d_lookback + 2;
    unsigned long long* st4 = st3 + tiles3;
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_lookback, 0, (static_cast<size_t>(tiles3) + tiles4 + 2) * sizeof(unsigned long long), c->stream));
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_counters + 1, 0, 2 * sizeof(uint32_t), c->stream));
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_raw, 0, ((nbytes + 15) & ~15ull) + 64, c->stream));

    EntropyParams p{};
    p.items = c->d_items;
    p.tile_off = c->d_tile_off;
    p.tile_cnt = c->d_tile_cnt;
    p.nranges = ranges;
    p.ntiles = tiles3;
    p.tables = c->d_tables;
    p.status = st3;
    p.ticket = c->d_counters + 1;
    p.raw = c->d_raw;
    p.total_out = totals;
    huffman_pack_kernel<<<tiles3, kPackThreads, 0, c->stream>>>(p);
    JPGENC_CUDA(c, cudaGetLastError());
    if (tiles4) {
        stuff_kernel<<<tiles4, kStuffThreads, 0, c->stream>>>(reinterpret_cast<const uint8_t*>(c->d_raw), nbytes, c->d_scan,
                                                              st4, c->d_counters + 2, totals + 1);
        JPGENC_CUDA(c, cudaGetLastError());
    }
    c->launches += 2;
    return JPGENC_OK;
}

}  // namespace jpgenc
