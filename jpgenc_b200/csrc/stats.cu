// K2 — DC differencing + run-length/category symbol generation, reduced to what the Huffman table build needs:
// a 4 x 256 symbol histogram and, per symbol, the position of its first occurrence in the reference's symbol text.
//
// Replaces Image::applyDCdifferenceCoding, Image::doRLEandCategoryCoding and the symbol-text collection
// (reference src/Image.cpp:638-735, 888-906; include/Coding.hpp:148-283).  The first-occurrence key is needed because
// generateHuffmanCode (src/Huffman.cpp:6-14) feeds package_merge in unordered_map iteration order, which depends on the
// order symbols first appear.  Text order (src/Image.cpp:892-906): Y blocks raster over the block grid,
// chroma = all Cb blocks then all Cr blocks; DC and AC symbols go to separate texts.
#include "blockwalk.cuh"

namespace jpgenc {

// One histogram update for the whole warp: lanes with the same bin are counted by their lowest lane, so the hot
// symbols (EOB, the common DC category) cost one shared-memory atomic instead of up to 32 serialised ones.
__device__ __forceinline__ void warp_count(uint32_t* s_hist, unsigned long long* s_first, bool has, int idx,
                                           unsigned long long key) {
    const int lane = threadIdx.x & 31;
    const unsigned grp = __match_any_sync(0xffffffffu, has ? idx : (0x10000 | lane));
    if (has) {
        if (lane == __ffs(grp) - 1) atomicAdd(&s_hist[idx], static_cast<uint32_t>(__popc(grp)));
        if (key < s_first[idx]) atomicMin(&s_first[idx], key);
    }
}

__global__ void __launch_bounds__(kTileBlocks) symbol_stats_kernel(const int16_t* __restrict__ coef, uint64_t nblocks,
                                                                   uint32_t mcu_w, uint64_t n_mcu,
                                                                   uint32_t* __restrict__ g_hist,
                                                                   unsigned long long* __restrict__ g_first) {
    extern __shared__ __align__(128) uint8_t smem[];
    const TileView tv = tile_view(smem);
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem + kTileSmemBytes);                           // [4][256]
    unsigned long long* s_first = reinterpret_cast<unsigned long long*>(smem + kTileSmemBytes + 4096);  // [4][256]
    const int tid = threadIdx.x;
    const uint64_t first = static_cast<uint64_t>(blockIdx.x) * kTileBlocks;
    const int nb = static_cast<int>(umin64(kTileBlocks, nblocks - first));

    for (int i = tid; i < 1024; i += kTileBlocks) { s_hist[i] = 0; s_first[i] = ~0ull; }
    stage_tile(tv, coef + first * kCoefPerBlock, nb, tid, kTileBlocks);
    __syncthreads();

    // every lane of the warp runs the same loop (the histogram update is a warp-wide operation); `live` masks lanes
    // past the end of the image
    const bool live = tid < nb;
    uint32_t lo = 0, hi = 0;
    int diff = 0;
    if (live) {
        load_mask(tv, tid, lo, hi);
        diff = tv.dc[tid] - dc_predictor(tv, coef, first, tid);
    }
    const uint64_t g = first + tid, mcu = g / kBlocksPerMcu;
    const int k = static_cast<int>(g % kBlocksPerMcu);
    const uint64_t mx = mcu % mcu_w, my = mcu / mcu_w;
    const unsigned long long text_key = 256ull * (k < 4 ? (my * 2 + (k >> 1)) * (2ull * mcu_w) + mx * 2 + (k & 1)
                                                          : static_cast<uint64_t>(k - 4) * n_mcu + mcu);
    const int tdc = (k < 4 ? 0 : 2) * 256, tac = tdc + 256;

    // DC entry: one warp-aggregated update per block row of the warp (few distinct categories per warp)
    warp_count(s_hist, s_first, live, tdc + category_of(diff), text_key);
    // AC entries: per-lane loop over the non-zero coefficients only
    int prev = 0;
    if (live) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t m = half ? hi : lo;
#pragma unroll 1
            while (m) {
                const int pos = half * 32 + __ffs(m) - 1;
                m &= m - 1;
                int run = pos - prev - 1;
                prev = pos;
                const int v = slot_coef(tv, tid, pos);
                if (run > 15) {                                                             // ZRL(s)
                    atomicAdd(&s_hist[tac + 0xF0], static_cast<uint32_t>(run >> 4));
                    if (text_key + 2 * pos < s_first[tac + 0xF0]) atomicMin(&s_first[tac + 0xF0], text_key + 2 * pos);
                    run &= 15;
                }
                const int idx = tac + ((run << 4) | category_of(v));
                atomicAdd(&s_hist[idx], 1u);
                if (text_key + 2 * pos + 1 < s_first[idx]) atomicMin(&s_first[idx], text_key + 2 * pos + 1);
            }
        }
    }
    // EOB: the same bin for every luma (resp. chroma) lane -> count with a ballot, one atomic per warp and table
    {
        const bool eob = live && prev != 63;
        const unsigned by = __ballot_sync(0xffffffffu, eob && k < 4), bc = __ballot_sync(0xffffffffu, eob && k >= 4);
        if ((tid & 31) == 0) {
            if (by) atomicAdd(&s_hist[256], static_cast<uint32_t>(__popc(by)));
            if (bc) atomicAdd(&s_hist[768], static_cast<uint32_t>(__popc(bc)));
        }
        if (eob && text_key + 129 < s_first[tac]) atomicMin(&s_first[tac], text_key + 129);
    }

    __syncthreads();
    for (int i = tid; i < 1024; i += kTileBlocks) {
        const uint32_t n = s_hist[i];
        if (n) {
            atomicAdd(&g_hist[i], n);
            atomicMin(&g_first[i], s_first[i]);
        }
    }
}

int launch_symbol_stats(jpgenc_ctx* c) {
    const uint64_t n_mcu = static_cast<uint64_t>(c->mcu_w) * c->mcu_h, nblocks = n_mcu * kBlocksPerMcu;
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_hist, 0, 4 * 256 * sizeof(uint32_t), c->stream));
    JPGENC_CUDA(c, cudaMemsetAsync(c->d_first, 0xFF, 4 * 256 * sizeof(unsigned long long), c->stream));
    const size_t smem = kTileSmemBytes + 4096 + 8192;
    JPGENC_CUDA(c, cudaFuncSetAttribute(symbol_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const unsigned grid = static_cast<unsigned>((nblocks + kTileBlocks - 1) / kTileBlocks);
    symbol_stats_kernel<<<grid, kTileBlocks, smem, c->stream>>>(c->d_coef, nblocks, c->mcu_w, n_mcu, c->d_hist, c->d_first);
    JPGENC_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return JPGENC_OK;
}

}  // namespace jpgenc
