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

__global__ void __launch_bounds__(kTileBlocks) symbol_stats_kernel(const int16_t* __restrict__ coef, uint64_t nblocks,
                                                                   uint32_t mcu_w, uint64_t n_mcu,
                                                                   uint32_t* __restrict__ g_hist,
                                                                   unsigned long long* __restrict__ g_first) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tile = smem;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem + kTileBytes);                           // [4][256]
    unsigned long long* s_first = reinterpret_cast<unsigned long long*>(smem + kTileBytes + 4096);  // [4][256]
    const int tid = threadIdx.x;
    const uint64_t first = static_cast<uint64_t>(blockIdx.x) * kTileBlocks;
    const int nb = static_cast<int>(umin64(kTileBlocks, nblocks - first));

    for (int i = tid; i < 1024; i += kTileBlocks) { s_hist[i] = 0; s_first[i] = ~0ull; }
    stage_tile(tile, coef + first * kCoefPerBlock, nb, tid, kTileBlocks);
    __syncthreads();

    if (tid < nb) {
        const int diff = slot_dc(tile, tid) - dc_predictor(tile, coef, first, tid);
        const uint64_t g = first + tid, mcu = g / kBlocksPerMcu;
        const int k = static_cast<int>(g % kBlocksPerMcu);
        const uint64_t mx = mcu % mcu_w, my = mcu / mcu_w;
        const uint64_t text_block = k < 4 ? (my * 2 + (k >> 1)) * (2ull * mcu_w) + mx * 2 + (k & 1)
                                          : static_cast<uint64_t>(k - 4) * n_mcu + mcu;
        const int tdc = k < 4 ? 0 : 2;
        uint32_t lo, hi;
        nonzero_mask(tile, tid, lo, hi);
        walk_block(tile, tid, diff, lo, hi, [&](int sym, int, int k) {
            const int idx = (k == 0 ? tdc : tdc + 1) * 256 + sym;
            atomicAdd(&s_hist[idx], 1u);
            const unsigned long long key = text_block * 256 + k;
            if (key < s_first[idx]) atomicMin(&s_first[idx], key);
        });
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
    const size_t smem = kTileBytes + 4096 + 8192;
    JPGENC_CUDA(c, cudaFuncSetAttribute(symbol_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const unsigned grid = static_cast<unsigned>((nblocks + kTileBlocks - 1) / kTileBlocks);
    symbol_stats_kernel<<<grid, kTileBlocks, smem, c->stream>>>(c->d_coef, nblocks, c->mcu_w, n_mcu, c->d_hist, c->d_first);
    JPGENC_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return JPGENC_OK;
}

}  // namespace jpgenc
