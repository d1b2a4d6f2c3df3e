// finalize_tables_kernel — the step between the device-side Huffman table build and K3 for BATCHES of frames.
//
// For one image the host does this work (capi.cu, entropy_frames): it turns the tables into the lookup form K3 reads and
// derives every size K3/K4 need from histogram x code length.  That costs a device-to-host copy, a synchronisation and a
// host-to-device copy in the middle of every pass, and on a host with few cores per GPU it is what a batch waits for.
// Here the same quantities are computed where the data already is:
//   * DeviceTables (entry / fast words) of every frame from the jpgenc_huff_table the device build produced
//     (reference semantics: Code = MSB-aligned code + length, include/Huffman.hpp:21-46);
//   * bits of every frame's scan = sum over symbols of count x (code length + magnitude bits) -- what
//     Image::doHuffmanEncoding would append (src/Image.cpp:751-769);
//   * the length of every frame's JFIF header (it varies with the number of symbols in the DHT segments,
//     JpegSegments.hpp:188-218);
//   * prefix sums over the frames of the pass: where each frame's raw scan, K4 tiles and output file start.
// One CTA per frame; the CTA that finishes last runs the prefix sums (a pass has at most 1024 frames) and writes the
// PassHeader, including the verdict whether the pass fits the buffers the host reserved.
#include "common.cuh"

namespace jpgenc {

constexpr size_t kStatsBytesF = 4096 + 8192;             // K2's per-frame statistics block (stats.cu)
constexpr unsigned long long kMissingSymbol = ~0ull;     // frame_bits marker: a symbol that occurs has no code

struct FinalizeParams {
    const jpgenc_huff_table* built;    // [4 * nframes]
    const uint8_t* stats;              // per frame: hist u32[4][256] | first-occurrence keys
    DeviceTables* tables;              // [nframes] out
    PassMeta meta;                     // out
    uint32_t nframes;
    uint32_t hdr_prefix_len;           // 0: K4 writes bare scans
    uint32_t tail;                     // bytes behind every scan (EOI)
    unsigned long long raw_limit, out_limit;
    uint32_t* ticket;
};

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += up;
    }
    return v;
}

__global__ void __launch_bounds__(256) finalize_tables_kernel(const __grid_constant__ FinalizeParams p) {
    __shared__ unsigned long long s_red[8];
    __shared__ uint32_t s_bad, s_last;
    const uint32_t f = blockIdx.x, s = threadIdx.x, lane = s & 31, warp = s >> 5, cat = s & 15;
    const uint32_t* hist = reinterpret_cast<const uint32_t*>(p.stats + static_cast<size_t>(f) * kStatsBytesF);
    if (s == 0) s_bad = 0;
    __syncthreads();
    unsigned long long bits = 0;
    bool bad = false;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const jpgenc_huff_table& tab = p.built[static_cast<size_t>(f) * 4 + t];
        const uint32_t len = tab.length[s];
        const uint32_t code = len ? tab.code_msb[s] >> (32 - len) : 0u;
        p.tables[f].entry[t][s] = len ? (len << 16) | code : 0u;
        p.tables[f].fast[t][s] = (len && len + cat <= 27) ? ((len + cat) << 27) | (code << cat) : 0u;
        const uint32_t h = hist[t * 256 + s];
        if (h) {
            bad |= len == 0;
            bits += static_cast<unsigned long long>(h) * (len + cat);
        }
    }
    if (bad) s_bad = 1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
    if (lane == 0) s_red[warp] = bits;
    __syncthreads();
    if (s == 0) {
        unsigned long long total = 0;
        for (int w = 0; w < 8; ++w) total += s_red[w];
        uint32_t hdr = 0;
        if (p.hdr_prefix_len) {
            hdr = p.hdr_prefix_len + 14;                                    // + SOS
            for (int t = 0; t < 4; ++t) {
                const jpgenc_huff_table& tab = p.built[static_cast<size_t>(f) * 4 + t];
                uint32_t nsym = 0;
                for (int i = 0; i < 16; ++i) nsym += tab.counts[i];
                hdr += 21 + nsym;
            }
        }
        p.meta.frame_bits[f] = s_bad ? kMissingSymbol : total;
        p.meta.raw_bytes[f] = (total + 7) / 8;
        p.meta.hdr_len[f] = hdr;
        __threadfence();
        s_last = atomicAdd(p.ticket, 1u) + 1 == p.nframes;
    }
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    // ---- the last CTA: prefix sums over the frames ----
    unsigned long long raw_c = 0, out_c = 0, sum_c = 0;
    uint32_t tile_c = 0, err = 0, err_frame = 0xFFFFFFFFu;
    for (uint32_t base = 0; base < p.nframes; base += 32) {
        const uint32_t g = base + lane;
        const bool valid = g < p.nframes;
        const unsigned long long fb = valid ? __ldcg(p.meta.frame_bits + g) : 1ull;
        const uint32_t e = !valid ? 0u : fb == 0 ? kPassNoSymbols : fb == kMissingSymbol ? kPassMissingSymbol : 0u;
        const unsigned long long nbytes = (valid && !e) ? __ldcg(p.meta.raw_bytes + g) : 0ull;
        const unsigned long long a = valid ? raw_slot_bytes(nbytes) : 0ull;
        const unsigned long long b = (nbytes + kK4TileBytes - 1) / kK4TileBytes;
        const unsigned long long c = valid ? __ldcg(p.meta.hdr_len + g) + nbytes + p.tail : 0ull;
        const unsigned long long ia = warp_incl_scan(a, lane), ib = warp_incl_scan(b, lane), ic = warp_incl_scan(c, lane),
                                 in = warp_incl_scan(nbytes, lane);
        if (valid) {
            p.meta.raw_off[g] = raw_c + ia - a;
            p.meta.k4_tile0[g] = tile_c + static_cast<uint32_t>(ib - b);
            p.meta.file_base[g] = out_c + ic - c;
        }
        raw_c += __shfl_sync(0xffffffffu, ia, 31);
        tile_c += static_cast<uint32_t>(__shfl_sync(0xffffffffu, ib, 31));
        out_c += __shfl_sync(0xffffffffu, ic, 31);
        sum_c += __shfl_sync(0xffffffffu, in, 31);
        const unsigned hit = __ballot_sync(0xffffffffu, e != 0);
        if (hit && err == 0) {
            const int first = __ffs(hit) - 1;
            err = __shfl_sync(0xffffffffu, e, first);
            err_frame = base + first;
        }
    }
    if (lane == 0) {
        p.meta.k4_tile0[p.nframes] = tile_c;
        if (raw_c > p.raw_limit) err |= kPassRawOverflow;
        if (out_c + sum_c > p.out_limit) err |= kPassOutOverflow;     // every scan byte could be an FF
        PassHeader h{};
        h.raw_total = raw_c;
        h.out_total = out_c;
        h.raw_sum = sum_c;
        h.k4_tiles = tile_c;
        h.error = err;
        h.error_frame = err_frame;
        h.nframes = p.nframes;
        *p.meta.hdr = h;
        *p.ticket = 0u;                                               // ready for the next pass
    }
}

int launch_finalize_tables(jpgenc_ctx* c) {
    FinalizeParams p{};
    p.built = c->d_built_tables;
    p.stats = c->d_stats;
    p.tables = c->d_tables;
    p.meta = pass_meta_view(c->d_meta, c->nframes);
    p.nframes = c->nframes;
    p.hdr_prefix_len = c->file_mode ? c->hdr_prefix_len : 0u;
    p.tail = c->file_mode ? 2u : 0u;
    p.raw_limit = c->raw_limit;
    p.out_limit = c->out_limit;
    p.ticket = c->d_counters + kCntFinalize;
    finalize_tables_kernel<<<c->nframes, 256, 0, c->stream>>>(p);
    JPGENC_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return JPGENC_OK;
}

}  // namespace jpgenc
