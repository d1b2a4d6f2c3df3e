// Device-side synthetic inputs and the L2 eviction helper used between timed iterations.
#include "common.cuh"

namespace jpgenc {

__device__ __forceinline__ uint32_t hash32(uint32_t v) {
    v ^= v >> 16; v *= 0x7feb352du;
    v ^= v >> 15; v *= 0x846ca68bu;
    v ^= v >> 16;
    return v;
}

// gradient + 32-px checker + hash noise; identical to jpgenc_b200/synth.py::synth_rgb
__global__ void synth_rgb_kernel(uint8_t* __restrict__ out, uint32_t w, uint32_t h, uint32_t seed) {
    const uint64_t n = static_cast<uint64_t>(w) * h * 3;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t c = static_cast<uint32_t>(i % 3);
        const uint64_t pix = i / 3;
        const uint32_t x = static_cast<uint32_t>(pix % w), y = static_cast<uint32_t>(pix / w);
        const uint32_t v = static_cast<uint32_t>(i) + seed * 0x9E3779B9u;     // ((y*W+x)*3 + c + s*golden) mod 2^32
        const int noise = static_cast<int>(hash32(v) & 31u) - 16;
        int base;
        if (c == 0) base = static_cast<int>((255ull * x) / max(w - 1, 1u));
        else if (c == 1) base = static_cast<int>((255ull * y) / max(h - 1, 1u));
        else base = static_cast<int>((255ull * (x + y)) / max(w + h - 2, 1u));
        const int chk = 40 * (((x >> 5) ^ (y >> 5)) & 1);
        out[i] = static_cast<uint8_t>(min(max(base + chk + noise, 0), 255));
    }
}

// microbenchmark samples: ((hash(i) & 255) - 128) as fp32, i = linear sample index
__global__ void synth_blocks_kernel(float* __restrict__ out, uint64_t n) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<float>(static_cast<int>(hash32(static_cast<uint32_t>(i)) & 255u) - 128);
}

__global__ void flush_kernel(uint4* __restrict__ buf, uint64_t n, uint32_t tag) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        buf[i] = make_uint4(tag, tag, tag, tag);
}

int launch_synth_rgb(jpgenc_ctx* c, uint8_t* d, uint32_t w, uint32_t h, uint32_t seed) {
    synth_rgb_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(d, w, h, seed);
    JPGENC_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return JPGENC_OK;
}

int launch_synth_blocks(jpgenc_ctx* c, float* d, uint64_t nblocks) {
    synth_blocks_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(d, nblocks * 64);
    JPGENC_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return JPGENC_OK;
}

int launch_flush(jpgenc_ctx* c) {
    static uint32_t tag = 0;
    flush_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(static_cast<uint4*>(c->d_flush), c->flush_bytes / 16, ++tag);
    JPGENC_CUDA(c, cudaGetLastError());
    return JPGENC_OK;
}

}  // namespace jpgenc
