// Stress of the table builders' hand-over (jpgenc_b200/csrc/host_pools.hpp: TablePool) without a GPU: arm / build cycles
// with 1..3 workers on random histograms, every result compared with a direct jpgenc_build_huffman call.
//   table_pool_probe <workers> <cycles>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../jpgenc_b200/csrc/host_pools.hpp"

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main(int argc, char** argv) {
    const int workers = argc > 1 ? std::atoi(argv[1]) : 1, cycles = argc > 2 ? std::atoi(argv[2]) : 1000;
    jpgenc::TablePool pool(workers);
    uint32_t seed = 12345u + static_cast<uint32_t>(workers);
    static uint32_t count[4][256];
    static uint64_t first[4][256];
    for (int it = 0; it < cycles; ++it) {
        for (int t = 0; t < 4; ++t) {
            const uint32_t nsym = 1 + lcg(seed) % (t == 1 ? 60 : t == 3 ? 30 : 12);
            std::memset(count[t], 0, sizeof count[t]);
            std::memset(first[t], 0xFF, sizeof first[t]);
            for (uint32_t k = 0; k < nsym; ++k) {
                const uint32_t s = lcg(seed) & 255u;
                count[t][s] = 1 + lcg(seed) % (k % 3 ? 9u : 100000u);      // many ties among the small counts
                first[t][s] = (static_cast<uint64_t>(lcg(seed)) << 8) | k;
            }
        }
        pool.arm();
        if (it % 7 == 3) {                                   // the error path of the caller: workers released without work
            if (pool.build(nullptr, nullptr, nullptr) != JPGENC_OK) { std::printf("release failed\n"); return 1; }
            continue;
        }
        for (volatile int spin = 0; spin < static_cast<int>(lcg(seed) % 3000); ++spin) {}   // K2 "runs" for a while
        jpgenc_huff_table got[4], want[4];
        std::memset(got, 0xAB, sizeof got);
        if (pool.build(count, first, got) != JPGENC_OK) { std::printf("cycle %d: build failed\n", it); return 1; }
        for (int t = 0; t < 4; ++t) {
            if (jpgenc_build_huffman(count[t], first[t], &want[t]) != JPGENC_OK) { std::printf("direct build failed\n"); return 1; }
            if (std::memcmp(&got[t], &want[t], sizeof want[t]) != 0) { std::printf("cycle %d table %d differs\n", it, t); return 1; }
        }
    }
    std::printf("ok %d workers %d cycles\n", workers, cycles);
    return 0;
}
