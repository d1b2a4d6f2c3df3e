// Offline fuzz of the three table builders (allocation-free, container-driven, array restatement) against each other:
//   g++ -O2 tools/fuzz_tables.cpp -o /tmp/fuzz_tables -Ljpgenc_b200/lib -ljpgenc_b200 -Wl,-rpath,$PWD/jpgenc_b200/lib && /tmp/fuzz_tables 150000 <seed>
// Six alphabet families (tiny, small, all-ties, huge counts, powers of two, up to 256 symbols).  10^6 alphabets: 0 mismatches.
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include "../include/jpgenc_b200.h"
static uint32_t s = 99;
static uint32_t rnd() { s = s * 1664525u + 1013904223u; return s >> 8; }
int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 100000;
    s = argc > 2 ? atoi(argv[2]) : 99;
    static uint32_t count[256]; static uint64_t first[256];
    long bad = 0;
    for (int t = 0; t < trials; ++t) {
        memset(count, 0, sizeof count); memset(first, 0xFF, sizeof first);
        const int mode = rnd() % 6;
        const uint32_t nsym = mode == 5 ? 1 + rnd() % 256 : 1 + rnd() % (mode == 0 ? 8 : mode == 1 ? 40 : 180);
        const uint32_t maxc = mode == 2 ? 3 : mode == 3 ? 0x7FFFFFFFu : mode == 4 ? 64 : 100000;
        for (uint32_t k = 0; k < nsym; ++k) {
            const uint32_t sym = rnd() & 255u;
            uint32_t c = 1 + rnd() % maxc;
            if (mode == 4) c = 1u << (rnd() % 24);          // powers of two: sums collide all the time
            count[sym] = c;
            first[sym] = ((uint64_t)rnd() << 20) ^ rnd();
        }
        jpgenc_huff_table a, b, c3;
        const int ra = jpgenc_build_huffman(count, first, &a), rb = jpgenc_build_huffman_containers(count, first, &b), rc = jpgenc_build_huffman_arrays(count, first, &c3);
        if (ra != rb || ra != rc || (ra == 0 && (memcmp(&a, &b, sizeof a) || memcmp(&a, &c3, sizeof a)))) { if (bad < 5) printf("trial %d mode %d nsym %u differs (%d %d %d)\n", t, mode, nsym, ra, rb, rc); ++bad; }
    }
    printf("%d trials, %ld mismatches\n", trials, bad);
    return bad != 0;
}
