// Stress of HostPool::parallel_for (jpgenc_b200/csrc/host_pools.hpp) without a GPU: several caller threads (the pipeline lanes
// of a threaded batch) issue calls at once on one pool; every job must run exactly once and every call must return only
// after all of its jobs have finished.
//   host_pool_probe <workers> <callers> <calls per caller>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../jpgenc_b200/csrc/host_pools.hpp"

int main(int argc, char** argv) {
    const unsigned workers = argc > 1 ? std::atoi(argv[1]) : 3, callers = argc > 2 ? std::atoi(argv[2]) : 4;
    const int calls = argc > 3 ? std::atoi(argv[3]) : 1000;
    jpgenc::HostPool pool(workers);
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (unsigned c = 0; c < callers; ++c)
        th.emplace_back([&, c] {
            uint32_t seed = 777u + c;
            std::vector<int> hits(64);
            for (int k = 0; k < calls; ++k) {
                seed = seed * 1664525u + 1013904223u;
                const uint32_t n = (seed >> 9) % 64;                         // 0 jobs is a legal call
                std::fill(hits.begin(), hits.end(), 0);
                std::atomic<uint32_t> done{0};
                pool.parallel_for(n, [&](uint32_t i) {
                    ++hits[i];                                                // each job owns its slot
                    for (volatile int spin = 0; spin < static_cast<int>((seed >> 3) % 200); ++spin) {}
                    done.fetch_add(1, std::memory_order_relaxed);
                });
                if (done.load() != n) bad.fetch_add(1);                       // returned before its jobs had finished
                for (uint32_t i = 0; i < 64; ++i)
                    if (hits[i] != (i < n ? 1 : 0)) bad.fetch_add(1);
            }
        });
    for (std::thread& t : th) t.join();
    if (bad.load()) { std::printf("%d violations\n", bad.load()); return 1; }
    std::printf("ok %u workers %u callers %d calls\n", workers, callers, calls);
    return 0;
}
