// Correctness + region-latency check of pathway_b200/csrc/host_pool.h: every part of every region runs exactly once,
// regions issued back to back and after idle gaps (workers asleep) both complete; prints the per-region overhead.
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <thread>

#include "../../pathway_b200/csrc/host_pool.h"

using namespace pwknn;
int main() {
    HostPool pool(6);
    std::vector<std::atomic<int>> hits(64);
    long total = 0;
    for (int round = 0; round < 3000; ++round) {
        const int nparts = 1 + round % 40;
        for (auto& h : hits) h.store(0);
        pool.parallel(nparts, [&](int part, int np) { if (np != nparts) std::abort(); hits[part].fetch_add(1); });
        for (int p = 0; p < 64; ++p) if (hits[p].load() != (p < nparts ? 1 : 0)) { printf("FAIL round %d part %d\n", round, p); return 1; }
        total += nparts;
        if (round % 500 == 499) std::this_thread::sleep_for(std::chrono::milliseconds(5));  // let the workers fall asleep
    }
    // an exception thrown by a part (on whichever thread ran it) reaches the caller; the pool stays usable
    int caught = 0;
    for (int i = 0; i < 50; ++i) {
        try { pool.parallel(12, [&](int part, int) { if (part == 7) throw std::runtime_error("boom"); }); }
        catch (const std::runtime_error&) { ++caught; }
    }
    if (caught != 50) { printf("FAIL exceptions %d\n", caught); return 1; }
    pool.parallel(8, [&](int part, int) { hits[part].store(5); });
    for (int p = 0; p < 8; ++p) if (hits[p].load() != 5) { printf("FAIL after exception\n"); return 1; }
    // latency of empty regions: back to back (workers polling) and after a sleep (workers blocked)
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 2000; ++i) pool.parallel(6, [&](int, int) {});
    const double hot = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 2000 * 1e6;
    double cold = 0;
    for (int i = 0; i < 20; ++i) {
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
        auto t1 = std::chrono::steady_clock::now();
        pool.parallel(6, [&](int, int) {});
        cold += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count() * 1e6 / 20;
    }
    printf("OK host_pool parts=%ld region_us hot=%.2f cold=%.2f\n", total, hot, cold);
    return 0;
}
