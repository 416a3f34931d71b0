// CPU unit test + micro-benchmark of pathway_b200/csrc/key_slot_map.h against std::unordered_map:
// randomised set / find / erase (backward-shift deletion keeps every probe chain intact), then the access pattern of a
// churn round (erase + re-insert of 1 % of the keys) with and without prefetching.  Prints "OK ..." and timings.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <unordered_map>

#include "../../pathway_b200/csrc/key_slot_map.h"

using namespace pwknn;
struct StdHash { size_t operator()(const Key& k) const { return KeyHash()(k); } };

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 200000;
    std::mt19937_64 rng(7);
    KeySlotMap m;
    std::unordered_map<Key, uint32_t, StdHash> ref;
    std::vector<Key> keys;
    // randomised differential test (small key space -> long probe chains, many collisions of erase with chains)
    for (int it = 0; it < 400000; ++it) {
        const Key k{rng() % 5000, rng() % 3};
        const int op = rng() % 4;
        uint32_t a = 0;
        if (op < 2) { const uint32_t v = static_cast<uint32_t>(rng() % 1000000); m.set(k, v); ref[k] = v; }
        else if (op == 2) { const bool f1 = m.erase(k, &a); auto f2 = ref.find(k); if (f1 != (f2 != ref.end()) || (f1 && a != f2->second)) { printf("FAIL erase\n"); return 1; } if (f1) ref.erase(f2); }
        else { const bool f1 = m.find(k, &a); auto f2 = ref.find(k); if (f1 != (f2 != ref.end()) || (f1 && a != f2->second)) { printf("FAIL find\n"); return 1; } }
        if (m.size() != ref.size()) { printf("FAIL size\n"); return 1; }
    }
    for (auto& kv : ref) { uint32_t a; if (!m.find(kv.first, &a) || a != kv.second) { printf("FAIL final\n"); return 1; } }
    // benchmark: n keys, then rounds of erase + re-insert of 1 %
    KeySlotMap big;
    keys.resize(n);
    for (size_t i = 0; i < n; ++i) { keys[i] = Key{rng(), rng()}; big.set(keys[i], static_cast<uint32_t>(i)); }
    const size_t nd = n / 100 ? n / 100 : 1;
    std::vector<size_t> victims(nd);
    double t_plain = 0, t_pref = 0;
    for (int round = 0; round < 20; ++round) {
        std::vector<Key> vk(nd);  // a delta batch hands the keys over as one contiguous array
        for (size_t i = 0; i < nd; ++i) { victims[i] = rng() % n; vk[i] = keys[victims[i]]; }
        for (int mode = 0; mode < 2; ++mode) {
            const auto t0 = std::chrono::steady_clock::now();
            uint32_t s;
            for (size_t i = 0; i < nd; ++i) { if (mode && i + 12 < nd) big.prefetch(vk[i + 12]); big.erase(vk[i], &s); }
            for (size_t i = 0; i < nd; ++i) { if (mode && i + 12 < nd) big.prefetch(vk[i + 12]); if (!big.find(vk[i], &s)) big.set(vk[i], static_cast<uint32_t>(victims[i])); }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            (mode ? t_pref : t_plain) += dt;
        }
    }
    if (big.size() != n) { printf("FAIL churn size %zu\n", big.size()); return 1; }
    printf("OK keys=%zu capacity=%zu ns_per_op plain=%.1f prefetch=%.1f\n", n, big.capacity(), t_plain / (20.0 * 3 * nd) * 1e9, t_pref / (20.0 * 3 * nd) * 1e9);
    return 0;
}
