// Differential test of pathway_b200/csrc/slot_book.h (batched, partly parallel swap-remove) against the literal
// sequential loop of brute_force_knn_integration.rs:146-170: random batches with unknown keys, keys listed twice, rows
// that move and are removed later in the same batch, batches that empty the index.  "Rows" are modelled by one integer
// per slot, moved by the (src, dst) pairs the way move_rows_kernel does it (all sources read from the pre-batch layout).
#include <cstdio>
#include <random>
#include <thread>
#include <unordered_map>

#include "../../pathway_b200/csrc/slot_book.h"

using namespace pwknn;
struct StdHash { size_t operator()(const Key& k) const { return KeyHash()(k); } };

static void run_threads(int nparts, const std::function<void(int)>& body) {
    std::vector<std::thread> ts;
    for (int i = 1; i < nparts; ++i) ts.emplace_back([&, i] { body(i); });
    body(0);
    for (auto& t : ts) t.join();
}

int main() {
    std::mt19937_64 rng(11);
    for (int trial = 0; trial < 300; ++trial) {
        const uint64_t n0 = trial < 250 ? 1 + rng() % 3000 : 20000 + rng() % 30000;
        // model
        std::vector<Key> m_s2k;
        std::unordered_map<Key, uint32_t, StdHash> m_map;
        std::vector<uint64_t> m_rows;
        // under test
        ShardedKeyMap map;
        std::vector<Key> s2k;
        std::vector<uint64_t> rows;
        std::vector<SlotStamp> stamp;
        uint32_t gen = 0;
        for (uint64_t i = 0; i < n0; ++i) {
            const Key k{rng(), rng() & 3};
            if (m_map.count(k)) continue;
            m_map[k] = static_cast<uint32_t>(m_s2k.size()); m_s2k.push_back(k); m_rows.push_back(rng());
            map.set(k, static_cast<uint32_t>(s2k.size())); s2k.push_back(k); rows.push_back(m_rows.back());
        }
        for (int batch = 0; batch < 6 && !m_s2k.empty(); ++batch) {
            const uint64_t nb = trial % 7 == 0 ? m_s2k.size() + 5 : 1 + rng() % (m_s2k.size() + 3);
            std::vector<uint64_t> keys;
            for (uint64_t j = 0; j < nb; ++j) {
                const uint64_t kind = rng() % 10;
                Key k;
                if (kind == 0) k = Key{rng(), 77};                                   // unknown key
                else if (kind == 1 && !keys.empty()) k = Key{keys[2 * (rng() % (keys.size() / 2))], 0}, k.hi = 0;  // (mostly unknown)
                else k = m_s2k[rng() % m_s2k.size()];                                // live key, possibly listed twice
                if (kind == 2 && !m_s2k.empty()) k = m_s2k[m_s2k.size() - 1 - rng() % std::min<uint64_t>(m_s2k.size(), 4)];  // near the tail
                keys.push_back(k.lo); keys.push_back(k.hi);
            }
            // ---- model: the reference loop
            std::vector<int32_t> m_status(nb);
            for (uint64_t j = 0; j < nb; ++j) {
                const Key k{keys[2 * j], keys[2 * j + 1]};
                auto it = m_map.find(k);
                if (it == m_map.end()) { m_status[j] = 1; continue; }
                const uint32_t rid = it->second, last = static_cast<uint32_t>(m_s2k.size() - 1);
                m_map.erase(it);
                if (rid != last) { m_s2k[rid] = m_s2k[last]; m_rows[rid] = m_rows[last]; m_map[m_s2k[rid]] = rid; }
                m_s2k.pop_back(); m_rows.pop_back();
                m_status[j] = 0;
            }
            // ---- under test
            std::vector<int32_t> status(nb, -1);
            std::vector<uint32_t> pairs;
            if (stamp.size() < s2k.size()) stamp.resize(s2k.size(), SlotStamp{0, 0, 0, 0});
            ++gen;
            const int threads = 1 + static_cast<int>(rng() % 6);
            const uint64_t ns = remove_batch(map, s2k, stamp, gen, keys.data(), nb, status.data(), &pairs, threads > 1 ? ParallelRunner(run_threads) : ParallelRunner(run_serial), threads);
            const std::vector<uint64_t> before = rows;
            for (size_t j = 0; j + 1 < pairs.size(); j += 2) {
                if (pairs[j] < ns || pairs[j + 1] >= ns) { printf("FAIL pair range\n"); return 1; }
                rows[pairs[j + 1]] = before[pairs[j]];
            }
            rows.resize(ns);
            if (status != m_status) { printf("FAIL status trial %d\n", trial); return 1; }
            if (ns != m_s2k.size() || s2k.size() != m_s2k.size()) { printf("FAIL size trial %d\n", trial); return 1; }
            for (uint64_t s = 0; s < ns; ++s)
                if (!(s2k[s] == m_s2k[s]) || rows[s] != m_rows[s]) { printf("FAIL layout trial %d slot %llu\n", trial, (unsigned long long)s); return 1; }
            if (map.size() != m_map.size()) { printf("FAIL map size trial %d: %zu vs %zu\n", trial, map.size(), m_map.size()); return 1; }
            for (auto& kv : m_map) { uint32_t a; if (!map.find(kv.first, &a) || a != kv.second) { printf("FAIL map trial %d\n", trial); return 1; } }
        }
    }
    printf("OK slot_book\n");
    return 0;
}
