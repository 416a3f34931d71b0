// Batched swap-remove on the host: the slot bookkeeping of
//   NonFilteringExternalIndex::remove  (/root/reference/src/external_integration/brute_force_knn_integration.rs:146-170)
// -- per key, IN ORDER: unknown key -> MissingKey; else the last row moves into the hole and is re-mapped -- for a whole
// batch at once, with the hash-table work spread over several host threads.  Plain C++ (no CUDA), differential-tested
// against the literal sequential loop (tests/native/slot_book_test.cpp).
//
// Why it can be parallel although the swaps are sequential: the hash table is only needed (A) to translate the batch's
// keys to their slots BEFORE the batch and (C) to apply the NET effect afterwards (erase the removed keys, re-map the
// keys of the rows that ended up in a hole).  Both are independent per key -- (A) is read-only, (C) is partitioned by
// table shard.  The order-dependent part (B) -- which row is last when, which hole it fills, a row that moved and is
// removed later in the same batch, a key listed twice -- is replayed on two slot-indexed arrays only:
//   cur_of_orig[s] = where the row that was at slot s before the batch is now (or "removed"),
//   orig_of_cur[p] = which pre-batch row sits at position p now,
// one cache line per removed row, prefetched from the slots phase (A) found.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

#include "key_slot_map.h"

namespace pwknn {

constexpr uint32_t kBookInvalid = 0xFFFFFFFFu;

// One record per slot, valid for the batch whose generation it carries.  add() uses {gen, a = batch row that last
// wrote the slot}; remove() uses {gen, a = cur_of_orig, b = orig_of_cur, listed}.
struct SlotStamp {
    uint32_t gen, a, b, listed;
};

// runs body(part) for part in [0, nparts) on some threads and returns when all are done
using ParallelRunner = std::function<void(int nparts, const std::function<void(int)>& body)>;

inline void run_serial(int nparts, const std::function<void(int)>& body) { for (int i = 0; i < nparts; ++i) body(i); }

// keys: n x {lo, hi}.  status (optional): 0 ok, 1 MissingKey.  pairs: (src, dst) row moves appended, src = slot in the
// PRE-batch layout (always >= the final size), dst < final size, every dst once.  Returns the new size.
// `stamp` must have >= old size records; `gen` is a fresh generation (no record carries it yet).
inline uint64_t remove_batch(ShardedKeyMap& map, std::vector<Key>& slot_to_key, std::vector<SlotStamp>& stamp, uint32_t gen,
                             const uint64_t* keys, uint64_t n, int32_t* status, std::vector<uint32_t>* pairs,
                             const ParallelRunner& run, int threads) {
    const uint64_t old_size = slot_to_key.size();
    if (threads < 1) threads = 1;
    const int parts = n >= 2048 ? threads : 1;
    // ---- (A) keys -> pre-batch slots (read-only probes, any thread may read any shard)
    std::vector<uint32_t> found(n);
    std::vector<size_t> hsh(n);  // every key is hashed exactly once (a u128 key costs ~15 ns to hash)
    run(parts, [&](int part) {
        const uint64_t b = n * part / parts, e = n * (part + 1) / parts;
        for (uint64_t i = b; i < e; ++i) hsh[i] = KeyHash()(Key{keys[2 * i], keys[2 * i + 1]});
        constexpr uint64_t kAhead = 8;
        for (uint64_t i = b; i < std::min(b + kAhead, e); ++i) map.prefetch_h(hsh[i]);
        for (uint64_t i = b; i < e; ++i) {
            if (i + kAhead < e) map.prefetch_h(hsh[i + kAhead]);
            uint32_t s;
            found[i] = map.find_h(Key{keys[2 * i], keys[2 * i + 1]}, hsh[i], &s) ? s : kBookInvalid;
        }
    });
    // ---- (B) replay the swaps in order on the slot-indexed records
    auto rec = [&](uint32_t s) -> SlotStamp& {
        SlotStamp& r = stamp[s];
        if (r.gen != gen) r = SlotStamp{gen, s, s, 0};
        return r;
    };
    uint64_t cur_size = old_size;
    std::vector<uint32_t> touched;  // positions that received a row at some point
    std::vector<uint8_t> removed(n, 0);
    constexpr uint64_t kAheadB = 12;
    for (uint64_t i = 0; i < n; ++i) {
        if (i + kAheadB < n && found[i + kAheadB] != kBookInvalid) __builtin_prefetch(&stamp[found[i + kAheadB]]);
        const uint32_t r0 = found[i];
        if (r0 == kBookInvalid) { if (status) status[i] = 1; continue; }
        SlotStamp& victim = rec(r0);
        if (victim.a == kBookInvalid) { if (status) status[i] = 1; continue; }  // the same key earlier in this batch
        const uint32_t p = victim.a;                          // where that row is now
        const uint32_t last = static_cast<uint32_t>(cur_size - 1);
        cur_size -= 1;
        victim.a = kBookInvalid;
        removed[i] = 1;
        if (status) status[i] = 0;
        if (p != last) {                                      // rs:158-166: the last row moves into the hole
            const uint32_t o = rec(last).b;                   // pre-batch identity of the row at the tail
            SlotStamp& hole = rec(p);
            hole.b = o;
            rec(o).a = p;
            if (!hole.listed) { hole.listed = 1; touched.push_back(p); }
        }
    }
    const uint64_t new_size = cur_size;
    // net effect on the slot -> key table and the device rows: the positions that ended up holding another row.  The
    // final (dst, src) list is built in order (cheap: stamp records of touched positions are hot), the key reads /
    // writes -- random accesses into the slot -> key table -- run on the pool.
    struct Moved { uint32_t dst, src; Key key; size_t hash; };
    std::vector<Moved> moved;
    moved.reserve(touched.size());
    for (uint32_t p : touched) {
        if (p >= new_size) continue;
        const uint32_t o = stamp[p].b;
        if (o == p) continue;
        moved.push_back(Moved{p, o, Key{0, 0}, 0});          // o >= new_size: still holds the pre-batch key
        if (pairs) { pairs->push_back(o); pairs->push_back(p); }
    }
    const size_t nm = moved.size();
    const int mparts = nm >= 2048 ? parts : 1;
    constexpr size_t kAheadN = 8;
    run(mparts, [&](int part) {  // sources are >= new_size, destinations < new_size: reads and writes never alias
        const size_t b = nm * part / mparts, e = nm * (part + 1) / mparts;
        for (size_t j = b; j < e; ++j) {
            if (j + kAheadN < e) { __builtin_prefetch(&slot_to_key[moved[j + kAheadN].src]); __builtin_prefetch(&slot_to_key[moved[j + kAheadN].dst], 1); }
            moved[j].key = slot_to_key[moved[j].src];
            moved[j].hash = KeyHash()(moved[j].key);
            slot_to_key[moved[j].dst] = moved[j].key;
        }
    });
    slot_to_key.resize(new_size);
    // ---- (C) net effect on the hash table: every part owns whole shards, gathers its own work list (so that the
    // probes can be prefetched ahead) and applies it
    const int cparts = std::min<int>(parts, ShardedKeyMap::kShards);
    run(cparts, [&](int part) {
        struct Mut { const Key* key; size_t hash; uint32_t slot; };  // slot == kBookInvalid: erase
        std::vector<Mut> w;
        w.reserve((n + nm) / cparts + 64);
        for (uint64_t i = 0; i < n; ++i)
            if (removed[i] && static_cast<int>(ShardedKeyMap::shard_of_hash(hsh[i]) % cparts) == part)
                w.push_back(Mut{reinterpret_cast<const Key*>(keys + 2 * i), hsh[i], kBookInvalid});
        for (size_t j = 0; j < nm; ++j)
            if (static_cast<int>(ShardedKeyMap::shard_of_hash(moved[j].hash) % cparts) == part)
                w.push_back(Mut{&moved[j].key, moved[j].hash, moved[j].dst});
        uint32_t dummy;
        constexpr size_t kAheadC = 8;
        for (size_t j = 0; j < std::min(kAheadC, w.size()); ++j) map.prefetch_h(w[j].hash);
        for (size_t j = 0; j < w.size(); ++j) {
            if (j + kAheadC < w.size()) map.prefetch_h(w[j + kAheadC].hash);
            if (w[j].slot == kBookInvalid) map.erase_h(*w[j].key, w[j].hash, &dummy);
            else map.set_h(*w[j].key, w[j].hash, w[j].slot);
        }
    });
    return new_size;
}

}  // namespace pwknn
