// Host half of KeyToU64IdMapper (/root/reference/src/external_integration/mod.rs:209-256): Key (u128) -> dense slot.
// Open addressing, linear probing, backward-shift deletion (no tombstones).  One 32-byte entry per bucket -- key and
// slot in the same cache line, two buckets per line -- because the table of a 10 M-key index (1 GiB) is far larger
// than any cache and a delta batch is a stream of random probes: one miss per probe instead of two, and `prefetch`
// lets a batch loop issue the miss of key i + 12 while it works on key i.  Plain C++ (no CUDA): unit-tested on the CPU
// (tests/test_key_slot_map.py).
#pragma once
#include <sys/mman.h>

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <new>

namespace pwknn {

struct Key {
    uint64_t lo, hi;
    bool operator==(const Key& o) const { return lo == o.lo && hi == o.hi; }
};
struct KeyHash {
    static uint64_t mix(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
        return x;
    }
    size_t operator()(const Key& k) const { return static_cast<size_t>(mix(k.lo ^ mix(k.hi + 0x9e3779b97f4a7c15ULL))); }
};

class alignas(64) KeySlotMap {  // one cache line per table object: shards owned by different threads never share one
public:
    static constexpr uint32_t kEmpty = 0xFFFFFFFFu;
    KeySlotMap() { rehash(1024); }
    size_t size() const { return size_; }
    size_t capacity() const { return mask_ + 1; }
    // *_h: the caller already has KeyHash()(k) (the sharded wrapper needs it for the shard anyway)
    bool find_h(const Key& k, size_t hash, uint32_t* slot) const {
        size_t i = hash & mask_;
        while (e_[i].slot != kEmpty) {
            if (e_[i].lo == k.lo && e_[i].hi == k.hi) { *slot = e_[i].slot; return true; }
            i = (i + 1) & mask_;
        }
        return false;
    }
    void set_h(const Key& k, size_t hash, uint32_t slot) {  // insert or overwrite
        if ((size_ + 1) * 2 > mask_ + 1) rehash((mask_ + 1) * 2);
        size_t i = hash & mask_;
        while (e_[i].slot != kEmpty) {
            if (e_[i].lo == k.lo && e_[i].hi == k.hi) { e_[i].slot = slot; return; }
            i = (i + 1) & mask_;
        }
        e_[i].lo = k.lo; e_[i].hi = k.hi; e_[i].slot = slot; e_[i].h32 = static_cast<uint32_t>(hash); ++size_;
    }
    bool erase_h(const Key& k, size_t hash, uint32_t* slot) {
        size_t i = hash & mask_;
        while (e_[i].slot != kEmpty) {
            if (e_[i].lo == k.lo && e_[i].hi == k.hi) {
                *slot = e_[i].slot;
                size_t hole = i, j = i;  // backward-shift: keep probe chains contiguous, no tombstones
                for (;;) {
                    j = (j + 1) & mask_;
                    if (e_[j].slot == kEmpty) break;
                    const size_t home = e_[j].h32 & mask_;  // the entry's own hash bits: no re-hashing of neighbours
                    if (((j - home) & mask_) >= ((j - hole) & mask_)) { e_[hole] = e_[j]; hole = j; }
                }
                e_[hole].slot = kEmpty; --size_;
                return true;
            }
            i = (i + 1) & mask_;
        }
        return false;
    }
    bool find(const Key& k, uint32_t* slot) const { return find_h(k, KeyHash()(k), slot); }
    void set(const Key& k, uint32_t slot) { set_h(k, KeyHash()(k), slot); }
    bool erase(const Key& k, uint32_t* slot) { return erase_h(k, KeyHash()(k), slot); }
    void reserve(size_t n) { size_t c = mask_ + 1; while (n * 2 > c) c *= 2; if (c != mask_ + 1) rehash(c); }
    void prefetch(const Key& k) const { __builtin_prefetch(&e_[KeyHash()(k) & mask_]); }
    void prefetch_h(size_t hash) const { __builtin_prefetch(&e_[hash & mask_]); }
private:
    // the low 32 hash bits ride along in the bucket (capacity <= 2^32): erase and rehash never hash a key again
    struct alignas(32) Entry { uint64_t lo, hi; uint32_t slot; uint32_t h32; uint32_t pad[2]; };
    // 2 MiB-aligned storage with MADV_HUGEPAGE: a random probe into a GiB-sized table is a TLB miss with 4 KiB pages
    // (the page walk costs as much as the data miss itself); huge pages cut the per-probe cost by about a third.
    static Entry* alloc_entries(size_t cap) {
        const size_t bytes = cap * sizeof(Entry), align = bytes >= (2u << 20) ? (2u << 20) : 4096;
        void* p = nullptr;
        if (posix_memalign(&p, align, (bytes + align - 1) / align * align) != 0) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
        if (bytes >= (2u << 20)) madvise(p, bytes, MADV_HUGEPAGE);
#endif
        Entry* e = static_cast<Entry*>(p);
        for (size_t i = 0; i < cap; ++i) { e[i].lo = 0; e[i].hi = 0; e[i].slot = kEmpty; e[i].h32 = 0; }
        return e;
    }
    void rehash(size_t cap) {
        Entry* old = e_;
        const size_t old_cap = old ? mask_ + 1 : 0;
        e_ = alloc_entries(cap); mask_ = cap - 1; size_ = 0;
        for (size_t i = 0; i < old_cap; ++i) {
            if (old[i].slot == kEmpty) continue;
            size_t j = old[i].h32 & mask_;  // keys are distinct: no comparison needed
            while (e_[j].slot != kEmpty) j = (j + 1) & mask_;
            e_[j] = old[i]; ++size_;
        }
        free(old);
    }
    Entry* e_ = nullptr;
    size_t mask_ = 0, size_ = 0;
public:
    ~KeySlotMap() { free(e_); }
    KeySlotMap(const KeySlotMap&) = delete;
    KeySlotMap& operator=(const KeySlotMap&) = delete;
};

// kShards independent tables, a key's shard chosen by the TOP bits of its hash (the bucket index uses the low ones).
// Same interface as one table for the sequential code; the batched swap-remove (slot_book.h) runs its hash-table
// phases on several host threads, each owning whole shards, so a mutation never races with another thread.
class ShardedKeyMap {
public:
    static constexpr uint32_t kShards = 16;
    static uint32_t shard_of_hash(size_t h) { return static_cast<uint32_t>(h >> 60) & (kShards - 1); }
    static uint32_t shard_of(const Key& k) { return shard_of_hash(KeyHash()(k)); }
    KeySlotMap& shard(uint32_t s) { return t_[s]; }
    const KeySlotMap& shard(uint32_t s) const { return t_[s]; }
    size_t size() const { size_t n = 0; for (const auto& t : t_) n += t.size(); return n; }
    bool find(const Key& k, uint32_t* slot) const { const size_t h = KeyHash()(k); return t_[shard_of_hash(h)].find_h(k, h, slot); }
    void set(const Key& k, uint32_t slot) { const size_t h = KeyHash()(k); t_[shard_of_hash(h)].set_h(k, h, slot); }
    bool erase(const Key& k, uint32_t* slot) { const size_t h = KeyHash()(k); return t_[shard_of_hash(h)].erase_h(k, h, slot); }
    // with the hash already at hand (a 128-bit key costs ~15 ns to hash: batch loops hash every key ONCE)
    bool find_h(const Key& k, size_t h, uint32_t* slot) const { return t_[shard_of_hash(h)].find_h(k, h, slot); }
    void set_h(const Key& k, size_t h, uint32_t slot) { t_[shard_of_hash(h)].set_h(k, h, slot); }
    bool erase_h(const Key& k, size_t h, uint32_t* slot) { return t_[shard_of_hash(h)].erase_h(k, h, slot); }
    void prefetch_h(size_t h) const { t_[shard_of_hash(h)].prefetch_h(h); }
    void reserve(size_t n) { for (auto& t : t_) t.reserve(n / kShards + n / (4 * kShards) + 64); }
    void prefetch(const Key& k) const { const size_t h = KeyHash()(k); t_[shard_of_hash(h)].prefetch_h(h); }
private:
    KeySlotMap t_[kShards];
};

}  // namespace pwknn
