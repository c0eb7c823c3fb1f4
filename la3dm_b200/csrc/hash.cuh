// la3dm_b200 -- the persistent block map's key -> slot table: open addressing, linear probing, <= 50 % load
// (replaces the reference's std::unordered_map<BlockHashKey, Block *> block_arr, include/bgkoctomap/bgkoctomap.h).
#pragma once
#include "common.cuh"

namespace la3dm_b200 {

__device__ inline int hash_find(const long long *__restrict__ hkeys, const int *__restrict__ hvals, size_t mask,
                                long long key) {
    size_t h = (size_t) mix64((unsigned long long) key) & mask;
    while (true) {
        const long long k = hkeys[h];
        if (k == key) return hvals[h];
        if (k == -1) return -1;
        h = (h + 1) & mask;
    }
}

__device__ inline void hash_insert(long long *hkeys, int *hvals, size_t mask, long long key, int val) {
    size_t h = (size_t) mix64((unsigned long long) key) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS((unsigned long long *) &hkeys[h], (unsigned long long) -1LL,
                                                  (unsigned long long) key);
        if (prev == (unsigned long long) -1LL || prev == (unsigned long long) key) { hvals[h] = val; return; }
        h = (h + 1) & mask;
    }
}

}  // namespace la3dm_b200
