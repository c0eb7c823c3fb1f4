// la3dm_b200 -- run detection on a sorted key array, shared by the voxel grid and the block binning.
#pragma once
#include "common.cuh"

namespace la3dm_b200 {

// ---- run detection on a sorted key array (shared with the binning stage) ------------------------------------------
// tile_sums[tile] = number of run heads among the first *d_n keys of the tile
static __global__ void k_run_count(const unsigned int *__restrict__ keys, const unsigned int *__restrict__ d_n,
                            unsigned int cap, unsigned int *tile_sums, const ScanCounters *__restrict__ c) {
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned int n = c->overflow ? 0u : min(*d_n, cap);
    const unsigned int base = blockIdx.x * kTile + threadIdx.x * kTileItems;
    unsigned int cnt = 0;
    if (base < n) {
        unsigned int prev = base ? keys[base - 1] : 0u;
#pragma unroll
        for (int k = 0; k < kTileItems; ++k) {
            const unsigned int i = base + k;
            if (i < n) {
                const unsigned int key = keys[i];
                cnt += (i == 0 || key != prev) ? 1u : 0u;
                prev = key;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_cnt;
}

}  // namespace la3dm_b200
