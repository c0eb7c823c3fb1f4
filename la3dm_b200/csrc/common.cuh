// la3dm_b200 -- shared declarations of the CUDA implementation (sm_100a).
//
// Whole library is compiled with -fmad=false: every fp32 product and sum below is rounded separately, exactly like a
// stock x86-64 build of the reference (no FMA contraction), which is what makes keys / memberships bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/la3dm_b200.h"

namespace la3dm_b200 {

// ------------------------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------------------------
struct CudaError {
    cudaError_t code;
    const char *file;
    int line;
};

#define LA3DM_CUDA(expr)                                                       \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) throw ::la3dm_b200::CudaError{_e, __FILE__, __LINE__}; \
    } while (0)

struct StatusError {
    int status;
    std::string msg;
};

// ------------------------------------------------------------------------------------------------------------------
// growable device buffer (grow-only, amortised; growth synchronises the stream)
// ------------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    // contents are NOT preserved
    void reserve(size_t bytes, cudaStream_t s) {
        if (bytes <= cap) return;
        size_t want = bytes + bytes / 2 + 256;
        if (p) { LA3DM_CUDA(cudaStreamSynchronize(s)); LA3DM_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
        LA3DM_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    // contents ARE preserved
    void grow_keep(size_t bytes, cudaStream_t s) {
        if (bytes <= cap) return;
        size_t want = bytes + bytes / 2 + 256;
        void *q = nullptr;
        LA3DM_CUDA(cudaMalloc(&q, want));
        if (p) {
            LA3DM_CUDA(cudaMemcpyAsync(q, p, cap, cudaMemcpyDeviceToDevice, s));
            LA3DM_CUDA(cudaStreamSynchronize(s));
            LA3DM_CUDA(cudaFree(p));
        }
        p = q;
        cap = want;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------------------------------
// device-visible parameter block
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxDepth = 6;        // node index is unsigned short upstream (bgkoctree.cpp:9-16): fine up to depth 6
constexpr int kMaxAxis = 8192;      // per-axis capacity of the float-stepped block grid (blocks per axis per scan)
constexpr int kStPRUNED = 3;        // BGK/BGKL/GP numbering; BGKLV uses 4 (see include/la3dm_b200.h)

struct DevParams {
    int method;
    int depth;              // block_depth
    int nodes;              // nodes per block = (8^depth-1)/7
    int finest;             // 8^(depth-1)
    int layer_off[kMaxDepth + 1];
    float resolution, block_size, half_size;
    float sf2, ell;
    float free_thresh, occupied_thresh, var_thresh, prior_A, prior_B;
    float min_W; int original_size;
    float noise, l, min_ivar, max_ivar, min_known_ivar;
    int pruned_state;       // 3 (4 for BGKLV)
    float def_a, def_b;     // default node floats: (prior_A, prior_B) or GP (0, min_ivar)
};

// per-scan block grid: restates get_blocks_in_bbox (src/bgkoctomap/bgkoctomap.cpp:486-495) as a Cartesian product of
// three per-axis index sets generated with the same fp32 stepping.
struct GridDesc {
    long long base[3];      // first block index (absolute, 0..2^20) on each axis
    int n[3];               // index span on each axis (last - first + 1)
    int irregular;          // stepping skipped or repeated an index
    int overflow;           // span > kMaxAxis or nx*ny*nz >= 2^32
    unsigned char present[3][kMaxAxis];
};

// counters the host reads back (pinned mirror)
struct ScanCounters {
    unsigned int n_ds_hits;      // voxel-grid output count of the cloud
    unsigned int n_hits;         // kept after the range filter
    unsigned int n_raw_frees;    // free samples before the second voxel grid
    unsigned int n_frees;        // after it
    unsigned int n_train;        // n_hits + n_frees (BGK/GP)
    unsigned int n_members;      // (block, entry) memberships
    unsigned int n_data_blocks;
    unsigned int n_cand;
    unsigned int n_test_blocks;
    unsigned int n_new_blocks;
    unsigned int vg_passthrough[2];
    unsigned long long visits, updates, pairs;
    unsigned int n_leaves;
    unsigned int pad;
};

// ------------------------------------------------------------------------------------------------------------------
// device helpers shared by several kernels
// ------------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// block_to_hash_key, one axis (src/bgkoctomap/bgkblock.cpp:73-77): int64(x / (double) size + 524288.5)
__host__ __device__ inline long long axis_index(float x, float block_size) {
    return (long long) ((double) x / (double) block_size + 524288.5);
}
// hash_key_to_block, one axis (src/bgkoctomap/bgkblock.cpp:79-83): (i - 524288) * size, int64 -> float then fp32 mul
__host__ __device__ inline float axis_center(long long i, float block_size) {
    return (float) (i - 524288) * block_size;
}
__host__ __device__ inline long long make_key(long long ix, long long iy, long long iz) {
    return (ix << 40) | (iy << 20) | iz;
}

// order-preserving float <-> uint for atomicMin/atomicMax
__device__ inline unsigned int float_flip(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ inline float float_unflip(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
#endif

inline int ceil_div(long long a, int b) { return (int) ((a + b - 1) / b); }

}  // namespace la3dm_b200
