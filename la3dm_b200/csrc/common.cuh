// la3dm_b200 -- shared declarations of the CUDA implementation (sm_100a).
//
// Whole library is compiled with -fmad=false: every fp32 product and sum below is rounded separately, exactly like a
// stock x86-64 build of the reference (no FMA contraction), which is what makes keys / memberships bit-exact.
// Where a fused multiply-add is wanted (polynomials) it is written explicitly as __fmaf_rn.
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
// growable device buffer (grow-only; growth synchronises the stream -- it only happens while capacities settle)
// ------------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    // contents are NOT preserved; returns true when the buffer moved
    bool reserve(size_t bytes, cudaStream_t s) {
        if (bytes <= cap) return false;
        size_t want = bytes + 256;
        if (p) { LA3DM_CUDA(cudaStreamSynchronize(s)); LA3DM_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
        LA3DM_CUDA(cudaMalloc(&p, want));
        cap = want;
        return true;
    }
    // contents ARE preserved
    bool grow_keep(size_t bytes, cudaStream_t s) {
        if (bytes <= cap) return false;
        size_t want = bytes + 256;
        void *q = nullptr;
        LA3DM_CUDA(cudaMalloc(&q, want));
        if (p) {
            LA3DM_CUDA(cudaMemcpyAsync(q, p, cap, cudaMemcpyDeviceToDevice, s));
            LA3DM_CUDA(cudaStreamSynchronize(s));
            LA3DM_CUDA(cudaFree(p));
        }
        p = q;
        cap = want;
        return true;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------------------------------
// device-visible parameter block
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxDepth = 6;        // node index is unsigned short upstream (bgkoctree.cpp:9-16): fine up to depth 6
constexpr int kMaxAxis = 8192;      // per-axis capacity of the float-stepped block grid (blocks per axis per scan)
constexpr unsigned int kHeavyTot = 64;   // test blocks above this many neighbourhood points are predicted first
constexpr unsigned int kMegaTot = 2048;  // ... above this many they are cut into chunks predicted by several warps
constexpr unsigned int kMegaChunkPts = 256;
constexpr int kStPRUNED = 3;        // BGK/BGKL/GP numbering; BGKLV uses 4 (see include/la3dm_b200.h)

struct DevParams {
    int method;
    int depth;              // block_depth
    int nodes;              // nodes per block = (8^depth-1)/7
    int finest;             // 8^(depth-1)
    int layer_off[kMaxDepth + 1];
    int st_off;             // byte offset of the state bytes inside a block record (= nodes * 8)
    int rec_bytes;          // bytes per block record (multiple of 16)
    float resolution, block_size, half_size;
    float sf2, ell;
    float free_thresh, occupied_thresh, var_thresh, prior_A, prior_B;
    float min_W; int original_size;
    float noise, l, min_ivar, max_ivar, min_known_ivar;
    int pruned_state;       // 3 (4 for BGKLV)
    float def_a, def_b;     // default node floats: (prior_A, prior_B) or GP (0, min_ivar)
    // block_depth 3 only: the node-centre LUT is separable per axis (init_key_loc_map, bgkblock.cpp:7-32: bits 4 / 2 / 1
    // of a child index pick x / y / z at every level) -- ax_off[a][j] = offset on axis a of: j = 0..3 the four finest
    // coordinates (2 * octant bit + child bit), j = 4, 5 the two depth-1 coordinates, j = 6 the root (0)
    float ax_off[3][8];
};

// Multi-GPU: every rank holds a replica of the map; the rank that predicts a test block stores the result straight into
// the other replicas' pools over NVLink (peer-mapped pointers), from inside the predict kernel.
constexpr int kMaxPeers = 8;
struct PeerTable {
    unsigned char *pool[kMaxPeers];          // pool base of every rank's replica, as mapped on THIS device ([rank]: own)
    unsigned long long *flags[kMaxPeers];    // flags[r][q], in rank r's memory: last scan that rank q has pushed completely
                                             // ([kMaxPeers + q]: last la3dm_peer_sync rank q has pushed completely)
    int world, rank;
    int deferred;                            // 1: no per-scan stores; the owner marks its blocks dirty, la3dm_peer_sync pushes
};

// Which replica predicts (and, with peers attached, owns) a test block.  Peers: a fixed function of the block key, so that
// a block's state lives on one rank from scan to scan; row exchange (la3dm_shard_*): test-block index modulo world.
__host__ __device__ inline int block_owner(long long key, unsigned int t, int world, bool by_key) {
    if (world <= 1) return 0;
    if (!by_key) return (int) (t % (unsigned int) world);
    unsigned long long x = (unsigned long long) key;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return (int) ((x >> 16) % (unsigned long long) world);
}

// Arguments of one insert_pointcloud call.  They live in device memory (copied from a pinned host mirror at the head
// of the scan) so that every kernel of the scan has launch parameters that do not change from scan to scan -- which is
// what lets the whole scan be replayed as one CUDA graph.
struct ScanArgs {
    const float *xyz;       // the cloud (device memory)
    unsigned int n;         // points in the cloud
    int stride_f;           // floats per point record
    float ox, oy, oz;       // sensor origin
    float ds, inv_ds;       // ds_resolution and 1 / ds_resolution (fp32, like pcl::VoxelGrid::setLeafSize)
    float fr;               // free_resolution
    float max_range;
    float free_label;       // 0 (BGK) / -1 (GP)
    int frontend_only;
    int training_data;      // 1: xyz holds pre-labelled points (x y z label); no front-end, no kbar guard (insert_training_data)
    int shard_rank, shard_world;
    unsigned int n_blocks;  // blocks in the map before this scan
    unsigned int pool_cap;  // block slots allocated
    const float *beam_tab;  // beam_tab[e] = fr + fr + ... (e fp32 additions): distance of beam sample e
    unsigned int beam_tab_n;
    unsigned int heavy_tot; // test blocks with more neighbourhood points than this are predicted first (kHeavyTot)
    unsigned int mega_tot, mega_chunk;   // ... with more than mega_tot they are cut into chunks of mega_chunk points (kMegaTot, kMegaChunkPts)
    const PeerTable *peers; // attached replicas (nullptr: none)
    unsigned long long scan_seq;   // sequence number of this scan (peer completion flags)
    int ab_flags;           // A/B switches for profiling (LA3DM_AB), 0 in production
    // ingest (server-side steps before insert_pointcloud, bgkoctomap_server.cpp:70-86): sensor-frame cloud -> map frame
    // by tf (3 x 4 row-major), VoxelGrid prefilter at pre_ds, scan skipped unless more than min_points survive
    const float *raw_xyz;   // the caller's cloud (sensor frame)
    int raw_stride_f;
    float tf[12];
    float scan_ds, scan_inv_ds;   // ds_resolution of the insert_pointcloud that follows the prefilter
    int min_points;
    float4 *stage_cloud;    // transformed cloud, then the prefiltered one
};

// per-scan block grid: restates get_blocks_in_bbox (src/bgkoctomap/bgkoctomap.cpp:486-495) as a Cartesian product of
// three per-axis index sets generated with the same fp32 stepping.
struct GridDesc {
    long long base[3];      // first block index (absolute, 0..2^20) on each axis
    int n[3];               // index span on each axis (last - first + 1)
    int irregular;          // stepping skipped or repeated an index
    unsigned int n_cells;   // n[0] * n[1] * n[2] (0 when the scan has no training data or the grid overflowed)
    unsigned char present[3][kMaxAxis];
};

// Per-scan workspace capacities (elements).  The whole scan is enqueued without host synchronisation against these
// capacities; every kernel checks the device-side counts against them and raises an overflow bit instead of writing
// out of bounds.  The host looks at the bits once, at the end of the scan, grows what was too small (using the sizes
// the device reports) and replays the scan -- the persistent map is only touched after all checks have passed.
struct Caps {
    unsigned int points;    // cloud points (also the capacity of the first voxel grid's output)
    unsigned int raw;       // beam samples before the second voxel grid
    unsigned int train;     // training set (= points + raw: exact bound)
    unsigned int members;   // (block, entry) memberships
    unsigned int cells;     // cells of the scan's dense block grid
    unsigned int tests;     // test blocks
    unsigned int vg_cells;  // linear voxel-grid index space (sets the radix-sort bit count)
    unsigned int gp_store;  // GP: floats of factor storage (packed L + alpha per data block)
    unsigned int gp_n_max;  // GP: largest data block the per-leaf scratch is sized for
    unsigned int lv_active; // BGKLV: voxels that have training data in their query box
    bool operator==(const Caps &o) const {
        return points == o.points && raw == o.raw && train == o.train && members == o.members && cells == o.cells &&
               tests == o.tests && vg_cells == o.vg_cells && gp_store == o.gp_store && gp_n_max == o.gp_n_max &&
               lv_active == o.lv_active;
    }
};

enum : unsigned int {
    OVF_RAW = 1u, OVF_MEMBERS = 2u, OVF_CELLS = 4u, OVF_TESTS = 8u, OVF_EXTENT = 16u, OVF_POOL = 32u, OVF_VGCELLS = 64u,
    OVF_GPSTORE = 128u, OVF_GPN = 256u, OVF_LVACTIVE = 512u, OVF_PEER = 1024u,
    OVF_FAST = 2048u      // a list too long for the sort-free front-end (frontend_fused.cu): replay on the legacy pipeline
};

// counters the host reads back once per scan (pinned mirror)
struct ScanCounters {
    unsigned int n_ds_hits;      // voxel-grid output count of the cloud
    unsigned int n_hits;         // kept after the range filter
    unsigned int n_raw_frees;    // free samples before the second voxel grid
    unsigned int n_frees;        // after it
    unsigned int n_train;        // n_hits + n_frees (BGK/GP)
    unsigned int n_members;      // (block, entry) memberships
    unsigned int n_data_blocks;
    unsigned int n_cells;        // cells of the block grid of this scan
    unsigned int n_test_blocks;
    unsigned int n_new_blocks;
    unsigned int vg_passthrough[2];
    unsigned long long visits, updates, pairs;
    unsigned int n_leaves;
    unsigned int overflow;       // OVF_* bits
    unsigned int n_long_runs[2]; // voxel-grid runs handed to the long-run kernel, one CTA each
    unsigned int n_mid_runs[2];  // ... one warp each
    unsigned int grid_irregular;
    unsigned int n_light;        // this rank's test blocks that are not heavy (light_list)
    unsigned int n_mega, n_mega_chunks, work_next2, work_next3;   // mega blocks (kMegaTot) and their chunks
    unsigned int vg_cells_needed;
    unsigned int gp_n_max;       // GP: largest data block of the scan
    unsigned int lv_active;      // BGKLV: active voxels of the scan
    unsigned int work_next;      // dynamic work distribution of the predict kernel (units handed out so far)
    unsigned int n_heavy;        // test blocks with more than kHeavyTot training points in their ExtendedBlock
    unsigned int ctas_done;      // CTAs of the predict kernel that have pushed everything to the peers
    unsigned long long gp_store_needed;
    unsigned int n_extra;        // fused binning: memberships beyond an entry's first
    unsigned int _pad_extra;
    unsigned int fz_nlong[3][2]; // fused front-end: spans longer than kShortRun ([voxel grid 1, 2, binning][warp, CTA])
};

#ifndef LA3DM_TILE
#define LA3DM_TILE 512
#endif
constexpr int kTile = LA3DM_TILE;          // elements per CTA in the two-kernel compactions (count per tile, then place)
constexpr int kTileThreads = 256;
constexpr int kTileItems = kTile / kTileThreads;
constexpr int kLongRun = 2048;       // voxel-grid runs longer than this are summed by a whole CTA each
constexpr int kMidRun = 48;          // ... longer than this by a warp each; shorter ones by one thread
constexpr int kMaxLongRuns = 4096;
constexpr int kMaxMidRuns = 32768;

inline int ceil_div(long long a, int b) { return (int) ((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------------------------
// s = fl(s + x) repeated n times (fp32, round-to-nearest-even), evaluated without n dependent additions.
// Used for voxel-grid runs that consist of one value repeated many times (the sensor origin is pushed once per hit:
// src/bgkoctomap/bgkoctomap.cpp:404, so its voxel holds n_hits identical points and pcl's centroid sums them one by
// one).  Inside one binade of s the increment of the integer significand is constant, so whole binades are crossed in
// one step; ties, binade crossings, sign changes and non-normal numbers fall back to real additions.
// ------------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define LA3DM_HD __host__ __device__
#else
#define LA3DM_HD
#endif
LA3DM_HD inline unsigned int f2u(float f) { union { float f; unsigned int u; } c; c.f = f; return c.u; }
LA3DM_HD inline float u2f(unsigned int u) { union { float f; unsigned int u; } c; c.u = u; return c.f; }

LA3DM_HD inline float add_repeat(float s, float x, unsigned int n) {
    while (n) {
        const unsigned int xb = f2u(x), sb = f2u(s);
        const unsigned int ex = (xb >> 23) & 0xFFu, es = (sb >> 23) & 0xFFu;
        const bool ok = n >= 4 && ex != 0 && ex != 255 && es != 0 && es != 255 && ((xb ^ sb) >> 31) == 0 &&
                        (sb & 0x7FFFFFFFu) >= (xb & 0x7FFFFFFFu);
        if (!ok) { s = s + x; --n; continue; }
        const unsigned int shift = es - ex;
        if (shift >= 25) return s;                               // x < ulp(s)/2 for good: absorbed
        const unsigned int X = (xb & 0x7FFFFFu) | 0x800000u;
        unsigned int m = (sb & 0x7FFFFFu) | 0x800000u;
        const unsigned int q = X >> shift;
        if (q + 2u >= (1u << 24)) { s = s + x; --n; continue; }
        const unsigned int L = (1u << 24) - q - 2u;              // a step from m <= L stays inside the binade
        if (m > L) { s = s + x; --n; continue; }
        const unsigned int rem = shift ? (X & ((1u << shift) - 1u)) : 0u;
        const unsigned int half = shift ? (1u << (shift - 1)) : 0u;
        unsigned int d;
        if (shift && rem == half) {                              // tie: round to even, then the parity is stable
            m += q + ((m + q) & 1u);
            --n;
            d = q + (q & 1u);
        } else {
            d = q + ((shift && rem > half) ? 1u : 0u);
        }
        const unsigned int sign_exp = sb & 0xFF800000u;
        if (d == 0) return u2f(sign_exp | (m & 0x7FFFFFu));
        if (m <= L && n) {
            unsigned int k = (L - m) / d + 1u;
            if (k > n) k = n;
            m += k * d;
            n -= k;
        }
        s = u2f(sign_exp | (m & 0x7FFFFFu));
    }
    return s;
}

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

// Sum of tile_sums[0 .. tile) and of tile_sums[0 .. n_tiles) computed by the whole CTA (blockDim.x threads, a multiple
// of 32, <= 1024); both results are valid in every thread.  smem: 66 elements.
template <typename T>
__device__ inline void block_tile_prefix(const T *__restrict__ tile_sums, unsigned int tile, unsigned int n_tiles,
                                         T *smem, T &prefix, T &total) {
    T acc = 0, all = 0;
    for (unsigned int j = threadIdx.x; j < n_tiles; j += blockDim.x) {
        const T v = tile_sums[j];
        all += v;
        if (j < tile) acc += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        all += __shfl_xor_sync(0xffffffffu, all, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { smem[threadIdx.x >> 5] = acc; smem[33 + (threadIdx.x >> 5)] = all; }
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < (blockDim.x >> 5) ? smem[threadIdx.x] : (T) 0;
        T w = threadIdx.x < (blockDim.x >> 5) ? smem[33 + threadIdx.x] : (T) 0;
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            w += __shfl_xor_sync(0xffffffffu, w, o);
        }
        if (threadIdx.x == 0) { smem[32] = v; smem[65] = w; }
    }
    __syncthreads();
    prefix = smem[32];
    total = smem[65];
    __syncthreads();
}

// CTA-wide exclusive scan of one value per thread (blockDim.x a multiple of 32, <= 1024); returns the exclusive prefix
// of this thread and the CTA total in `total`.  smem: 33 elements.
template <typename T>
__device__ inline T block_exclusive_scan(T v, T *smem, T &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    __syncthreads();
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        T w = lane < nw ? smem[lane] : (T) 0;
        T winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const T up = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += up;
        }
        if (lane < nw) smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    total = smem[32];
    const T r = smem[warp] + inc - v;
    __syncthreads();
    return r;
}
#endif

}  // namespace la3dm_b200
