// la3dm_b200 -- the scan front-end, the block binning and the test-block plan as THREE cooperative kernels whose phases
// are separated by grid barriers instead of kernel boundaries, and without a single sort.
//
// What it replaces: frontend.cu + binning.cu (the "legacy" pipeline: ~42 launches per scan, three CUB radix sorts), i.e.
// get_training_data / downsample / beam_sample (src/bgkoctomap/bgkoctomap.cpp:383-458), the per-scan R-tree
// (:240-243, :497-552) and the TRAIN loop's bookkeeping (:250-284).  Same outputs, bit for bit -- the legacy pipeline
// stays in the library as the fallback (BGKL / BGKLV front-ends, lists too long for the ordering scheme below) and as
// the on-device cross-check of this file (LA3DM_LEGACY_FRONTEND=1).
//
// How grouping works without sorting (pcl::VoxelGrid needs, per voxel, its points IN INPUT ORDER -- the centroid is a
// sequential fp32 sum -- and the voxels in ascending index order; the binning needs per block its entries in entry order):
//   voxel grid : one bit per cell of the grid's bounding box; the rank of a set bit (prefix popcount) IS the voxel's
//                position in index order.  Points take an arrival ticket per voxel (atomicAdd), an exclusive scan of
//                the counts gives every voxel its span, points drop their index into the span (unordered), then every
//                point counts how many indices in its span are smaller than its own: that is its place in input order.
//                Spans are short (mean 2), the quadratic count is cheaper than one radix pass; a span above kFastRun
//                raises OVF_FAST and the host replays the scan on the legacy pipeline.
//   sensor origin: pushed once per kept hit (bgkoctomap.cpp:404) -- n_hits identical points in one voxel.  They never
//                enter the lists: the origin voxel's sum is rebuilt from the (few) other samples in it, each of which
//                knows how many origins precede it (its hit's ordinal + 1), with add_repeat() in between.
//   binning    : the same ticket / scan / drop / count scheme over the dense cells of the scan's block grid.
//
// Launched with cudaLaunchCooperativeKernel (co-residency of all CTAs is guaranteed by the driver, so the barrier
// cannot deadlock whatever else runs on the device); one CTA of 1024 threads per SM.
#include "engine.cuh"
#include "hash.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kFT = 1024;                      // threads per CTA
constexpr unsigned int kPad = 0xFFFFFFFFu;
constexpr unsigned int kFastRun = 2048;        // longest span ordered by counting
constexpr unsigned int kShortRun = 16;         // spans up to this long: every index counts for itself; longer: a warp per span
constexpr int kBeamTile = 128;                 // hits per tile of the beam phase
constexpr unsigned int kWideRun = 128;         // spans longer than this are ordered by a whole CTA (list B), else by a warp (list A)
constexpr int kTW = 64;                        // bitmap words per tile of the test-block compaction
constexpr unsigned int kMidVox = 48;           // voxels with more points than this are summed by a warp (kMidRun)
constexpr int kMidStage = 64;                  // points a warp stages at a time
constexpr unsigned int kMaxExtra = 1u << 16;   // memberships beyond an entry's first (points on a block boundary)

struct Extra { unsigned int entry, cell, k; };

struct FusedArgs {
    const ScanArgs *A;
    ScanCounters *c;
    unsigned int *mm;               // [3][6] flipped min / max: cloud, raw frees, training set
    const DevParams *P;
    GridDesc *g;
    unsigned int *bar;              // [4] grid-barrier counters (zeroed by k_fused_begin)
    unsigned long long *trace;      // LA3DM_FUSED_TRACE=1: %globaltimer of CTA 0 after every barrier ([3][64]); else null
    // voxel grid
    unsigned int *bits, *wpre;      // cell bitmap and the prefix popcount of its words
    unsigned int bits_words;        // capacity (words)
    unsigned int vg_cells_cap;
    unsigned int *vkey, *varr, *vcnt, *vstart, *vlist, *vsorted;
    unsigned int vcnt_cap;
    unsigned long long *tsum;       // tile sums (any phase)
    unsigned long long *bsum;       // (kept hits, free points) per kBeamTile downsampled hits (cleared in P0)
    unsigned int bsum_n;
    unsigned int *longs;            // spans longer than kShortRun: list A at [0, long_capA), list B behind it
    unsigned int long_capA, long_capB;
    float4 *hits_ds, *frees_raw, *xy;
    unsigned int *hit_cnt;
    unsigned int points_cap, raw_cap;
    // binning
    unsigned int *cell_cnt, *cell_db, *test_bits, *cell_test;
    unsigned int *mcell, *mk, *mlist;
    Extra *extra;
    float4 *pts;
    unsigned int *db_id, *db_start;
    unsigned int *skeys, *svals;    // BGKL: the (cell, entry) pairs in block order
    unsigned int cells_cap, members_cap, train_cap;
    // plan
    unsigned int *test_id;
    NeighbourPlan *plan;
    unsigned int *plan_db, *heavy_list, *light_list;
    uint4 *mega_list;
    unsigned int *chunk_mega;
    unsigned char *dirty, *touched;
    long long *hkeys;
    int *hvals;
    size_t hmask;
    long long *keys;
    unsigned char *pool;
    unsigned int *new_sums;         // new blocks per tile of 1024 test blocks
    unsigned int tests_cap;
    int init_records;
};

// ---- grid barrier ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_volatile(const unsigned int *p) {
    return *reinterpret_cast<const volatile unsigned int *>(p);
}

// all CTAs of the grid; `epoch` = arrivals expected so far (thread 0's copy is the one that counts)
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_mark(unsigned long long *trace, unsigned int slot) {
    if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[slot] = global_ns();
}
__device__ __forceinline__ void grid_sync(unsigned int *bar, unsigned int &epoch, unsigned long long *trace = nullptr) {
    __syncthreads();
    if (gridDim.x == 1 && !trace) return;                   // one CTA: block-level visibility is all that is needed
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        const unsigned long long t0 = trace ? global_ns() : 0ull;
        // release: everything this CTA wrote (ordered before by the bar.sync above) is visible to whoever acquires the
        // counter; acquire: ... and what the other CTAs wrote is visible to this CTA after the bar.sync below
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        while (ld_acquire(bar) < epoch) {}
        if (trace && blockIdx.x == 0) {
            const unsigned int ph = epoch / gridDim.x;             // 1-based phase just finished
            trace[2 * ph] = t0;                                    // CTA 0 arrived
            trace[2 * ph + 1] = global_ns();                       // everyone arrived
        }
    }
    __syncthreads();
}

// ---- CTA helpers ---------------------------------------------------------------------------------------------------------
// sum of ts[lo .. hi) by the whole CTA, valid in every thread.  smem: 34 elements.
template <typename T>
__device__ inline T block_sum_range(const T *ts, unsigned int lo, unsigned int hi, T *smem) {
    T a = 0;
    for (unsigned int j = lo + threadIdx.x; j < hi; j += blockDim.x) a += ts[j];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < (blockDim.x >> 5) ? smem[threadIdx.x] : (T) 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) smem[33] = v;
    }
    __syncthreads();
    const T r = smem[33];
    __syncthreads();
    return r;
}

template <typename T>
__device__ inline T block_sum(T a, T *smem) {
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < (blockDim.x >> 5) ? smem[threadIdx.x] : (T) 0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) smem[33] = v;
    }
    __syncthreads();
    const T r = smem[33];
    __syncthreads();
    return r;
}

// A CTA walks its tiles in ascending order; the exclusive prefix of a tile = the prefix of the CTA's previous tile +
// the tile sums in between.
template <typename T>
struct TileCarry {
    unsigned int next = 0;
    T acc = 0;
    __device__ T prefix(const T *ts, unsigned int tile, T *smem) {
        acc += block_sum_range(ts, next, tile, smem);
        next = tile;
        return acc;
    }
};

struct Box {
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    bool any = false;
    __device__ void add(float x, float y, float z) {
        mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
        mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
        mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
        any = true;
    }
    // every thread of the CTA; s_mm: 6 words
    __device__ void flush(unsigned int *mm, unsigned int *s_mm) {
        if (threadIdx.x < 6) s_mm[threadIdx.x] = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
        __syncthreads();
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float lo = mn[a], hi = mx[a];
                for (int o = 16; o > 0; o >>= 1) {
                    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                if ((threadIdx.x & 31) == 0) { atomicMin(&s_mm[a], float_flip(lo)); atomicMax(&s_mm[3 + a], float_flip(hi)); }
            }
        }
        __syncthreads();
        if (threadIdx.x < 3) { if (s_mm[threadIdx.x] != 0xFFFFFFFFu) atomicMin(&mm[threadIdx.x], s_mm[threadIdx.x]); }
        else if (threadIdx.x < 6) { if (s_mm[threadIdx.x] != 0u) atomicMax(&mm[threadIdx.x], s_mm[threadIdx.x]); }
        __syncthreads();
    }
};

__device__ inline void zero_words(unsigned int *p, size_t n) {
    uint4 *p4 = reinterpret_cast<uint4 *>(p);       // cudaMalloc'ed: 256-byte aligned
    const size_t n4 = n >> 2;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) p4[i] = z;
    for (size_t i = (n4 << 2) + (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = 0u;
}

// Set a bit unless a (possibly stale, never wrongly set) look says it is there already: next to the sensor thousands of
// points share a bitmap word, and same-address atomics serialise in L2.
__device__ __forceinline__ void set_bit(unsigned int *bits, unsigned int i) {
    const unsigned int m = 1u << (i & 31);
    if (!(*reinterpret_cast<volatile unsigned int *>(bits + (i >> 5)) & m)) atomicOr(&bits[i >> 5], m);
}

// a span longer than kShortRun goes on list A (a warp orders / sums it) or B (a CTA orders it); false: too long
__device__ __forceinline__ bool note_long_span(const FusedArgs &F, int pass, unsigned int r, unsigned int L) {
    if (L <= kShortRun) return true;
    if (L > kFastRun) return false;
    const int which = L > kWideRun ? 1 : 0;
    const unsigned int slot = atomicAdd(&F.c->fz_nlong[pass][which], 1u);
    if (slot >= (which ? F.long_capB : F.long_capA)) return false;
    F.longs[(which ? F.long_capA : 0u) + slot] = r;
    return true;
}

// ---- voxel grid (pcl::VoxelGrid restated, see frontend.cu / DESIGN.md) ---------------------------------------------------
struct VGFrame {
    int min_b[3];
    int mul1, mul2;
    bool passthrough;
    long long cells;
};

__device__ inline VGFrame vg_frame(const unsigned int *mm, float inv) {
    VGFrame f;
    float mn[3], mx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { mn[a] = float_unflip(mm[a]); mx[a] = float_unflip(mm[3 + a]); }
    const long long dx = (long long) ((mx[0] - mn[0]) * inv) + 1;
    const long long dy = (long long) ((mx[1] - mn[1]) * inv) + 1;
    const long long dz = (long long) ((mx[2] - mn[2]) * inv) + 1;
    f.passthrough = (dx * dy * dz) > 2147483647LL;   // "Leaf size is too small ... would overflow": output = input
    int div_b[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        f.min_b[a] = (int) floorf(mn[a] * inv);
        div_b[a] = (int) floorf(mx[a] * inv) - f.min_b[a] + 1;
    }
    f.mul1 = div_b[0];
    f.mul2 = div_b[0] * div_b[1];
    f.cells = (long long) div_b[0] * div_b[1] * div_b[2];
    return f;
}

__device__ __forceinline__ unsigned int vg_cell(const VGFrame &f, float inv, float x, float y, float z) {
    const int i0 = (int) (floorf(x * inv) - (float) f.min_b[0]);
    const int i1 = (int) (floorf(y * inv) - (float) f.min_b[1]);
    const int i2 = (int) (floorf(z * inv) - (float) f.min_b[2]);
    return (unsigned int) (i0 + i1 * f.mul1 + i2 * f.mul2);
}

// per-pass view: W = 0 the cloud -> hits_ds, W = 1 the raw free samples -> xy behind the hits
template <int W>
struct VGPass {
    const float *in;
    int stride;
    unsigned int n;
    bool pass;          // output = input (ds_resolution < 0, or pcl's index-overflow passthrough)
    bool dead;
    VGFrame f;
    float inv;
    __device__ VGPass(const FusedArgs &F, bool need_frame) {
        const ScanArgs *A = F.A;
        dead = ld_volatile(&F.c->overflow) != 0u;
        if (W == 0) { in = A->xyz; stride = A->stride_f; n = A->n; }
        else { in = reinterpret_cast<const float *>(F.frees_raw); stride = 4; n = F.c->n_raw_frees; }
        if (dead) n = 0;
        inv = A->inv_ds;
        const bool identity = A->ds < 0;
        f.passthrough = true;
        f.cells = 0;
        (void) need_frame;
        if (!identity && n > 0) f = vg_frame(F.mm + 6 * W, inv);
        pass = identity || f.passthrough;
    }
    __device__ unsigned int words() const { return pass ? 0u : (unsigned int) ((f.cells + 31) >> 5); }
};

// P1 / P11: cell of every point, one bit per occupied cell
template <int W>
__device__ void vg_keys(const FusedArgs &F) {
    VGPass<W> V(F, true);
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    if (gt == 0 && !V.dead && V.n > 0) {
        F.c->vg_passthrough[W] = V.pass ? 1u : 0u;
        if (!V.pass && (unsigned long long) V.f.cells > (unsigned long long) F.vg_cells_cap) {
            atomicOr(&F.c->overflow, OVF_VGCELLS);
            atomicMax(&F.c->vg_cells_needed,
                      (unsigned int) (V.f.cells > 0x80000000ll ? 0x80000000ll : V.f.cells));
        }
        if (W == 1 && !V.pass && V.n > 0 && (unsigned long long) V.f.cells <= (unsigned long long) F.vg_cells_cap) {
            // the sensor origin's voxel exists whatever else falls into it (the origin copies are not listed)
            const unsigned int cell = vg_cell(V.f, V.inv, F.A->ox, F.A->oy, F.A->oz);
            set_bit(F.bits, cell);
        }
    }
    if (V.pass || (unsigned long long) V.f.cells > (unsigned long long) F.vg_cells_cap) return;
    for (unsigned int i0 = gt; i0 < V.n; i0 += 4 * gn) {      // four points in flight per thread
        float px[4], py[4], pz[4];
        int pw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int i = i0 + k * gn;
            pw[k] = 0;
            if (i < V.n) {
                if (W == 1) {
                    const float4 v = reinterpret_cast<const float4 *>(V.in)[i];
                    px[k] = v.x; py[k] = v.y; pz[k] = v.z; pw[k] = __float_as_int(v.w);
                } else {
                    const float *p = V.in + (size_t) i * V.stride;
                    px[k] = p[0]; py[k] = p[1]; pz[k] = p[2];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int i = i0 + k * gn;
            if (i >= V.n) continue;
            if (W == 1 && pw[k] < 0) { F.vkey[i] = kPad; continue; }      // an origin copy
            const unsigned int cell = vg_cell(V.f, V.inv, px[k], py[k], pz[k]);
            F.vkey[i] = cell;
            set_bit(F.bits, cell);
        }
    }
}

// P2: popcount per tile of 4096 bitmap words (four words per thread)
constexpr unsigned int kWT = 4 * kFT;

template <int W>
__device__ void vg_bits_count(const FusedArgs &F, unsigned long long *smem) {
    VGPass<W> V(F, true);
    const unsigned int words = V.dead ? 0u : min(V.words(), F.bits_words);
    const unsigned int n_tiles = (words + kWT - 1) / kWT;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned int w = t * kWT + 4 * threadIdx.x;
        const uint4 b = w < words ? *reinterpret_cast<const uint4 *>(F.bits + w) : make_uint4(0u, 0u, 0u, 0u);   // (tail words are zero)
        const unsigned long long s = block_sum<unsigned long long>(
            (unsigned long long) (__popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w)), smem);
        if (threadIdx.x == 0) F.tsum[t] = s;
    }
}

// P3: wpre[w] = set bits in words 0 .. w - 1; the total = number of voxels
template <int W>
__device__ void vg_bits_prefix(const FusedArgs &F, unsigned long long *smem) {
    VGPass<W> V(F, true);
    const unsigned int words = V.dead ? 0u : min(V.words(), F.bits_words);
    const unsigned int n_tiles = (words + kWT - 1) / kWT;
    if (blockIdx.x == 0) {
        const unsigned long long total = block_sum_range<unsigned long long>(F.tsum, 0, n_tiles, smem);
        if (threadIdx.x == 0 && !V.dead) {
            const unsigned int nv = V.pass ? V.n : (unsigned int) total;
            if (W == 0) F.c->n_ds_hits = nv; else F.c->n_frees = nv;
        }
    }
    TileCarry<unsigned long long> carry;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned int w = t * kWT + 4 * threadIdx.x;
        const uint4 b = w < words ? *reinterpret_cast<const uint4 *>(F.bits + w) : make_uint4(0u, 0u, 0u, 0u);
        const unsigned long long prefix = carry.prefix(F.tsum, t, smem);
        const unsigned int c0 = __popc(b.x), c1 = __popc(b.y), c2 = __popc(b.z), c3 = __popc(b.w);
        unsigned long long tot;
        const unsigned int ex = (unsigned int) (prefix + block_exclusive_scan<unsigned long long>((unsigned long long) (c0 + c1 + c2 + c3), smem, tot));
        if (w < words) *reinterpret_cast<uint4 *>(F.wpre + w) = make_uint4(ex, ex + c0, ex + c0 + c1, ex + c0 + c1 + c2);
    }
}

// P4: voxel (= rank of the cell's bit) and arrival ticket of every point
template <int W>
__device__ void vg_rank(const FusedArgs &F) {
    VGPass<W> V(F, true);
    if (V.pass || V.dead) return;
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    for (unsigned int i0 = gt; i0 < V.n; i0 += 4 * gn) {
        unsigned int cell[4], pre[4], bw[4], r[4], tk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned int i = i0 + k * gn; cell[k] = i < V.n ? F.vkey[i] : kPad; }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (cell[k] != kPad) { pre[k] = F.wpre[cell[k] >> 5]; bw[k] = F.bits[cell[k] >> 5]; }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (cell[k] != kPad) {
                r[k] = pre[k] + (unsigned int) __popc(bw[k] & ((1u << (cell[k] & 31)) - 1u));
                tk[k] = atomicAdd(&F.vcnt[r[k]], 1u);
            }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (cell[k] != kPad) { const unsigned int i = i0 + k * gn; F.vkey[i] = r[k]; F.varr[i] = tk[k]; }
    }
}

template <int W>
__device__ __forceinline__ unsigned int vg_voxels(const FusedArgs &F) { return W == 0 ? F.c->n_ds_hits : F.c->n_frees; }

// P5 / P6: vstart = exclusive scan of the voxel counts (four voxels per thread)
template <int W>
__device__ void vg_span_count(const FusedArgs &F, unsigned long long *smem) {
    VGPass<W> V(F, false);
    const unsigned int nv = (V.pass || V.dead) ? 0u : min(vg_voxels<W>(F), F.vcnt_cap - 8u);
    const unsigned int n_tiles = (nv + kWT - 1) / kWT;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned int r = t * kWT + 4 * threadIdx.x;
        uint4 c = r < nv ? *reinterpret_cast<const uint4 *>(F.vcnt + r) : make_uint4(0u, 0u, 0u, 0u);
        if (r + 1 >= nv) c.y = 0; if (r + 2 >= nv) c.z = 0; if (r + 3 >= nv) c.w = 0;
        const unsigned long long s = block_sum<unsigned long long>((unsigned long long) c.x + c.y + c.z + c.w, smem);
        if (threadIdx.x == 0) F.tsum[t] = s;
    }
}
template <int W>
__device__ void vg_span_place(const FusedArgs &F, unsigned long long *smem) {
    VGPass<W> V(F, false);
    const unsigned int nv = (V.pass || V.dead) ? 0u : min(vg_voxels<W>(F), F.vcnt_cap - 8u);
    const unsigned int n_tiles = (nv + kWT - 1) / kWT;
    TileCarry<unsigned long long> carry;
    bool ok = true;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned int r = t * kWT + 4 * threadIdx.x;
        uint4 c = r < nv ? *reinterpret_cast<const uint4 *>(F.vcnt + r) : make_uint4(0u, 0u, 0u, 0u);
        if (r + 1 >= nv) c.y = 0; if (r + 2 >= nv) c.z = 0; if (r + 3 >= nv) c.w = 0;
        const unsigned long long prefix = carry.prefix(F.tsum, t, smem);
        unsigned long long tot;
        const unsigned int ex = (unsigned int) (prefix + block_exclusive_scan<unsigned long long>((unsigned long long) c.x + c.y + c.z + c.w, smem, tot));
        if (r < nv) {
            // (the sentinel lands at [nv]: counts past nv read as zero)
            const unsigned int s1 = ex + c.x, s2 = s1 + c.y, s3 = s2 + c.z, s4 = s3 + c.w;
            F.vstart[r] = ex;
            if (r + 1 <= nv) F.vstart[r + 1] = s1;
            if (r + 2 <= nv) F.vstart[r + 2] = s2;
            if (r + 3 <= nv) F.vstart[r + 3] = s3;
            if (r + 4 == nv) F.vstart[r + 4] = s4;
            ok = note_long_span(F, W, r, c.x) && ok;
            ok = note_long_span(F, W, r + 1, c.y) && ok;
            ok = note_long_span(F, W, r + 2, c.z) && ok;
            ok = note_long_span(F, W, r + 3, c.w) && ok;
        }
    }
    if (!ok) atomicOr(&F.c->overflow, OVF_FAST);
}

// P7: indices into the voxel's span, in arrival order
template <int W>
__device__ void vg_drop(const FusedArgs &F) {
    VGPass<W> V(F, false);
    if (V.pass || V.dead) return;
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    for (unsigned int i0 = gt; i0 < V.n; i0 += 4 * gn) {
        unsigned int r[4], tk[4], st[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int i = i0 + k * gn;
            r[k] = kPad;
            if (i < V.n) { r[k] = F.vkey[i]; tk[k] = F.varr[i]; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (r[k] != kPad) st[k] = F.vstart[r[k]];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (r[k] != kPad) F.vlist[st[k] + tk[k]] = i0 + k * gn;
    }
}

// Spans longer than kShortRun.  List A (up to kWideRun indices): a warp per span, the span in registers (four indices per
// lane), every index counts the smaller ones through shuffles.  List B (up to kFastRun): a CTA per span, the span in shared
// memory, a thread per index.  emit(position, index, span) for every index.  The lists hand consecutive (= neighbouring,
// = similarly long) spans to different warps / CTAs.
template <typename Emit>
__device__ inline void order_long_spans(const FusedArgs &F, int pass, const unsigned int *start, const unsigned int *list,
                                        unsigned int *s_list /* [kFastRun + 4] */, Emit emit) {
    const int lane = threadIdx.x & 31;
    const unsigned int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const unsigned int nA = min(F.c->fz_nlong[pass][0], F.long_capA), nB = min(F.c->fz_nlong[pass][1], F.long_capB);
    for (unsigned int q = blockIdx.x; q < nB; q += gridDim.x) {
        const unsigned int r = F.longs[F.long_capA + q];
        const unsigned int f = start[r], l = min(start[r + 1] - f, kFastRun);
        __syncthreads();
        for (unsigned int j = threadIdx.x; j < ((l + 3u) & ~3u); j += kFT) s_list[j] = j < l ? list[f + j] : kPad;
        __syncthreads();
        for (unsigned int j0 = threadIdx.x; j0 < l; j0 += kFT) {
            const unsigned int x = s_list[j0];
            unsigned int rank = 0;
            for (unsigned int j = 0; j < l; j += 4) {
                const uint4 y = *reinterpret_cast<const uint4 *>(s_list + j);
                rank += (y.x < x ? 1u : 0u) + (y.y < x ? 1u : 0u) + (y.z < x ? 1u : 0u) + (y.w < x ? 1u : 0u);
            }
            emit(f + rank, x, r);
        }
    }
    constexpr int kPer = kWideRun / 32;
    for (unsigned int q = gw; q < nA; q += nw) {
        const unsigned int r = F.longs[q];
        const unsigned int f = start[r], l = min(start[r + 1] - f, kWideRun);
        unsigned int v[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) v[k] = lane + 32u * k < l ? list[f + lane + 32u * k] : kPad;
#pragma unroll
        for (int kx = 0; kx < kPer; ++kx) {
            if (32u * kx >= l) break;
            const unsigned int x = v[kx];
            unsigned int rank = 0;
#pragma unroll
            for (int ky = 0; ky < kPer; ++ky) {
                if (32u * ky >= l) break;
#pragma unroll
                for (int sft = 0; sft < 32; ++sft) rank += __shfl_sync(0xffffffffu, v[ky], sft) < x ? 1u : 0u;
            }
            if (lane + 32u * kx < l) emit(f + rank, x, r);
        }
    }
}

// P8: ... and into input order: the place of an index = how many indices of the span are smaller
template <int W>
__device__ void vg_order(const FusedArgs &F, unsigned char *scratch) {
    VGPass<W> V(F, false);
    if (V.pass || V.dead) return;
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    unsigned int *vsorted = F.vsorted;
    order_long_spans(F, W, F.vstart, F.vlist, reinterpret_cast<unsigned int *>(scratch), [vsorted](unsigned int pos, unsigned int idx, unsigned int) { vsorted[pos] = idx; });
    for (unsigned int i0 = gt; i0 < V.n; i0 += 4 * gn) {
        unsigned int r[4], first[4], last[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned int i = i0 + k * gn; r[k] = i < V.n ? F.vkey[i] : kPad; }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (r[k] != kPad) { first[k] = F.vstart[r[k]]; last[k] = F.vstart[r[k] + 1]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (r[k] == kPad || last[k] - first[k] > kShortRun) continue;
            const unsigned int i = i0 + k * gn;
            unsigned int rank = 0;
            for (unsigned int li = first[k]; li < last[k]; ++li) rank += F.vlist[li] < i ? 1u : 0u;
            F.vsorted[first[k] + rank] = i;
        }
    }
}

// acc + v[0] + v[1] + ... one by one (pcl's CentroidPoint); v is 16-byte aligned shared memory
__device__ __forceinline__ float seq_sum(float acc, const float *v, unsigned int m) {
    unsigned int j = 0;
    for (; j + 8 <= m; j += 8) {
        const float4 a = *reinterpret_cast<const float4 *>(v + j), b = *reinterpret_cast<const float4 *>(v + j + 4);
        acc += a.x; acc += a.y; acc += a.z; acc += a.w;
        acc += b.x; acc += b.y; acc += b.z; acc += b.w;
    }
    for (; j < m; ++j) acc += v[j];
    return acc;
}

// per downsampled hit: 0 if the range filter drops it (bgkoctomap.cpp:394-398), else 1 + number of free points it emits
// (origin once per kept hit :404, samples d = fr, 2fr.. < l with fp32 accumulation :451-455, tail sample :456-457)
__device__ inline unsigned int hit_free_count(const float4 h, const ScanArgs *A) {
    const float dx = h.x - A->ox, dy = h.y - A->oy, dz = h.z - A->oz;
    const float s = dx * dx + dy * dy + dz * dz;
    if (A->max_range > 0) {
        const double l = sqrt((double) s);                 // point3f::norm() (point3f.h:207-214)
        if (l > (double) A->max_range) return 0u;
    }
    const float l = (float) sqrt((double) s);
    const float fr = A->fr;
    const float *__restrict__ tab = A->beam_tab;
    const unsigned int tab_n = A->beam_tab_n;
    const float est = l / fr;
    unsigned int lo = est < (float) tab_n ? (unsigned int) est : tab_n;
    while (lo < tab_n && tab[lo] < l) ++lo;
    while (lo > 0 && !(tab[lo - 1] < l)) --lo;
    unsigned int cnt = 1 + lo;                              // the origin + regular samples
    if (lo == A->beam_tab_n) {
        float d = tab[lo - 1] + fr;
        while (d < l) { ++cnt; const float nd = d + fr; if (nd == d) break; d = nd; }
    }
    if (l > fr) ++cnt;
    return cnt;
}

// P9 / P19: centroids.  A thread per voxel of up to kShortRun points (tiles of 1024 voxels), a warp per longer one (from
// the lists of long spans).  W = 0 also computes the free-point count of every downsampled hit, adds it to the sums per
// kBeamTile hits of the beam phase, and clears the bitmap and the counts for the second pass; W = 1 accumulates the
// bounding box of the free centroids.
template <int W>
__device__ __forceinline__ void centroid_done(const FusedArgs &F, unsigned int r, float4 v, Box &box) {
    if (W == 0) {
        const ScanArgs *A = F.A;
        const unsigned int cnt = hit_free_count(v, A);
        F.hit_cnt[r] = cnt;
        if (cnt) {
            atomicAdd(&F.bsum[r / kBeamTile], (1ull << 32) | (unsigned long long) cnt);
            // Bounding box of this beam's free points (getMinMax3D of the second voxel grid) without generating them:
            // every sample is o + n * d, rounded, and rounding is monotonic in d, so on every axis the extremes are at the
            // origin, the nearest and the farthest sample (same expressions as beam_fill).
            const float ox = A->ox, oy = A->oy, oz = A->oz, fr = A->fr;
            const float dx = v.x - ox, dy = v.y - oy, dz = v.z - oz;
            const float l = (float) sqrt((double) (dx * dx + dy * dy + dz * dz));
            const float nx = dx / l, ny = dy / l, nz = dz / l;
            const unsigned int tail = l > fr ? 1u : 0u, n_reg = cnt - 1u - tail;
            box.add(ox, oy, oz);
            if (n_reg) {
                const float d0 = A->beam_tab[0];
                const unsigned int e = n_reg - 1u;
                const float d1 = e < A->beam_tab_n ? A->beam_tab[e] : add_repeat(fr, fr, e);
                box.add(ox + nx * d0, oy + ny * d0, oz + nz * d0);
                box.add(ox + nx * d1, oy + ny * d1, oz + nz * d1);
            }
            if (tail) { const float d = l - fr; box.add(ox + nx * d, oy + ny * d, oz + nz * d); }
        }
    } else box.add(v.x, v.y, v.z);
}

template <int W>
__device__ void vg_centroid(const FusedArgs &F, unsigned long long *smem, unsigned char *scratch) {
    VGPass<W> V(F, true);
    const ScanArgs *A = F.A;
    const unsigned int nv = V.dead ? 0u : min(vg_voxels<W>(F), W == 0 ? F.points_cap : F.raw_cap);
    const unsigned int n_hits = W == 1 ? F.c->n_hits : 0u;
    const unsigned int off = W == 1 ? n_hits : 0u;
    float4 *out = W == 0 ? F.hits_ds : F.xy;
    const float label = W == 0 ? 1.0f : A->free_label;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    float(*wstage)[3][kMidStage] = reinterpret_cast<float(*)[3][kMidStage]>(scratch);   // [32][3][64]
    __shared__ unsigned int s_mm[6];
    (void) smem;
    // the origin's voxel (W = 1)
    unsigned int r0 = kPad;
    if (W == 1 && !V.pass && nv > 0 && n_hits > 0) {
        const unsigned int cell = vg_cell(V.f, V.inv, A->ox, A->oy, A->oz);
        r0 = F.wpre[cell >> 5] + (unsigned int) __popc(F.bits[cell >> 5] & ((1u << (cell & 31)) - 1u));
    }
    if (W == 1 && blockIdx.x == 0 && threadIdx.x == 0 && !V.dead) F.c->n_train = n_hits + nv;
    Box box;
    // ---- long voxels first (they are the long poles).  List B: a CTA each, all points gathered into shared memory at
    // once, three threads add them up; list A: a warp each, staged kMidStage points at a time.
    if (!V.pass) {
        const unsigned int gw = gt >> 5, nw = gn >> 5;
        const unsigned int nA = min(F.c->fz_nlong[W][0], F.long_capA), nB = min(F.c->fz_nlong[W][1], F.long_capB);
        float(*cstage)[kFastRun] = reinterpret_cast<float(*)[kFastRun]>(scratch);     // [3][2048]
        for (unsigned int q = blockIdx.x; q < nB; q += gridDim.x) {
            const unsigned int r2 = F.longs[F.long_capA + q];
            if (W == 1 && r2 == r0) continue;
            const unsigned int run_first = F.vstart[r2], l = min(F.vstart[r2 + 1] - run_first, kFastRun);
            __syncthreads();
            for (unsigned int j = threadIdx.x; j < l; j += kFT) {
                const float *p = V.in + (size_t) F.vsorted[run_first + j] * V.stride;
                cstage[0][j] = p[0]; cstage[1][j] = p[1]; cstage[2][j] = p[2];
            }
            __syncthreads();
            if (warp == 0) {
                float acc = 0.f;
                if (lane < 3) acc = seq_sum(0.f, cstage[lane], l) / (float) l;
                const float cx = __shfl_sync(0xffffffffu, acc, 0), cy = __shfl_sync(0xffffffffu, acc, 1),
                            cz = __shfl_sync(0xffffffffu, acc, 2);
                if (lane == 0) {
                    const float4 v = make_float4(cx, cy, cz, label);
                    out[off + r2] = v;
                    centroid_done<W>(F, r2, v, box);
                }
            }
        }
        __syncthreads();
        for (unsigned int q = gw; q < nA; q += nw) {
            const unsigned int r2 = F.longs[q];
            if (W == 1 && r2 == r0) continue;
            const unsigned int run_first = F.vstart[r2], run_last = F.vstart[r2 + 1];
            float acc = 0.f;
            float px[kMidStage / 32], py[kMidStage / 32], pz[kMidStage / 32];
            auto fetch = [&](unsigned int s0) {
                const unsigned int mcount = min((unsigned int) kMidStage, run_last - s0);
#pragma unroll
                for (int k = 0; k < kMidStage / 32; ++k) {
                    const unsigned int j = lane + 32 * k;
                    if (j < mcount) {
                        const float *p = V.in + (size_t) F.vsorted[s0 + j] * V.stride;
                        px[k] = p[0]; py[k] = p[1]; pz[k] = p[2];
                    }
                }
            };
            fetch(run_first);
            for (unsigned int s0 = run_first; s0 < run_last; s0 += kMidStage) {
                const unsigned int mcount = min((unsigned int) kMidStage, run_last - s0);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < kMidStage / 32; ++k) {
                    const unsigned int j = lane + 32 * k;
                    if (j < mcount) { wstage[warp][0][j] = px[k]; wstage[warp][1][j] = py[k]; wstage[warp][2][j] = pz[k]; }
                }
                __syncwarp();
                if (s0 + kMidStage < run_last) fetch(s0 + kMidStage);     // in flight while three lanes add this chunk
                if (lane < 3) acc = seq_sum(acc, wstage[warp][lane], mcount);
            }
            acc = acc / (float) (run_last - run_first);
            const float cx = __shfl_sync(0xffffffffu, acc, 0), cy = __shfl_sync(0xffffffffu, acc, 1),
                        cz = __shfl_sync(0xffffffffu, acc, 2);
            if (lane == 0) {
                const float4 v = make_float4(cx, cy, cz, label);
                out[off + r2] = v;
                centroid_done<W>(F, r2, v, box);
            }
        }
    }
    // ---- everything else: a thread per voxel, four points in flight
    for (unsigned int r = gt; r < nv; r += gn) {
        float4 v;
        if (V.pass) {
            const float *p = V.in + (size_t) r * V.stride;
            v = make_float4((0.f + p[0]) / 1.0f, (0.f + p[1]) / 1.0f, (0.f + p[2]) / 1.0f, label);
        } else {
            const unsigned int first = F.vstart[r], last = F.vstart[r + 1];
            if (W == 1 && r == r0) {
                // origin copies interleaved with the other samples of this voxel, in push order
                float acc[3] = {0.f, 0.f, 0.f};
                const float o[3] = {A->ox, A->oy, A->oz};
                unsigned int done = 0;
                for (unsigned int li = first; li < last; ++li) {
                    const float4 w4 = F.frees_raw[F.vsorted[li]];
                    const unsigned int k = (unsigned int) __float_as_int(w4.w) + 1u;   // origins pushed before this sample
                    const float pv[3] = {w4.x, w4.y, w4.z};
#pragma unroll
                    for (int a = 0; a < 3; ++a) { acc[a] = add_repeat(acc[a], o[a], k - done); acc[a] += pv[a]; }
                    done = k;
                }
                const float cnt = (float) (n_hits + (last - first));
#pragma unroll
                for (int a = 0; a < 3; ++a) acc[a] = add_repeat(acc[a], o[a], n_hits - done) / cnt;
                v = make_float4(acc[0], acc[1], acc[2], label);
            } else if (last - first > kShortRun) {
                continue;                                   // a warp's job (above)
            } else {
                float sx = 0.f, sy = 0.f, sz = 0.f;
                for (unsigned int l0 = first; l0 < last; l0 += 4) {
                    unsigned int idx[4];
                    float px[4], py[4], pz[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (l0 + k < last) idx[k] = F.vsorted[l0 + k];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (l0 + k < last) {
                            const float *p = V.in + (size_t) idx[k] * V.stride;
                            px[k] = p[0]; py[k] = p[1]; pz[k] = p[2];
                        }
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (l0 + k < last) { sx += px[k]; sy += py[k]; sz += pz[k]; }
                }
                const float cnt = (float) (last - first);
                v = make_float4(sx / cnt, sy / cnt, sz / cnt, label);
            }
        }
        out[off + r] = v;
        centroid_done<W>(F, r, v, box);
    }
    box.flush(F.mm + (W == 1 ? 12 : 6), s_mm);
    if (W == 0 && !V.pass) {
        // second pass starts from a clear bitmap and clear counts
        zero_words(F.bits, min(V.words(), F.bits_words));
        zero_words(F.vcnt, min((size_t) nv + 1, (size_t) F.vcnt_cap));
    }
}

// P10: kept hits -> xy[0 .. n_hits), all free points -> frees_raw in the reference's push order (origin :404, samples
// :451-457); w of a free point = ordinal of its hit among the kept ones (an origin copy: -1).  Tiles of kBeamTile hits,
// the free points of a tile are written by all threads (coalesced 16-byte stores), each finding its hit by bisection.
__device__ void beam_fill(const FusedArgs &F, unsigned long long *smem, unsigned char *scratch) {
    const ScanArgs *A = F.A;
    ScanCounters *c = F.c;
    const bool dead = ld_volatile(&c->overflow) != 0u;
    const unsigned int n = dead ? 0u : min(c->n_ds_hits, F.points_cap);
    const unsigned int n_tiles = min((n + kBeamTile - 1) / kBeamTile, F.bsum_n);
    unsigned int *s_off = reinterpret_cast<unsigned int *>(scratch);             // [kBeamTile + 1]
    unsigned int *s_cnt = s_off + kBeamTile + 32;
    unsigned int *s_ord = s_cnt + kBeamTile;
    float4 *s_beam = reinterpret_cast<float4 *>(scratch + 4096);
    __shared__ unsigned int s_mm[6];
    const unsigned long long total = block_sum_range<unsigned long long>(F.bsum, 0, n_tiles, smem);
    const unsigned int n_hits = (unsigned int) (total >> 32), n_raw = (unsigned int) (total & 0xFFFFFFFFull);
    if (blockIdx.x == 0 && threadIdx.x == 0 && !dead) {
        c->n_hits = n_hits;
        c->n_raw_frees = n_raw;
        if (n_raw > F.raw_cap) atomicOr(&c->overflow, OVF_RAW);
    }
    if (n_raw > F.raw_cap) return;
    const float ox = A->ox, oy = A->oy, oz = A->oz, fr = A->fr;
    // the second voxel grid's frame is known (bounding box from vg_centroid<0>): cell of every sample, one bit per cell
    const float inv = A->inv_ds;
    VGFrame f2;
    f2.passthrough = true; f2.cells = 0;
    if (!(A->ds < 0) && n_raw > 0) f2 = vg_frame(F.mm + 6, inv);
    const bool pass2 = A->ds < 0 || f2.passthrough;
    const bool cells_ok = (unsigned long long) f2.cells <= (unsigned long long) F.vg_cells_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0 && !dead && n_raw > 0) {
        c->vg_passthrough[1] = pass2 ? 1u : 0u;
        if (!pass2 && !cells_ok) {
            atomicOr(&c->overflow, OVF_VGCELLS);
            atomicMax(&c->vg_cells_needed, (unsigned int) (f2.cells > 0x80000000ll ? 0x80000000ll : f2.cells));
        }
        // the sensor origin's voxel exists whatever else falls into it (the origin copies are not listed)
        if (!pass2 && cells_ok) set_bit(F.bits, vg_cell(f2, inv, ox, oy, oz));
    }
    const bool keys2 = !pass2 && cells_ok;
    const unsigned int tab_n = A->beam_tab_n;
    const float *__restrict__ tab = A->beam_tab;
    Box hbox;
    TileCarry<unsigned long long> carry;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned int i = t * kBeamTile + threadIdx.x;
        const bool mine_hit = threadIdx.x < kBeamTile && i < n;
        const unsigned int cnt = mine_hit ? F.hit_cnt[i] : 0u;
        float4 hit = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cnt) hit = F.hits_ds[i];
        const unsigned long long prefix = carry.prefix(F.bsum, t, smem);
        unsigned long long cta_total;
        const unsigned long long mine = cnt ? ((1ull << 32) | (unsigned long long) cnt) : 0ull;
        const unsigned long long excl = block_exclusive_scan<unsigned long long>(mine, smem, cta_total);
        if (threadIdx.x <= kBeamTile) s_off[threadIdx.x] = (unsigned int) (excl & 0xFFFFFFFFull);
        if (threadIdx.x < kBeamTile) s_cnt[threadIdx.x] = cnt;
        if (cnt) {
            const unsigned int ord = (unsigned int) ((prefix + excl) >> 32);
            F.xy[ord] = make_float4(hit.x, hit.y, hit.z, 1.0f);
            hbox.add(hit.x, hit.y, hit.z);
            const float dx = hit.x - ox, dy = hit.y - oy, dz = hit.z - oz;
            const float l = (float) sqrt((double) (dx * dx + dy * dy + dz * dz));     // beam_sample's preamble (:437-449)
            s_beam[threadIdx.x] = make_float4(dx / l, dy / l, dz / l, l);
            s_ord[threadIdx.x] = ord;
        }
        __syncthreads();
        const unsigned int tile_total = (unsigned int) (cta_total & 0xFFFFFFFFull);
        float4 *out = F.frees_raw + (unsigned int) (prefix & 0xFFFFFFFFull);
        unsigned int *kout = F.vkey + (unsigned int) (prefix & 0xFFFFFFFFull);
        for (unsigned int q = threadIdx.x; q < tile_total; q += kFT) {
            // h = the last hit of the tile whose first free point is at or before q
            unsigned int lo = 0, hi = kBeamTile;
            while (hi - lo > 1) { const unsigned int mid = (lo + hi) >> 1; if (s_off[mid] <= q) lo = mid; else hi = mid; }
            const unsigned int h = lo, e1 = q - s_off[h], ch = s_cnt[h];
            float sx = ox, sy = oy, sz = oz;
            int w = -1;
            if (e1) {
                const float4 bm = s_beam[h];
                const unsigned int tail = bm.w > fr ? 1u : 0u;
                const unsigned int n_reg = ch - 1u - tail;                                    // samples with d < l (:451-455)
                const unsigned int e = e1 - 1u;
                const float d = e < n_reg ? (e < tab_n ? tab[e] : add_repeat(fr, fr, e)) : bm.w - fr;   // :453 | :457
                sx = ox + bm.x * d; sy = oy + bm.y * d; sz = oz + bm.z * d;
                w = (int) s_ord[h];
            }
            out[q] = make_float4(sx, sy, sz, __int_as_float(w));
            if (keys2) {
                unsigned int cell = kPad;                                  // (an origin copy stays out of the lists)
                if (e1) { cell = vg_cell(f2, inv, sx, sy, sz); atomicOr(&F.bits[cell >> 5], 1u << (cell & 31)); }
                kout[q] = cell;
            }
        }
        __syncthreads();
    }
    hbox.flush(F.mm + 12, s_mm);
}

// ---- block grid of the scan (get_blocks_in_bbox, bgkoctomap.cpp:486-495): CTA 0, see k_grid in binning.cu -------------
constexpr int kGridChunk = 2048;

__device__ void block_grid(const FusedArgs &F, unsigned char *scratch) {
    GridDesc *g = F.g;
    ScanCounters *c = F.c;
    const DevParams *P = F.P;
    const unsigned int *mm = F.mm + 12;
    float(*xs)[kGridChunk] = reinterpret_cast<float(*)[kGridChunk]>(scratch);      // [3][2048]
    __shared__ float s_x[3];
    __shared__ long long s_last[3];
    __shared__ int s_cnt[3], s_steps[3], s_done[3], s_bad, s_irr;
    {   // *g starts zeroed
        unsigned int *gw = reinterpret_cast<unsigned int *>(g);
        for (unsigned int i = threadIdx.x; i < sizeof(GridDesc) / 4; i += blockDim.x) gw[i] = 0u;
    }
    __syncthreads();
    const bool live = ld_volatile(&c->overflow) == 0u && c->n_train > 0;
    const bool worker = threadIdx.x < 384;
    const int a = worker ? (int) (threadIdx.x >> 7) : 0, t = worker ? (int) (threadIdx.x & 127) : 1;
    const float bs = P->block_size;
    const float mn = float_unflip(mm[a]), mx = float_unflip(mm[3 + a]);
    const float hi = mx + 2 * bs;
    const long long first = axis_index(mn - bs, bs);
    if (threadIdx.x == 0) { s_bad = 0; s_irr = 0; }
    if (worker && t == 0) { s_x[a] = mn - bs; s_steps[a] = 0; s_done[a] = live ? 0 : 1; s_last[a] = first - 1; }
    __syncthreads();
    while (!(s_done[0] && s_done[1] && s_done[2]) && !s_bad) {
        if (worker && t == 0) {
            int n = 0;
            if (!s_done[a]) {
                float x = s_x[a];
                while (n < kGridChunk && x <= hi) { xs[a][n++] = x; x += bs; }
                s_x[a] = x;
            }
            s_cnt[a] = n;
        }
        __syncthreads();
        const int n = s_cnt[a], base = s_steps[a];
        const long long last_before = s_last[a];
        __syncthreads();
        if (worker) {
            for (int j = t; j < n; j += 128) {
                const long long idx = axis_index(xs[a][j], bs);
                const long long prev = j ? axis_index(xs[a][j - 1], bs) : last_before;
                if (base + j > 0 && idx != prev + 1) s_irr = 1;
                const long long rel = idx - first;
                if (rel < 0 || rel >= kMaxAxis || base + j >= kMaxAxis) s_bad = 1;
                else g->present[a][rel] = 1;
                if (j == n - 1) s_last[a] = idx;
            }
            if (t == 0) {
                s_steps[a] = base + n;
                if (n < kGridChunk) s_done[a] = 1;          // the stepping passed max + 2 block_size
            }
        }
        __syncthreads();
    }
    if (live && worker && t == 0) {
        g->base[a] = first;
        g->n[a] = s_steps[a] == 0 ? 0 : (int) (s_last[a] - first + 1);
    }
    __syncthreads();
    if (live && threadIdx.x == 0) {
        if (s_irr) g->irregular = 1;
        const unsigned long long cells = (unsigned long long) g->n[0] * (unsigned long long) g->n[1] *
                                         (unsigned long long) g->n[2];
        if (s_bad || cells >= 0x7FFFFFF0ull) atomicOr(&c->overflow, OVF_EXTENT);
        else {
            g->n_cells = (unsigned int) cells;
            c->n_cells = (unsigned int) cells;
            c->grid_irregular = (unsigned int) g->irregular;
            if (cells > (unsigned long long) F.cells_cap) atomicOr(&c->overflow, OVF_CELLS);
        }
    }
}

// candidate block indices of one coordinate (closed box, bgkoctomap.cpp:497-503, rtree.h:1519-1532): see binning.cu
__device__ inline unsigned int axis_candidates(float q, float bs, float half, const GridDesc *g, int a, int &rel0) {
    const long long i0 = axis_index(q, bs);
    rel0 = (int) (i0 - g->base[a]);
    unsigned int mask = 0;
#pragma unroll
    for (int k = -1; k <= 1; ++k) {
        const long long ii = i0 + k;
        const float c = axis_center(ii, bs);
        const float lo = c - half, hi = c + half;
        if (lo > q || q > hi) continue;
        const long long rel = ii - g->base[a];
        if (rel < 0 || rel >= g->n[a] || !g->present[a][rel]) continue;   // block not enumerated this scan
        mask |= 1u << (k + 1);
    }
    return mask;
}

// B1: memberships of every training entry; a ticket per (cell, entry).  Two entries in flight per thread.
__device__ void bin_members(const FusedArgs &F) {
    ScanCounters *c = F.c;
    const GridDesc *g = F.g;
    const bool dead = ld_volatile(&c->overflow) != 0u;
    const unsigned int n = dead ? 0u : min(c->n_train, F.train_cap);
    const float bs = F.P->block_size, half = F.P->half_size;
    const unsigned int n1 = (unsigned int) g->n[1], n2 = (unsigned int) g->n[2];
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    bool ovf = false;
    for (unsigned int i0 = gt; i0 < n; i0 += 4 * gn) {
        float4 p[4];
        unsigned int mx[4], my[4], mz[4], cell0[4], tk[4];
        int rx[4], ry[4], rz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned int i = i0 + k * gn; if (i < n) p[k] = F.xy[i]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int i = i0 + k * gn;
            mx[k] = my[k] = mz[k] = 0;
            if (i >= n) continue;
            mx[k] = axis_candidates(p[k].x, bs, half, g, 0, rx[k]);
            my[k] = axis_candidates(p[k].y, bs, half, g, 1, ry[k]);
            mz[k] = axis_candidates(p[k].z, bs, half, g, 2, rz[k]);
        }
        // the common case first: one block per entry, both tickets in flight together
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cell0[k] = kPad;
            if (!(mx[k] && my[k] && mz[k])) continue;
            const int a = __ffs(mx[k]) - 1, b = __ffs(my[k]) - 1, d = __ffs(mz[k]) - 1;
            cell0[k] = ((unsigned int) (rx[k] + a - 1) * n1 + (unsigned int) (ry[k] + b - 1)) * n2 + (unsigned int) (rz[k] + d - 1);
            tk[k] = atomicAdd(&F.cell_cnt[cell0[k]], 1u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int i = i0 + k * gn;
            if (i >= n) continue;
            F.mcell[i] = cell0[k];
            if (cell0[k] == kPad) continue;
            F.mk[i] = tk[k];
            if (__popc(mx[k]) * __popc(my[k]) * __popc(mz[k]) == 1) continue;
            // a point on a block boundary: its other blocks, in the order of the closed-box enumeration
            bool first = true;
            for (int a = 0; a < 3; ++a) {
                if (!(mx[k] & (1u << a))) continue;
                for (int b = 0; b < 3; ++b) {
                    if (!(my[k] & (1u << b))) continue;
                    for (int d = 0; d < 3; ++d) {
                        if (!(mz[k] & (1u << d))) continue;
                        if (first) { first = false; continue; }
                        const unsigned int cell = ((unsigned int) (rx[k] + a - 1) * n1 + (unsigned int) (ry[k] + b - 1)) * n2 +
                                                  (unsigned int) (rz[k] + d - 1);
                        const unsigned int t2 = atomicAdd(&F.cell_cnt[cell], 1u);
                        const unsigned int e = atomicAdd(&c->n_extra, 1u);
                        if (e < kMaxExtra) F.extra[e] = Extra{i, cell, t2};
                        else ovf = true;
                    }
                }
            }
        }
    }
    if (ovf) atomicOr(&c->overflow, OVF_FAST);
}

// B2 / B3: scan over the dense cells, four per thread: (data blocks, memberships) before each cell
__device__ __forceinline__ unsigned long long cell_item(unsigned int cnt) {
    return cnt ? ((1ull << 32) | (unsigned long long) cnt) : 0ull;
}
__device__ __forceinline__ uint4 load_cells4(const unsigned int *cell_cnt, unsigned int id, unsigned int nc) {
    uint4 c = id < nc ? *reinterpret_cast<const uint4 *>(cell_cnt + id) : make_uint4(0u, 0u, 0u, 0u);
    if (id + 1 >= nc) c.y = 0;
    if (id + 2 >= nc) c.z = 0;
    if (id + 3 >= nc) c.w = 0;
    return c;
}

__device__ void bin_cell_count(const FusedArgs &F, unsigned long long *smem) {
    const bool dead = ld_volatile(&F.c->overflow) != 0u;
    const unsigned int nc = dead ? 0u : min(F.g->n_cells, F.cells_cap);
    const unsigned int n_tiles = (nc + kWT - 1) / kWT;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint4 cn = load_cells4(F.cell_cnt, t * kWT + 4 * threadIdx.x, nc);
        const unsigned long long s = block_sum<unsigned long long>(cell_item(cn.x) + cell_item(cn.y) + cell_item(cn.z) + cell_item(cn.w), smem);
        if (threadIdx.x == 0) F.tsum[t] = s;
    }
}

__device__ void bin_cell_place(const FusedArgs &F, unsigned long long *smem) {
    ScanCounters *c = F.c;
    const GridDesc *g = F.g;
    const bool dead = ld_volatile(&c->overflow) != 0u;
    const unsigned int nc = dead ? 0u : min(g->n_cells, F.cells_cap);
    const unsigned int n_tiles = (nc + kWT - 1) / kWT;
    const unsigned long long total = block_sum_range<unsigned long long>(F.tsum, 0, n_tiles, smem);
    const unsigned int n_db = (unsigned int) (total >> 32), n_mem = (unsigned int) (total & 0xFFFFFFFFull);
    if (blockIdx.x == 0 && threadIdx.x == 0 && !dead) {
        c->n_members = n_mem;
        c->n_data_blocks = n_db;
        if (n_mem > F.members_cap) atomicOr(&c->overflow, OVF_MEMBERS);
        else F.db_start[n_db] = n_mem;
    }
    if (n_mem > F.members_cap) return;
    TileCarry<unsigned long long> carry;
    bool ok = true;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned int id0 = t * kWT + 4 * threadIdx.x;
        const uint4 cn = load_cells4(F.cell_cnt, id0, nc);
        const unsigned long long prefix = carry.prefix(F.tsum, t, smem);
        const unsigned int cnt[4] = {cn.x, cn.y, cn.z, cn.w};
        unsigned long long tot;
        unsigned long long ex = prefix + block_exclusive_scan<unsigned long long>(
                                             cell_item(cn.x) + cell_item(cn.y) + cell_item(cn.z) + cell_item(cn.w), smem, tot);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!cnt[j]) continue;
            const unsigned int pos = (unsigned int) (ex >> 32);
            F.db_id[pos] = id0 + j;
            F.db_start[pos] = (unsigned int) (ex & 0xFFFFFFFFull);
            F.cell_db[id0 + j] = pos + 1;
            ok = note_long_span(F, 2, pos, cnt[j]) && ok;
            ex += cell_item(cnt[j]);
        }
    }
    if (!ok) atomicOr(&c->overflow, OVF_FAST);
}

// B4: entries into their block's span in arrival order; one bit for each of the 7 blocks whose ExtendedBlock holds a data
// block (itself and its 6 face neighbours, bgkblock.cpp:85-101) = the test-block candidates
__device__ void bin_drop(const FusedArgs &F) {
    ScanCounters *c = F.c;
    const GridDesc *g = F.g;
    const bool dead = ld_volatile(&c->overflow) != 0u;
    const unsigned int n = dead ? 0u : min(c->n_train, F.train_cap);
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    for (unsigned int i0 = gt; i0 < n; i0 += 4 * gn) {
        unsigned int cell[4], tk[4], d[4], st[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int i = i0 + k * gn;
            cell[k] = kPad;
            if (i < n) { cell[k] = F.mcell[i]; tk[k] = F.mk[i]; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (cell[k] != kPad) d[k] = F.cell_db[cell[k]] - 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (cell[k] != kPad) st[k] = F.db_start[d[k]];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (cell[k] != kPad) F.mlist[st[k] + tk[k]] = i0 + k * gn;
    }
    const unsigned int ne = dead ? 0u : min(c->n_extra, kMaxExtra);
    for (unsigned int e = gt; e < ne; e += gn) {
        const Extra x = F.extra[e];
        F.mlist[F.db_start[F.cell_db[x.cell] - 1] + x.k] = x.entry;
    }
    const unsigned int n_db = dead ? 0u : min(c->n_data_blocks, F.members_cap);
    const int nz = g->n[2], ny = g->n[1], nx = g->n[0];
    for (unsigned int q0 = gt; q0 < n_db; q0 += gn) {
        const unsigned int id = F.db_id[q0];
        const int z = (int) (id % (unsigned int) nz), y = (int) ((id / (unsigned int) nz) % (unsigned int) ny),
                  x = (int) (id / ((unsigned int) nz * (unsigned int) ny));
        const int dx[7] = {0, 1, -1, 0, 0, 0, 0}, dy[7] = {0, 0, 0, 1, -1, 0, 0}, dz[7] = {0, 0, 0, 0, 0, 1, -1};
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const int xx = x + dx[q], yy = y + dy[q], zz = z + dz[q];
            if (xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz && g->present[0][xx] &&
                g->present[1][yy] && g->present[2][zz]) {
                const unsigned int nid = ((unsigned int) xx * (unsigned int) ny + (unsigned int) yy) *
                                             (unsigned int) nz + (unsigned int) zz;
                set_bit(F.test_bits, nid);
            }
        }
    }
}

// B5: ... into entry order, and the block-sorted training array (coordinates pre-scaled for the method's kernel)
struct BinEmit {
    const FusedArgs &F;
    int method;
    float ell, gp_scale;
    __device__ void operator()(unsigned int pos, unsigned int i, unsigned int cell) const {
        const float4 p = F.xy[i];
        // covSparse's  x / ell  (bgkinference.h:114) | GP's  scale * x  (gpregressor.h:115) hoisted
        F.pts[pos] = method == LA3DM_GP ? make_float4(gp_scale * p.x, gp_scale * p.y, gp_scale * p.z, p.w)
                                        : make_float4(p.x / ell, p.y / ell, p.z / ell, p.w);
        if (F.svals) { F.svals[pos] = i; F.skeys[pos] = cell; }
    }
};

__device__ __forceinline__ void bin_order_one(const FusedArgs &F, const BinEmit &emit, unsigned int i, unsigned int cell) {
    const unsigned int d = F.cell_db[cell] - 1;
    const unsigned int first = F.db_start[d], last = F.db_start[d + 1];
    if (last - first > kShortRun) return;                      // the warp-per-span pass has it
    unsigned int rank = 0;
    for (unsigned int li = first; li < last; ++li) rank += F.mlist[li] < i ? 1u : 0u;
    emit(first + rank, i, cell);
}

__device__ void bin_order(const FusedArgs &F, unsigned char *scratch) {
    ScanCounters *c = F.c;
    const bool dead = ld_volatile(&c->overflow) != 0u;
    const unsigned int n = dead ? 0u : min(c->n_train, F.train_cap);
    const float ell = F.P->ell;
    const BinEmit emit{F, F.P->method, ell, (float) (1.73205 / (double) ell)};   // gpregressor.h:115 (float(1.73205 / ell))
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    const unsigned int n_db = dead ? 0u : min(c->n_data_blocks, F.members_cap);
    const unsigned int *db_id = F.db_id, *db_start = F.db_start;
    // long spans: a warp / a CTA each (the cell of a span's entries = its data block's cell)
    if (!dead) order_long_spans(F, 2, db_start, F.mlist, reinterpret_cast<unsigned int *>(scratch), [&emit, db_id](unsigned int pos, unsigned int i, unsigned int d) { emit(pos, i, db_id[d]); });
    (void) n_db;
    for (unsigned int i0 = gt; i0 < n; i0 += 2 * gn) {
        unsigned int cell[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) { const unsigned int i = i0 + k * gn; cell[k] = i < n ? F.mcell[i] : kPad; }
#pragma unroll
        for (int k = 0; k < 2; ++k) if (cell[k] != kPad) bin_order_one(F, emit, i0 + k * gn, cell[k]);
    }
    const unsigned int ne = dead ? 0u : min(c->n_extra, kMaxExtra);
    for (unsigned int e = gt; e < ne; e += gn) {
        const Extra x = F.extra[e];
        bin_order_one(F, emit, x.entry, x.cell);
    }
}

// ---- test blocks = set bits of the cell bitmap, in ascending cell order (tiles of kTW words: the bitmap is dense) ------
__device__ void test_count(const FusedArgs &F, unsigned long long *smem) {
    const bool dead = ld_volatile(&F.c->overflow) != 0u;
    const unsigned int words = dead ? 0u : min((F.c->n_cells + 31u) >> 5, (F.cells_cap + 31u) >> 5);
    const unsigned int n_tiles = (words + kTW - 1) / kTW;
    (void) smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a warp per tile
    for (unsigned int t = blockIdx.x * (kFT / 32) + warp; t < n_tiles; t += gridDim.x * (kFT / 32)) {
        unsigned int s = 0;
        for (int j = lane; j < kTW; j += 32) { const unsigned int w = t * kTW + j; s += w < words ? __popc(F.test_bits[w]) : 0; }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) F.tsum[t] = s;
    }
}

__device__ inline long long test_block_key(unsigned int id, const GridDesc *g, int &x, int &y, int &z) {
    const int nz = g->n[2], ny = g->n[1];
    z = (int) (id % (unsigned int) nz);
    y = (int) ((id / (unsigned int) nz) % (unsigned int) ny);
    x = (int) (id / ((unsigned int) nz * (unsigned int) ny));
    return make_key(g->base[0] + x, g->base[1] + y, g->base[2] + z);
}

// test_id in ascending cell order: a CTA per tile, every thread writes ids (the q-th set bit of the tile)
__device__ void test_place(const FusedArgs &F, unsigned long long *smem, unsigned char *scratch) {
    ScanCounters *c = F.c;
    const bool dead = ld_volatile(&c->overflow) != 0u;
    const unsigned int words = dead ? 0u : min((c->n_cells + 31u) >> 5, (F.cells_cap + 31u) >> 5);
    const unsigned int n_tiles = (words + kTW - 1) / kTW;
    const unsigned int total = (unsigned int) block_sum_range<unsigned long long>(F.tsum, 0, n_tiles, smem);
    if (blockIdx.x == 0 && threadIdx.x == 0 && !dead) {
        c->n_test_blocks = total;
        if (total > F.tests_cap) atomicOr(&c->overflow, OVF_TESTS);
    }
    if (total > F.tests_cap) return;
    unsigned int *s_word = reinterpret_cast<unsigned int *>(scratch), *s_pre = s_word + kTW;      // [kTW], [kTW + 1]
    TileCarry<unsigned long long> carry;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned int w = t * kTW + threadIdx.x;
        const unsigned int word = (threadIdx.x < kTW && w < words) ? F.test_bits[w] : 0u;
        const unsigned int prefix = (unsigned int) carry.prefix(F.tsum, t, smem);
        unsigned long long tot;
        const unsigned int ex = (unsigned int) block_exclusive_scan<unsigned long long>((unsigned long long) __popc(word), smem, tot);
        if (threadIdx.x < kTW) { s_word[threadIdx.x] = word; s_pre[threadIdx.x] = ex; }
        __syncthreads();
        const unsigned int tile_total = (unsigned int) tot;
        for (unsigned int q = threadIdx.x; q < tile_total; q += kFT) {
            unsigned int lo = 0, hi = kTW;            // the last word whose prefix is <= q
            while (hi - lo > 1) { const unsigned int mid = (lo + hi) >> 1; if (s_pre[mid] <= q) lo = mid; else hi = mid; }
            unsigned int wd = s_word[lo];
            for (unsigned int k = q - s_pre[lo]; k > 0; --k) wd &= wd - 1;      // drop the lower set bits
            const unsigned int b = (unsigned int) __ffs(wd) - 1u;
            F.test_id[prefix + q] = (t * kTW + lo) * 32u + b;
        }
        __syncthreads();
    }
}

// slot of every test block in the map (0xFFFFFFFF: does not exist yet); the new ones are counted per tile of the next
// phase (1024 test blocks) with one atomic per warp on a cleared array.  Read-only on the map.  Two probes in flight.
__device__ void plan_find(const FusedArgs &F) {
    ScanCounters *c = F.c;
    const bool dead = ld_volatile(&c->overflow) != 0u;
    const unsigned int T = dead ? 0u : min(c->n_test_blocks, F.tests_cap);
    const unsigned int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    for (unsigned int t0 = gt; t0 < ((T + 31u) & ~31u); t0 += 2 * gn) {
        unsigned int id[2];
        int slot[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 2; ++k) { const unsigned int t = t0 + k * gn; id[k] = t < T ? F.test_id[t] : kPad; }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (id[k] == kPad) continue;
            int x, y, z;
            slot[k] = hash_find(F.hkeys, F.hvals, F.hmask, test_block_key(id[k], F.g, x, y, z));
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const unsigned int t = t0 + k * gn;                 // (gn is a multiple of 32: a warp's 32 t are one aligned group)
            const bool is_new = id[k] != kPad && slot[k] < 0;
            if (id[k] != kPad) F.plan[t].slot = (unsigned int) slot[k];
            const unsigned int m = __ballot_sync(0xffffffffu, is_new);
            if ((threadIdx.x & 31) == 0 && m) atomicAdd(&F.new_sums[t / kFT], (unsigned int) __popc(m));
        }
    }
}

// creation of the new blocks in test-block order, the 7 neighbour ranges, this rank's work lists: k_plan (binning.cu).
// The first phase of the scan that writes to the persistent map; every capacity check has been made by now.
__device__ void plan_fill(const FusedArgs &F, unsigned long long *smem, unsigned int *s_new) {
    ScanCounters *c = F.c;
    const ScanArgs *A = F.A;
    const GridDesc *g = F.g;
    const DevParams *P = F.P;
    if (ld_volatile(&c->overflow) != 0u) return;
    const unsigned int T = min(c->n_test_blocks, F.tests_cap);
    const unsigned int n_tiles = (T + kFT - 1) / kFT;
    if (blockIdx.x == 0) {
        const unsigned int total = block_sum_range<unsigned int>(F.new_sums, 0, n_tiles, reinterpret_cast<unsigned int *>(smem));
        if (threadIdx.x == 0) c->n_new_blocks = total;
    }
    const unsigned int *db_start = F.db_start;
    TileCarry<unsigned int> carry;
    for (unsigned int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const unsigned int t = tl * kFT + threadIdx.x;
        const bool valid = t < T;
        NeighbourPlan pl;
        pl.slot = valid ? F.plan[t].slot : 0u;
        const unsigned int tid_cell = valid ? F.test_id[t] : 0u;
        const unsigned int prefix = carry.prefix(F.new_sums, tl, reinterpret_cast<unsigned int *>(smem));
        pl.is_new = (valid && pl.slot == 0xFFFFFFFFu) ? 1u : 0u;
        unsigned int tot;
        const unsigned int rank = block_exclusive_scan<unsigned int>(pl.is_new, reinterpret_cast<unsigned int *>(smem), tot);
        int x = 0, y = 0, z = 0;
        long long key = 0;
        if (valid) key = test_block_key(tid_cell, g, x, y, z);
        if (pl.is_new) {
            pl.slot = A->n_blocks + prefix + rank;                       // < pool_cap: the host keeps room for `tests`
            F.keys[pl.slot] = key;
            hash_insert(F.hkeys, F.hvals, F.hmask, key, (int) pl.slot);
        }
        // the neighbour ranges first (their loads are in flight while the default records are written)
        const int nz = g->n[2], ny = g->n[1], nx = g->n[0];
        unsigned int totp = 0;
        // peers: a block that another rank owns only needs its slot here; ranges, plan entry and work lists are the owner's
        const bool mine_blk = valid && (!A->peers || block_owner(key, t, A->shard_world, true) == A->shard_rank);
        if (mine_blk) {
            const int dx[7] = {0, 1, -1, 0, 0, 0, 0}, dy[7] = {0, 0, 0, 1, -1, 0, 0}, dz[7] = {0, 0, 0, 0, 0, 1, -1};
            unsigned int dd[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int xx = x + dx[k], yy = y + dy[k], zz = z + dz[k];
                dd[k] = 0;
                if (xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz)
                    dd[k] = F.cell_db[((unsigned int) xx * (unsigned int) ny + (unsigned int) yy) * (unsigned int) nz + (unsigned int) zz];
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                unsigned int start = 0, count = 0;
                if (dd[k]) { start = db_start[dd[k] - 1]; count = db_start[dd[k]] - start; }
                if (F.plan_db) F.plan_db[(size_t) t * 8 + k] = dd[k];      // data block index + 1 (GP: locates the regressor)
                pl.start[k] = start;
                pl.count[k] = count;
                totp += count;
            }
        }
        if (F.init_records) {
            // Default records of the tile's new blocks, 16 bytes per thread and step, by the whole CTA.  Multi-GPU: the
            // block's owner writes the record into EVERY replica -- a peer may be ahead of us, and its results must not be
            // overwritten by a default record that we write later; stores of one GPU to one peer arrive in order, so the
            // owner's defaults land before the owner's results.
            const PeerTable *PT = A->peers;
            const int world = (PT && !PT->deferred) ? PT->world : 1, my_rank = PT ? PT->rank : 0;
            const bool mine = !PT || block_owner(key, t, PT->world, true) == PT->rank;
            if (PT && PT->deferred && mine && pl.is_new) F.dirty[pl.slot] = 1;
            __syncthreads();
            if (pl.is_new) s_new[rank] = mine ? pl.slot : kPad;
            __syncthreads();
            const int nodes = P->nodes, st_off = P->st_off, words = P->rec_bytes >> 4;
            const float da = P->def_a, db = P->def_b;
            const unsigned int leaves = (unsigned int) (P->finest & 0xFF);
            for (unsigned int idx = threadIdx.x; idx < tot * (unsigned int) words; idx += kFT) {
                const unsigned int sl = s_new[idx / (unsigned int) words];
                if (sl == kPad) continue;
                const int w = (int) (idx % (unsigned int) words);
                const size_t rec_off = (size_t) sl * (size_t) P->rec_bytes;
                uint4 v;
                unsigned int *vw = reinterpret_cast<unsigned int *>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int byte0 = 16 * w + 4 * q;
                    unsigned int word;
                    if (byte0 + 4 <= st_off) word = __float_as_uint(((byte0 >> 2) & 1) ? db : da);
                    else {
                        word = 0;
#pragma unroll
                        for (int bb = 0; bb < 4; ++bb) {
                            const int nn = byte0 + bb - st_off;
                            const unsigned int by = nn < nodes ? (unsigned int) LA3DM_UNKNOWN : (nn == nodes ? leaves : 0u);
                            word |= by << (8 * bb);
                        }
                    }
                    vw[q] = word;
                }
                reinterpret_cast<uint4 *>(F.pool + rec_off)[w] = v;
                for (int p = 0; p < world; ++p)
                    if (p != my_rank) reinterpret_cast<uint4 *>(PT->pool[p] + rec_off)[w] = v;
            }
        }
        // this rank's work lists; one atomic per warp and list (every test block of the scan passes here: 10^5 atomics on
        // one counter would serialise in L2)
        {
            const bool own = mine_blk && F.heavy_list && block_owner(key, t, A->shard_world, A->peers != nullptr) == A->shard_rank;
            const bool mega = own && totp > A->mega_tot;
            const bool heavy = own && !mega && totp > A->heavy_tot;
            const bool light = own && !mega && !heavy && A->shard_world > 1;     // (one rank: walked in cell order)
            const int lane = threadIdx.x & 31;
            const unsigned int lt = (1u << lane) - 1u;
            const unsigned int mh = __ballot_sync(0xffffffffu, heavy), ml = __ballot_sync(0xffffffffu, light);
            unsigned int bh = 0, bl = 0;
            if (lane == 0) {
                if (mh) bh = atomicAdd(&c->n_heavy, (unsigned int) __popc(mh));
                if (ml) bl = atomicAdd(&c->n_light, (unsigned int) __popc(ml));
            }
            bh = __shfl_sync(0xffffffffu, bh, 0);
            bl = __shfl_sync(0xffffffffu, bl, 0);
            if (heavy) F.heavy_list[bh + __popc(mh & lt)] = t;
            if (light) F.light_list[bl + __popc(ml & lt)] = t;
            if (mega) {                    // cut into chunks, each predicted as a unit of its own (predict_bgk.cu)
                const unsigned int nch = (totp + A->mega_chunk - 1u) / A->mega_chunk;
                const unsigned int first = atomicAdd(&c->n_mega_chunks, nch), m = atomicAdd(&c->n_mega, 1u);
                F.mega_list[m] = make_uint4(t, first, nch, 0u);
                for (unsigned int q = 0; q < nch; ++q) F.chunk_mega[first + q] = m;
            }
        }
        if (!valid) continue;
        F.touched[pl.slot] = 1;                // read side: la3dm_export_touched
        if (F.cell_test) F.cell_test[tid_cell] = t + 1;
        uint4 *dst = reinterpret_cast<uint4 *>(F.plan + t);
        const uint4 *src = reinterpret_cast<const uint4 *>(&pl);
        if (mine_blk) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; }
        dst[3] = src[3];                   // (count[5..6], slot, is_new)
    }
}

constexpr int kScratchBytes = 28 * 1024;

// ---- the three kernels -----------------------------------------------------------------------------------------------
__global__ void k_fused_begin(ScanCounters *c, unsigned int *mm, int reset_counters) {
    if (threadIdx.x == 0 && reset_counters) *c = ScanCounters();
    if (threadIdx.x < 18 && reset_counters) mm[threadIdx.x] = (threadIdx.x % 6) < 3 ? 0xFFFFFFFFu : 0u;   // flipped min | max
}

// The barrier counter goes back to zero when the last CTA leaves the kernel (every CTA that counts itself out has left its
// last barrier, so nobody is spinning on the counter any more): the next kernel / scan starts from a clean one.
__device__ __forceinline__ void grid_leave(unsigned int *bar, unsigned int *out_count) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(out_count, 1u);
        if (prev == gridDim.x - 1u) { *bar = 0u; *out_count = 0u; __threadfence(); }
    }
}

// get_training_data: both voxel grids and the beam sampling (19 phases)
__device__ __forceinline__ void run_frontend(const FusedArgs &F, unsigned int *bar, unsigned int &epoch, unsigned long long *tr,
                                             unsigned long long *smem, unsigned char *scratch) {
    // P0: bounding box of the cloud (the bitmap, the ticket counters and the beam sums were cleared before)
    {
        const ScanArgs *A = F.A;
        Box b;
        if (ld_volatile(&F.c->overflow) == 0u)
            for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < A->n; i += gridDim.x * blockDim.x) {
                const float *p = A->xyz + (size_t) i * A->stride_f;
                b.add(p[0], p[1], p[2]);
            }
        unsigned int *s_mm = reinterpret_cast<unsigned int *>(scratch);
        b.flush(F.mm, s_mm);
    }
    grid_sync(bar, epoch, tr);
    vg_keys<0>(F);                          grid_sync(bar, epoch, tr);
    vg_bits_count<0>(F, smem);              grid_sync(bar, epoch, tr);
    vg_bits_prefix<0>(F, smem);             grid_sync(bar, epoch, tr);
    vg_rank<0>(F);                          grid_sync(bar, epoch, tr);
    vg_span_count<0>(F, smem);              grid_sync(bar, epoch, tr);
    vg_span_place<0>(F, smem);              grid_sync(bar, epoch, tr);
    vg_drop<0>(F);                          grid_sync(bar, epoch, tr);
    vg_order<0>(F, scratch);                grid_sync(bar, epoch, tr);
    vg_centroid<0>(F, smem, scratch);       grid_sync(bar, epoch, tr);
    beam_fill(F, smem, scratch);            grid_sync(bar, epoch, tr);
    vg_bits_count<1>(F, smem);              grid_sync(bar, epoch, tr);
    vg_bits_prefix<1>(F, smem);             grid_sync(bar, epoch, tr);
    vg_rank<1>(F);                          grid_sync(bar, epoch, tr);
    vg_span_count<1>(F, smem);              grid_sync(bar, epoch, tr);
    vg_span_place<1>(F, smem);              grid_sync(bar, epoch, tr);
    vg_drop<1>(F);                          grid_sync(bar, epoch, tr);
    vg_order<1>(F, scratch);                grid_sync(bar, epoch, tr);
    vg_centroid<1>(F, smem, scratch);
}

__device__ __forceinline__ void frontend_clear(const FusedArgs &F) {
    zero_words(F.bits, (size_t) F.bits_words + 4);
    zero_words(F.vcnt, F.vcnt_cap);
    zero_words(reinterpret_cast<unsigned int *>(F.bsum), 2 * (size_t) F.bsum_n);
}

// the per-scan R-tree's job: training entries grouped by block, data blocks, test-block candidates (6 phases)
__device__ __forceinline__ void run_binning(const FusedArgs &F, unsigned int *bar, unsigned int &epoch, unsigned long long *tr,
                                            unsigned long long *smem, unsigned char *scratch) {
    // B0: CTA 0 steps the block grid; everyone clears the dense per-cell tables
    if (blockIdx.x == 0) block_grid(F, scratch);
    zero_words(F.cell_cnt, F.cells_cap);
    zero_words(F.cell_db, F.cells_cap);
    zero_words(F.test_bits, (F.cells_cap + 31u) / 32u + 1u);
    if (F.cell_test) zero_words(F.cell_test, F.cells_cap);
    zero_words(F.new_sums, F.tests_cap / kFT + 2);
    grid_sync(bar, epoch, tr);
    bin_members(F);                         grid_sync(bar, epoch, tr);
    bin_cell_count(F, smem);                grid_sync(bar, epoch, tr);
    bin_cell_place(F, smem);                grid_sync(bar, epoch, tr);
    bin_drop(F);                            grid_sync(bar, epoch, tr);
    bin_order(F, scratch);
}

// test blocks, their slots in the map, their neighbour plans (4 phases)
__device__ __forceinline__ void run_plan(const FusedArgs &F, unsigned int *bar, unsigned int &epoch, unsigned long long *tr,
                                         unsigned long long *smem, unsigned char *scratch) {
    test_count(F, smem);                    grid_sync(bar, epoch, tr);
    test_place(F, smem, scratch);           grid_sync(bar, epoch, tr);
    plan_find(F);                           grid_sync(bar, epoch, tr);
    plan_fill(F, smem, reinterpret_cast<unsigned int *>(scratch));
}

__global__ void __launch_bounds__(kFT, 1) k_fused_frontend(const FusedArgs F) {
    __shared__ unsigned long long smem[66];
    __shared__ __align__(16) unsigned char scratch[kScratchBytes];
    unsigned int epoch = 0;
    unsigned long long *tr = F.trace;
    trace_mark(tr, 0);
    frontend_clear(F);
    run_frontend(F, F.bar, epoch, tr, smem, scratch);
    __syncthreads();
    trace_mark(tr, 1);
    grid_leave(F.bar, F.bar + 3);
}

__global__ void __launch_bounds__(kFT, 1) k_fused_binning(const FusedArgs F) {
    __shared__ unsigned long long smem[66];
    __shared__ __align__(16) unsigned char scratch[kScratchBytes];
    unsigned int epoch = 0;
    unsigned long long *tr = F.trace ? F.trace + 64 : nullptr;
    trace_mark(tr, 0);
    run_binning(F, F.bar, epoch, tr, smem, scratch);
    __syncthreads();
    trace_mark(tr, 1);
    grid_leave(F.bar, F.bar + 3);
}

__global__ void __launch_bounds__(kFT, 1) k_fused_plan(const FusedArgs F) {
    __shared__ unsigned long long smem[66];
    __shared__ __align__(16) unsigned char scratch[4096];
    unsigned int epoch = 0;
    unsigned long long *tr = F.trace ? F.trace + 128 : nullptr;
    trace_mark(tr, 0);
    run_plan(F, F.bar, epoch, tr, smem, scratch);
    __syncthreads();
    trace_mark(tr, 1);
    grid_leave(F.bar, F.bar + 3);
}

// (One kernel for all 29 phases was tried: under the 64-register cap of a 1 024-thread CTA the merged function spills in its
// hot loops and every phase got slower -- 298 us against 280 us for the three kernels including their launch gaps.)

}  // namespace

// buffers of the fused pipeline for the current `caps` (called from ensure_workspace)
bool Map::ensure_fused_workspace() {
    bool moved = false;
    const size_t n_vox = std::max<size_t>(caps.points, caps.raw) + 8;
    const size_t words = (size_t) caps.vg_cells / 32 + 16;
    moved |= fz_vcnt.reserve(n_vox * 4, stream);
    moved |= fz_bits.reserve(words * 4, stream);
    moved |= fz_wpre.reserve(words * 4, stream);
    moved |= fz_cell_cnt.reserve(((size_t) caps.cells + 8) * 4, stream);
    moved |= fz_extra.reserve((size_t) kMaxExtra * sizeof(Extra), stream);
    moved |= fz_bsum.reserve(((size_t) caps.points / kBeamTile + 8) * 8, stream);
    {
        const size_t n_max = std::max<size_t>(std::max<size_t>(caps.points, caps.raw), caps.members);
        moved |= fz_long.reserve((n_max / kShortRun + n_max / kWideRun + 256) * 4, stream);
    }
    moved |= fz_newsums.reserve(((size_t) caps.tests / kFT + 8) * 4, stream);
    moved |= fz_tsum.reserve((std::max<size_t>(std::max<size_t>(n_vox, words), std::max<size_t>(caps.cells, caps.tests)) / kFT + 8) * 8, stream);
    if (!fz_bar) {
        LA3DM_CUDA(cudaMalloc(&fz_bar, 16 + 3 * 64 * 8));
        LA3DM_CUDA(cudaMemset(fz_bar, 0, 16 + 3 * 64 * 8));
        // the grid barrier needs every CTA resident: one CTA of 1 024 threads per SM must fit and the device must take
        // cooperative launches, else the sort-based pipeline is the one that runs
        int coop = 0, o0 = 0, o1 = 0, o2 = 0;
        LA3DM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        LA3DM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, k_fused_frontend, kFT, 0));
        LA3DM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_fused_binning, kFT, 0));
        LA3DM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_fused_plan, kFT, 0));
        if (!coop || o0 < 1 || o1 < 1 || o2 < 1) use_fused = false;
    }
    return moved;
}

bool Map::fused_applicable(int mode) const {
    if (!use_fused) return false;
    if (hp.method != LA3DM_BGK && hp.method != LA3DM_GP) return false;
    // two cooperative grids cannot share the device: a replica waiting for its peer (k_peer_wait) would keep the peer's
    // front-end from starting -- replicas on ONE device (tests) stay on the legacy pipeline
    if (peers_attached && peers_share_device) return false;
    if (caps.vg_cells > (1u << 28)) return false;          // bitmap of the voxel grid: 32 MB at most
    return mode == 0 || mode == 1 || mode == 2 || mode == 3;
}

// stage: 0 front-end, 1 binning, 2 plan
void Map::enqueue_fused(int stage) {
    FusedArgs F{};
    F.A = d_args; F.c = d_cnt; F.mm = d_mm; F.P = d_params; F.g = d_grid; F.bar = fz_bar;
    static const bool trace = getenv("LA3DM_FUSED_TRACE") != nullptr;
    F.trace = trace ? reinterpret_cast<unsigned long long *>(fz_bar + 4) : nullptr;
    F.bits = fz_bits.as<unsigned int>(); F.wpre = fz_wpre.as<unsigned int>();
    F.bits_words = (unsigned int) ((size_t) caps.vg_cells / 32 + 8);
    F.vg_cells_cap = caps.vg_cells;
    F.vkey = sort_keys[0].as<unsigned int>(); F.varr = sort_keys[1].as<unsigned int>();
    F.vcnt = fz_vcnt.as<unsigned int>(); F.vstart = run_start.as<unsigned int>();
    F.vlist = sort_vals[0].as<unsigned int>(); F.vsorted = sort_vals[1].as<unsigned int>();
    F.vcnt_cap = (unsigned int) (std::max<size_t>(caps.points, caps.raw) + 1);
    F.tsum = fz_tsum.as<unsigned long long>();
    F.bsum = fz_bsum.as<unsigned long long>(); F.bsum_n = caps.points / kBeamTile + 4;
    {
        const size_t n_max = std::max<size_t>(std::max<size_t>(caps.points, caps.raw), caps.members);
        F.longs = fz_long.as<unsigned int>();
        F.long_capA = (unsigned int) (n_max / kShortRun + 64); F.long_capB = (unsigned int) (n_max / kWideRun + 64);
    }
    F.hits_ds = hits_ds.as<float4>(); F.frees_raw = frees_raw.as<float4>(); F.xy = xy.as<float4>();
    F.hit_cnt = hit_cnt.as<unsigned int>();
    F.points_cap = caps.points; F.raw_cap = caps.raw;
    F.cell_cnt = fz_cell_cnt.as<unsigned int>(); F.cell_db = cell_db.as<unsigned int>();
    F.test_bits = test_bits.as<unsigned int>();
    F.cell_test = hp.method == LA3DM_GP ? cell_test.as<unsigned int>() : nullptr;
    F.mcell = sort_keys[0].as<unsigned int>(); F.mk = sort_keys[1].as<unsigned int>();
    F.mlist = sort_vals[0].as<unsigned int>();
    F.extra = fz_extra.as<Extra>();
    F.pts = pts_sorted.as<float4>();
    F.db_id = db_id.as<unsigned int>(); F.db_start = db_start.as<unsigned int>();
    F.skeys = nullptr; F.svals = nullptr;
    F.cells_cap = caps.cells; F.members_cap = caps.members; F.train_cap = caps.train;
    F.test_id = test_id.as<unsigned int>(); F.plan = plan.as<NeighbourPlan>();
    F.plan_db = hp.method == LA3DM_GP ? plan_db.as<unsigned int>() : nullptr;
    F.heavy_list = hp.method == LA3DM_BGK ? heavy_list.as<unsigned int>() : nullptr;
    F.light_list = light_list.as<unsigned int>();
    F.mega_list = mega_list.as<uint4>(); F.chunk_mega = chunk_mega.as<unsigned int>();
    F.dirty = dirty.as<unsigned char>(); F.touched = touched.as<unsigned char>();
    F.hkeys = hkeys.as<long long>(); F.hvals = hvals.as<int>(); F.hmask = hash_cap - 1;
    F.keys = keys.as<long long>(); F.pool = pool.as<unsigned char>();
    F.tests_cap = caps.tests;
    F.new_sums = fz_newsums.as<unsigned int>();
    F.init_records = hp.method == LA3DM_BGK ? 1 : 0;
    void *args[] = {&F};
    const void *fn = stage == 0 ? (const void *) k_fused_frontend
                                : (stage == 1 ? (const void *) k_fused_binning : (const void *) k_fused_plan);
    // small scans: fewer CTAs, cheaper barriers (one CTA: the barrier is a __syncthreads)
    static const size_t per_cta = getenv("LA3DM_FUSED_ITEMS_PER_CTA") ? (size_t) atoll(getenv("LA3DM_FUSED_ITEMS_PER_CTA")) : 1024;
    const size_t items = stage == 0 ? std::max<size_t>(caps.points, caps.raw)
                                    : (stage == 1 ? std::max<size_t>(caps.train, caps.cells / 4) : std::max<size_t>(caps.tests, caps.cells / 32));
    const unsigned int grid = (unsigned int) std::min<size_t>((size_t) num_sms, std::max<size_t>(1, (items + per_cta - 1) / per_cta));
    LA3DM_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kFT), args, 0, stream));
    ++launches;
}

// LA3DM_FUSED_TRACE=1: per-phase times of the last scan (CTA 0's clock), printed to stderr
void Map::dump_fused_trace() {
    static const bool trace = getenv("LA3DM_FUSED_TRACE") != nullptr;
    if (!trace || !fz_bar) return;
    unsigned long long h[3 * 64];
    LA3DM_CUDA(cudaMemcpy(h, fz_bar + 4, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) {
        const unsigned long long *t = h + 64 * k;
        if (!t[0]) continue;
        fprintf(stderr, "[fused %d] start %+.1f end %+.1f total %.1f us:", k, ((double) t[0] - (double) h[0]) / 1000.0,
                ((double) t[1] - (double) h[0]) / 1000.0, (double) (t[1] - t[0]) / 1000.0);
        unsigned long long prev = t[0];
        for (int ph = 1; ph < 31 && t[2 * ph]; ++ph) {
            fprintf(stderr, " %d:%.1f+%.1f", ph, (double) (t[2 * ph] - prev) / 1000.0, (double) (t[2 * ph + 1] - t[2 * ph]) / 1000.0);
            prev = t[2 * ph + 1];
        }
        fprintf(stderr, " last:%.1f\n", (double) (t[1] - prev) / 1000.0);
    }
    LA3DM_CUDA(cudaMemset(fz_bar + 4, 0, sizeof(h)));
}

void Map::enqueue_fused_begin(int reset_counters) {
    k_fused_begin<<<1, 32, 0, stream>>>(d_cnt, d_mm, reset_counters);
    ++launches;
}

}  // namespace la3dm_b200
