// la3dm_b200 -- GPU front-end of insert_pointcloud: voxel-grid downsampling, range filter and free-space beam
// sampling.  Replaces get_training_data / downsample / beam_sample
// (src/bgkoctomap/bgkoctomap.cpp:383-458; PCL's VoxelGrid is restated from its published algorithm).
//
// All arithmetic mirrors the reference's fp32 evaluation order (library built with -fmad=false).
//
// Nothing here synchronises with the host: element counts live in ScanCounters on the device, kernels are launched
// over the workspace capacities (Caps) and check the counts themselves, and a count that does not fit raises an OVF_*
// bit that the host inspects once at the end of the scan (engine.cu).
#include <cub/cub.cuh>

#include "engine.cuh"
#include "runs.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kThreads = 256;
constexpr unsigned int kPad = 0xFFFFFFFFu;

// which point set a voxel-grid pass reads: 0 = the cloud (ScanArgs), 1 = the raw free samples
template <int W>
__device__ __forceinline__ void vg_source(const ScanArgs *A, const ScanCounters *c, const float4 *frees_raw,
                                          const float *&p, int &stride, unsigned int &n) {
    if (W == 0) { p = A->xyz; stride = A->stride_f; n = A->n; }
    else { p = reinterpret_cast<const float *>(frees_raw); stride = 4; n = c->n_raw_frees; }
}

__global__ void k_scan_begin(ScanCounters *c, unsigned int *mm) {
    if (threadIdx.x == 0) *c = ScanCounters();
    if (threadIdx.x < 18) mm[threadIdx.x] = (threadIdx.x % 6) < 3 ? 0xFFFFFFFFu : 0u;   // flipped min | max
}

__device__ inline void block_minmax_box(bool valid, float *mn, float *mx, unsigned int *mm, unsigned int *s_mm);

__device__ inline void minmax_accumulate(const float *__restrict__ in, int stride, unsigned int n, unsigned int *mm) {
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    bool any = false;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *p = in + (size_t) i * stride;
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = p[a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
        any = true;
    }
    __shared__ unsigned int s_mm[6];
    block_minmax_box(any, mn, mx, mm, s_mm);
}

// warp-level min/max of one point per lane into mm[0..2] (min) / mm[3..5] (max), flipped-uint encoding.
// Every lane of the warp must call it; `valid` says whether this lane contributes.
__device__ inline void warp_minmax3(bool valid, float x, float y, float z, unsigned int *mm) {
    if (!__any_sync(0xffffffffu, valid)) return;
    float mn[3] = {valid ? x : 3.402823466e+38f, valid ? y : 3.402823466e+38f, valid ? z : 3.402823466e+38f};
    float mx[3] = {valid ? x : -3.402823466e+38f, valid ? y : -3.402823466e+38f, valid ? z : -3.402823466e+38f};
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&mm[a], float_flip(mn[a]));
            atomicMax(&mm[3 + a], float_flip(mx[a]));
        }
    }
}

// same for a per-lane box
__device__ inline void warp_minmax_box(bool valid, float *mn, float *mx, unsigned int *mm) {
    if (!__any_sync(0xffffffffu, valid)) return;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!valid) { mn[a] = 3.402823466e+38f; mx[a] = -3.402823466e+38f; }
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&mm[a], float_flip(mn[a]));
            atomicMax(&mm[3 + a], float_flip(mx[a]));
        }
    }
}

// CTA-level version: warps combine in shared memory (s_mm: 6 words), then ONE set of global atomics per CTA -- thousands
// of warps hitting the same six addresses serialise in L2.  Every thread of the CTA must call it.
__device__ inline void block_minmax_box(bool valid, float *mn, float *mx, unsigned int *mm, unsigned int *s_mm) {
    if (threadIdx.x < 6) s_mm[threadIdx.x] = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
    __syncthreads();
    warp_minmax_box(valid, mn, mx, s_mm);
    __syncthreads();
    if (threadIdx.x < 3) { if (s_mm[threadIdx.x] != 0xFFFFFFFFu) atomicMin(&mm[threadIdx.x], s_mm[threadIdx.x]); }
    else if (threadIdx.x < 6) { if (s_mm[threadIdx.x] != 0u) atomicMax(&mm[threadIdx.x], s_mm[threadIdx.x]); }
    __syncthreads();
}

// getMinMax3D of a voxel-grid input
template <int W>
__global__ void k_vg_minmax(const ScanArgs *__restrict__ A, const ScanCounters *__restrict__ c,
                            const float4 *__restrict__ frees_raw, unsigned int *mm) {
    if (c->overflow) return;
    const float *in; int stride; unsigned int n;
    vg_source<W>(A, c, frees_raw, in, stride, n);
    minmax_accumulate(in, stride, n, mm);
}

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

// (linear voxel index, point index) for every point; slots past the point count get the pad key so that the
// fixed-size radix sort leaves them at the end.  ds_resolution < 0 means downsample() is the identity
// (src/bgkoctomap/bgkoctomap.cpp:420-423): handled like pcl's overflow passthrough, key = point index.
template <int W>
__global__ void k_vg_keys(const ScanArgs *__restrict__ A, ScanCounters *c, const float4 *__restrict__ frees_raw,
                          const unsigned int *__restrict__ mm, unsigned int *keys, unsigned int *vals,
                          unsigned int cap, unsigned int vg_cells_cap) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const float *in; int stride; unsigned int n;
    vg_source<W>(A, c, frees_raw, in, stride, n);
    if (c->overflow || i >= n) { keys[i] = kPad; vals[i] = 0; return; }
    const bool identity = A->ds < 0;
    VGFrame f;
    f.passthrough = true;
    f.cells = 0;
    if (!identity) f = vg_frame(mm, A->inv_ds);
    const bool pass = identity || f.passthrough;
    if (i == 0) {
        c->vg_passthrough[W] = pass ? 1u : 0u;
        const unsigned long long need = pass ? (unsigned long long) n : (unsigned long long) f.cells;
        if (need > (unsigned long long) vg_cells_cap) {
            atomicOr(&c->overflow, OVF_VGCELLS);
            atomicMax(&c->vg_cells_needed, (unsigned int) (need > 0x80000000ull ? 0x80000000ull : need));
        }
    }
    unsigned int key;
    if (pass) key = i;
    else {
        const float inv = A->inv_ds;
        const float *p = in + (size_t) i * stride;
        const int i0 = (int) (floorf(p[0] * inv) - (float) f.min_b[0]);
        const int i1 = (int) (floorf(p[1] * inv) - (float) f.min_b[1]);
        const int i2 = (int) (floorf(p[2] * inv) - (float) f.min_b[2]);
        key = (unsigned int) (i0 + i1 * f.mul1 + i2 * f.mul2);
    }
    keys[i] = key;
    vals[i] = i;
}

// run_start[r] = first sorted position of run r; run_start[total] = n; *d_total = total
__global__ void k_vg_run_place(const unsigned int *__restrict__ keys, const unsigned int *__restrict__ d_n,
                               unsigned int cap, const unsigned int *__restrict__ tile_sums, unsigned int n_tiles,
                               unsigned int *run_start, unsigned int *d_total, const ScanCounters *__restrict__ c) {
    __shared__ unsigned int smem[66];
    const unsigned int n = c->overflow ? 0u : min(*d_n, cap);
    unsigned int prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    const unsigned int base = blockIdx.x * kTile + threadIdx.x * kTileItems;
    unsigned int flags = 0, cnt = 0;
    if (base < n) {
        unsigned int prev = base ? keys[base - 1] : 0u;
#pragma unroll
        for (int k = 0; k < kTileItems; ++k) {
            const unsigned int i = base + k;
            if (i < n) {
                const unsigned int key = keys[i];
                if (i == 0 || key != prev) { flags |= 1u << k; ++cnt; }
                prev = key;
            }
        }
    }
    unsigned int cta_total;
    unsigned int pos = prefix + block_exclusive_scan(cnt, smem, cta_total);
#pragma unroll
    for (int k = 0; k < kTileItems; ++k)
        if (flags & (1u << k)) run_start[pos++] = base + k;
    if (blockIdx.x == 0 && threadIdx.x == 0) { run_start[total] = n; *d_total = total; }
}

// one thread per voxel: sequential fp32 sum in ascending input order, then / float(count)  (pcl CentroidPoint);
// runs longer than kMidRun are handed to k_vg_long (a warp each; a CTA each beyond kLongRun).
// Pass 1 (free centroids = the tail of the training set) also feeds the training-set bounding box mm_xy.
template <int W>
__global__ void k_vg_centroid(const ScanArgs *__restrict__ A, ScanCounters *c, const float4 *__restrict__ frees_raw,
                              const unsigned int *__restrict__ vals, const unsigned int *__restrict__ run_start,
                              const unsigned int *__restrict__ d_total, float4 *out, unsigned int *long_list,
                              unsigned int *mm_xy) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    bool done = false;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (!c->overflow && r < *d_total) {
        const float *in; int stride; unsigned int n;
        vg_source<W>(A, c, frees_raw, in, stride, n);
        const unsigned int first = run_start[r], last = run_start[r + 1];
        bool handed = false;
        if (last - first > (unsigned int) kLongRun) {
            const unsigned int slot = atomicAdd(&c->n_long_runs[W], 1u);
            if (slot < (unsigned int) kMaxLongRuns) { long_list[slot] = r; handed = true; }
        } else if (last - first > (unsigned int) kMidRun) {
            const unsigned int slot = atomicAdd(&c->n_mid_runs[W], 1u);
            if (slot < (unsigned int) kMaxMidRuns) { long_list[kMaxLongRuns + slot] = r; handed = true; }
        }
        if (!handed) {
            float sx = 0.f, sy = 0.f, sz = 0.f;
            for (unsigned int li = first; li < last; ++li) {
                const float *p = in + (size_t) vals[li] * stride;
                sx += p[0]; sy += p[1]; sz += p[2];
            }
            const float cnt = (float) (last - first);
            const unsigned int off = W == 0 ? 0u : c->n_hits;
            const float label = W == 0 ? 1.0f : A->free_label;
            cx = sx / cnt; cy = sy / cnt; cz = sz / cnt;
            out[off + r] = make_float4(cx, cy, cz, label);
            done = true;
        }
    }
    if (W == 1) {   // CTA-level box first: one set of global atomics per CTA
        __shared__ unsigned int s_mm[6];
        if (threadIdx.x < 6) s_mm[threadIdx.x] = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
        __syncthreads();
        warp_minmax3(done, cx, cy, cz, s_mm);
        __syncthreads();
        if (threadIdx.x < 3) { if (s_mm[threadIdx.x] != 0xFFFFFFFFu) atomicMin(&mm_xy[threadIdx.x], s_mm[threadIdx.x]); }
        else if (threadIdx.x < 6) { if (s_mm[threadIdx.x] != 0u) atomicMax(&mm_xy[threadIdx.x], s_mm[threadIdx.x]); }
    }
}

// Long runs.  The sorted array is cut into aligned sub-chunks of kSub positions; k_vg_long_flags (whole grid) records
// for every sub-chunk that lies inside a long run whether it is one point repeated (the sensor origin is pushed once
// per hit, so its voxel holds n_hits copies).  k_vg_long then walks each long run in order with one CTA: consecutive
// repeated-point sub-chunks are added with ONE add_repeat (exact, no loads), anything else is staged through shared
// memory and added element by element by three lanes (x, y, z).  Medium runs get one warp each.
constexpr int kLongThreads = 512;

// acc + v[0] + v[1] + ... + v[m - 1], added one by one in that order (pcl's CentroidPoint); v is 16-byte aligned shared
// memory, read eight values ahead of the dependent chain of additions
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
constexpr int kSub = 256;                      // points per sub-chunk (8 per lane)
constexpr int kMidSub = 128;                   // points a warp stages at a time for a medium run
constexpr int kSubsPerPass = 1024;             // sub-chunk flags staged in shared memory at a time

template <int W>
__global__ void __launch_bounds__(kLongThreads)
k_vg_long_flags(const ScanArgs *__restrict__ A, const ScanCounters *__restrict__ c,
                const float4 *__restrict__ frees_raw, const unsigned int *__restrict__ vals,
                const unsigned int *__restrict__ run_start, const unsigned int *__restrict__ long_list,
                float *lf_first, unsigned char *lf_same, unsigned int n_sub_cap) {
    if (c->overflow) return;
    const unsigned int nl = min(c->n_long_runs[W], (unsigned int) kMaxLongRuns);
    const float *in; int stride; unsigned int n;
    vg_source<W>(A, c, frees_raw, in, stride, n);
    const int lane = threadIdx.x & 31;
    const unsigned int gw = (blockIdx.x * kLongThreads + threadIdx.x) >> 5, n_w = gridDim.x * (kLongThreads / 32);
    for (unsigned int idx = 0; idx < nl; ++idx) {
        const unsigned int r = long_list[idx];
        const unsigned int g0 = (run_start[r] + kSub - 1) / kSub, g1 = run_start[r + 1] / kSub;   // whole sub-chunks
        for (unsigned int g = g0 + gw; g < g1 && g < n_sub_cap; g += n_w) {
            unsigned int vi[kSub / 32];
#pragma unroll
            for (int k = 0; k < kSub / 32; ++k) vi[k] = vals[g * kSub + lane + 32 * k];
            float px[kSub / 32], py[kSub / 32], pz[kSub / 32];
#pragma unroll
            for (int k = 0; k < kSub / 32; ++k) {
                const float *p = in + (size_t) vi[k] * stride;
                px[k] = p[0]; py[k] = p[1]; pz[k] = p[2];
            }
            const float fx = __shfl_sync(0xffffffffu, px[0], 0), fy = __shfl_sync(0xffffffffu, py[0], 0),
                        fz = __shfl_sync(0xffffffffu, pz[0], 0);
            bool eq = true;
#pragma unroll
            for (int k = 0; k < kSub / 32; ++k)
                eq = eq && __float_as_uint(px[k]) == __float_as_uint(fx) && __float_as_uint(py[k]) == __float_as_uint(fy) &&
                     __float_as_uint(pz[k]) == __float_as_uint(fz);
            eq = __all_sync(0xffffffffu, eq);
            if (lane == 0) {
                lf_first[3 * (size_t) g] = fx; lf_first[3 * (size_t) g + 1] = fy; lf_first[3 * (size_t) g + 2] = fz;
                lf_same[g] = eq ? 1 : 0;
            }
        }
    }
}

template <int W>
__global__ void __launch_bounds__(kLongThreads, 1)
k_vg_long(const ScanArgs *__restrict__ A, ScanCounters *c, const float4 *__restrict__ frees_raw,
          const unsigned int *__restrict__ vals, const unsigned int *__restrict__ run_start, float4 *out,
          const unsigned int *__restrict__ long_list, const float *__restrict__ lf_first,
          const unsigned char *__restrict__ lf_same, unsigned int n_sub_cap, unsigned int *mm_xy) {
    __shared__ float first[kSubsPerPass][3];   // the sub-chunk's first point
    __shared__ unsigned char same[kSubsPerPass];
    __shared__ __align__(16) float stage[3][kSub];
    __shared__ __align__(16) float wstage[kLongThreads / 32][3][kMidSub];
    if (c->overflow) return;
    // last kernel of the free-space voxel grid: the training set is complete (hits, then free centroids)
    if (W == 1 && blockIdx.x == 0 && threadIdx.x == 0) c->n_train = c->n_hits + c->n_frees;
    const unsigned int nl = min(c->n_long_runs[W], (unsigned int) kMaxLongRuns);
    const unsigned int nm = min(c->n_mid_runs[W], (unsigned int) kMaxMidRuns);
    const float *in; int stride; unsigned int n;
    vg_source<W>(A, c, frees_raw, in, stride, n);
    const unsigned int off = W == 0 ? 0u : c->n_hits;
    const float label = W == 0 ? 1.0f : A->free_label;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- medium runs: one warp each, staged kMidSub points at a time
    for (unsigned int idx = blockIdx.x * (kLongThreads / 32) + warp; idx < nm; idx += gridDim.x * (kLongThreads / 32)) {
        const unsigned int r = long_list[kMaxLongRuns + idx];
        const unsigned int run_first = run_start[r], run_last = run_start[r + 1];
        float acc = 0.f;
        float px[kMidSub / 32], py[kMidSub / 32], pz[kMidSub / 32];
        auto fetch = [&](unsigned int s0) {
            const unsigned int m = min((unsigned int) kMidSub, run_last - s0);
#pragma unroll
            for (int k = 0; k < kMidSub / 32; ++k) {
                const unsigned int j = lane + 32 * k;
                if (j < m) {
                    const float *p = in + (size_t) vals[s0 + j] * stride;
                    px[k] = p[0]; py[k] = p[1]; pz[k] = p[2];
                }
            }
        };
        fetch(run_first);
        for (unsigned int s0 = run_first; s0 < run_last; s0 += kMidSub) {
            const unsigned int m = min((unsigned int) kMidSub, run_last - s0);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kMidSub / 32; ++k) {
                const unsigned int j = lane + 32 * k;
                if (j < m) { wstage[warp][0][j] = px[k]; wstage[warp][1][j] = py[k]; wstage[warp][2][j] = pz[k]; }
            }
            __syncwarp();
            if (s0 + kMidSub < run_last) fetch(s0 + kMidSub);     // in flight while three lanes add this chunk
            if (lane < 3) acc = seq_sum(acc, wstage[warp][lane], m);
        }
        float *o = reinterpret_cast<float *>(out + off + r);
        if (lane < 3) {
            const float v = acc / (float) (run_last - run_first);
            o[lane] = v;
            if (W == 1) { atomicMin(&mm_xy[lane], float_flip(v)); atomicMax(&mm_xy[3 + lane], float_flip(v)); }
        } else if (lane == 3) o[3] = label;
    }
    // ---- long runs: one CTA each
    // (dealt from the last CTA backwards: the medium runs above fill the first CTAs)
    for (unsigned int idx = gridDim.x - 1 - blockIdx.x; idx < nl; idx += gridDim.x) {
        const unsigned int r = long_list[idx];
        const unsigned int run_first = run_start[r], run_last = run_start[r + 1];
        const unsigned int g0 = min((run_first + kSub - 1) / kSub, n_sub_cap), g1 = min(run_last / kSub, n_sub_cap);
        float acc = 0.f;
        // pieces: head [run_first, g0 * kSub), whole sub-chunks g0 .. g1 - 1, tail [g1 * kSub, run_last)
        unsigned int pos = run_first;
        unsigned int g = g0;
        // The head and the tail are fetched up front, together with the first pass of flags: three dependent round
        // trips (index, point, sum) in parallel instead of one after the other.  Threads 0..255 hold the head piece,
        // 256..511 the tail piece.
        const unsigned int a_len = (g0 < g1 && run_first == g0 * kSub)
                                       ? 0u : min(((g0 < g1) ? g0 * kSub : run_last) - run_first, (unsigned int) kSub);
        const unsigned int b_pos = g1 * kSub;
        const unsigned int b_len = (g0 < g1 && g1 - g0 <= (unsigned int) kSubsPerPass && b_pos < run_last)
                                       ? min(run_last - b_pos, (unsigned int) kSub) : 0u;
        float pre[3] = {0.f, 0.f, 0.f};
        {
            const unsigned int tq = threadIdx.x & (kSub - 1);
            const bool is_a = threadIdx.x < (unsigned int) kSub;
            if (is_a ? tq < a_len : tq < b_len) {
                const float *p = in + (size_t) vals[(is_a ? run_first : b_pos) + tq] * stride;
                pre[0] = p[0]; pre[1] = p[1]; pre[2] = p[2];
            }
        }
        while (pos < run_last) {
            if (g < g1 && pos == g * kSub) {
                const unsigned int nb = min(g1 - g, (unsigned int) kSubsPerPass);
                __syncthreads();
                for (unsigned int j = threadIdx.x; j < nb; j += kLongThreads) {
                    first[j][0] = lf_first[3 * (size_t) (g + j)]; first[j][1] = lf_first[3 * (size_t) (g + j) + 1];
                    first[j][2] = lf_first[3 * (size_t) (g + j) + 2];
                    same[j] = lf_same[g + j];
                }
                __syncthreads();
                for (unsigned int sc = 0; sc < nb; ++sc) {
                    if (same[sc]) {
                        // merge the following sub-chunks that repeat the same point: one add_repeat for all of them
                        unsigned int m = kSub;
                        const unsigned int f0 = __float_as_uint(first[sc][0]), f1 = __float_as_uint(first[sc][1]),
                                           f2 = __float_as_uint(first[sc][2]);
                        while (sc + 1 < nb && same[sc + 1] && __float_as_uint(first[sc + 1][0]) == f0 &&
                               __float_as_uint(first[sc + 1][1]) == f1 && __float_as_uint(first[sc + 1][2]) == f2) {
                            ++sc;
                            m += kSub;
                        }
                        if (threadIdx.x < 3) acc = add_repeat(acc, first[sc][threadIdx.x], m);
                    } else {
                        const unsigned int s0 = (g + sc) * kSub;
                        if (threadIdx.x < kSub) {
                            const float *p = in + (size_t) vals[s0 + threadIdx.x] * stride;
                            stage[0][threadIdx.x] = p[0]; stage[1][threadIdx.x] = p[1]; stage[2][threadIdx.x] = p[2];
                        }
                        __syncthreads();
                        if (threadIdx.x < 3) acc = seq_sum(acc, stage[threadIdx.x], (unsigned int) kSub);
                        __syncthreads();
                    }
                }
                g += nb;
                pos = g * kSub;
            } else {
                // head, tail, or sub-chunks beyond the flag capacity: staged
                const unsigned int end = (g < g1) ? g * kSub : run_last;
                const unsigned int m = min(end - pos, (unsigned int) kSub);
                if (pos == run_first && a_len == m) {                  // the prefetched head
                    if (threadIdx.x < m) { stage[0][threadIdx.x] = pre[0]; stage[1][threadIdx.x] = pre[1]; stage[2][threadIdx.x] = pre[2]; }
                } else if (b_len && pos == b_pos && b_len == m) {      // the prefetched tail
                    const unsigned int tq = threadIdx.x - (unsigned int) kSub;
                    if (threadIdx.x >= (unsigned int) kSub && tq < m) { stage[0][tq] = pre[0]; stage[1][tq] = pre[1]; stage[2][tq] = pre[2]; }
                } else if (threadIdx.x < m) {
                    const float *p = in + (size_t) vals[pos + threadIdx.x] * stride;
                    stage[0][threadIdx.x] = p[0]; stage[1][threadIdx.x] = p[1]; stage[2][threadIdx.x] = p[2];
                }
                __syncthreads();
                if (threadIdx.x < 3) acc = seq_sum(acc, stage[threadIdx.x], m);
                __syncthreads();
                pos += m;
            }
        }
        float *o = reinterpret_cast<float *>(out + off + r);
        if (threadIdx.x < 3) {
            const float v = acc / (float) (run_last - run_first);
            o[threadIdx.x] = v;
            if (W == 1) { atomicMin(&mm_xy[threadIdx.x], float_flip(v)); atomicMax(&mm_xy[3 + threadIdx.x], float_flip(v)); }
        } else if (threadIdx.x == 3) o[3] = label;
    }
}

// ---- range filter + beam sampling ---------------------------------------------------------------------------------
// per downsampled hit: 0 if the range filter drops it (src/bgkoctomap/bgkoctomap.cpp:394-398), else 1 + number of free
// points it emits (origin once per kept hit :404, samples d = fr, 2fr.. < l with fp32 accumulation :451-455, tail
// sample :456-457)
__device__ inline unsigned int hit_free_count(const float4 h, const ScanArgs *A) {
    const float dx = h.x - A->ox, dy = h.y - A->oy, dz = h.z - A->oz;
    const float s = dx * dx + dy * dy + dz * dz;
    if (A->max_range > 0) {
        const double l = sqrt((double) s);                 // point3f::norm() (point3f.h:207-214)
        if (l > (double) A->max_range) return 0u;
    }
    const float l = (float) sqrt((double) s);
    const float fr = A->fr;
    // number of e with beam_tab[e] < l (the table is non-decreasing, beam_tab[e] ~ (e + 1) fr): start from the estimate
    // l / fr and walk to the exact boundary; beyond the table continue the accumulation
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

// tile_sums[tile] = (kept hits << 32) | free points of the tile (kHitTile hits, one per thread);
// hit_cnt[i] = per-hit count (0 = dropped)
constexpr int kHitTile = 256;

__global__ void __launch_bounds__(kHitTile)
k_hit_count(const float4 *__restrict__ hits, const ScanCounters *__restrict__ c, const ScanArgs *__restrict__ A,
            unsigned int *hit_cnt, unsigned long long *tile_sums) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const unsigned int n = c->overflow ? 0u : c->n_ds_hits;
    unsigned long long acc = 0;
    const unsigned int i = blockIdx.x * kHitTile + threadIdx.x;
    if (i < n) {
        const unsigned int cnt = hit_free_count(hits[i], A);
        hit_cnt[i] = cnt;
        if (cnt) acc = (1ull << 32) | (unsigned long long) cnt;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_sum;
}

// writes the kept hits (label 1) to xy[0 .. n_hits) and every free point to frees_raw, in the reference's push order.
// Positions come from a scan over the hits; the samples of one hit are then written by a whole warp (contiguous
// 16-byte stores).  Sample e of a beam sits at d_e = fr + fr + ... (e fp32 additions, :451-455) = add_repeat(fr, fr, e).
constexpr int kFillThreads = 4 * 256;   // = 4 * kHitTile: the free points of a hit are written by 4 threads

__global__ void __launch_bounds__(kFillThreads)
k_hit_fill(const float4 *__restrict__ hits, ScanCounters *c, const ScanArgs *__restrict__ A,
           const unsigned int *__restrict__ hit_cnt, const unsigned long long *__restrict__ tile_sums,
           unsigned int n_tiles, float4 *xy, float4 *frees, unsigned int raw_cap, unsigned int *mm_raw,
           unsigned int *mm_xy) {
    __shared__ unsigned long long smem[66];
    __shared__ unsigned int s_off[kHitTile];       // first free point of hit h, relative to the tile
    __shared__ unsigned int s_cnt[kHitTile];
    __shared__ float4 s_beam[kHitTile];            // (nx, ny, nz, l) of beam_sample's preamble (:437-449)
    const unsigned int n = c->overflow ? 0u : c->n_ds_hits;
    unsigned long long prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    const unsigned int n_hits = (unsigned int) (total >> 32), n_raw = (unsigned int) (total & 0xFFFFFFFFull);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->n_hits = n_hits;
        c->n_raw_frees = n_raw;
        if (n_raw > raw_cap) atomicOr(&c->overflow, OVF_RAW);
    }
    if (n_raw > raw_cap) return;
    const unsigned int i = blockIdx.x * kHitTile + threadIdx.x;
    const unsigned int cnt = (threadIdx.x < kHitTile && i < n) ? hit_cnt[i] : 0u;
    unsigned long long cta_total;
    const unsigned long long mine = cnt ? ((1ull << 32) | (unsigned long long) cnt) : 0ull;
    const unsigned long long excl = block_exclusive_scan(mine, smem, cta_total);
    const float ox = A->ox, oy = A->oy, oz = A->oz, fr = A->fr;
    // bounding boxes on the fly: the free samples (getMinMax3D of the second voxel grid) and the kept hits (their part
    // of the training-set bbox, src/bgkoctomap/bgkoctomap.cpp:464-484)
    float smn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, smx[3] = {-smn[0], -smn[0], -smn[0]};
    float hmn[3] = {smn[0], smn[0], smn[0]}, hmx[3] = {-smn[0], -smn[0], -smn[0]};
    bool any_s = false, any_h = false;
    // ---- one thread per hit: the kept hit itself (:399) and its beam's direction and length
    if (threadIdx.x < kHitTile) {
        s_off[threadIdx.x] = (unsigned int) (excl & 0xFFFFFFFFull);
        s_cnt[threadIdx.x] = cnt;
    }
    if (cnt) {
        const float4 hit = hits[i];
        xy[(unsigned int) ((prefix + excl) >> 32)] = make_float4(hit.x, hit.y, hit.z, 1.0f);
        hmn[0] = hmx[0] = hit.x; hmn[1] = hmx[1] = hit.y; hmn[2] = hmx[2] = hit.z;
        any_h = true;
        const float dx = hit.x - ox, dy = hit.y - oy, dz = hit.z - oz;
        const float l = (float) sqrt((double) (dx * dx + dy * dy + dz * dz));
        s_beam[threadIdx.x] = make_float4(dx / l, dy / l, dz / l, l);
    }
    __syncthreads();
    // ---- four threads per hit write its free points in the reference's push order (origin :404, samples :451-457)
    (void) cta_total;
    float4 *out = frees + (unsigned int) (prefix & 0xFFFFFFFFull);
    const unsigned int tab_n = A->beam_tab_n;
    const float *__restrict__ tab = A->beam_tab;
    {
        const unsigned int h = threadIdx.x >> 2, ch = s_cnt[h];
        if (ch) {
            const unsigned int first = s_off[h];
            const float4 bm = s_beam[h];
            const unsigned int tail = bm.w > fr ? 1u : 0u;
            const unsigned int n_reg = ch - 1u - tail;                                    // samples with d < l (:451-455)
            for (unsigned int e1 = threadIdx.x & 3u; e1 < ch; e1 += 4u) {
                float sx = ox, sy = oy, sz = oz;
                if (e1) {
                    const unsigned int e = e1 - 1u;
                    const float d = e < n_reg ? (e < tab_n ? tab[e] : add_repeat(fr, fr, e)) : bm.w - fr;   // :453 | :457
                    sx = ox + bm.x * d; sy = oy + bm.y * d; sz = oz + bm.z * d;
                }
                out[first + e1] = make_float4(sx, sy, sz, 0.f);
                smn[0] = fminf(smn[0], sx); smx[0] = fmaxf(smx[0], sx);
                smn[1] = fminf(smn[1], sy); smx[1] = fmaxf(smx[1], sy);
                smn[2] = fminf(smn[2], sz); smx[2] = fmaxf(smx[2], sz);
                any_s = true;
            }
        }
    }
    __shared__ unsigned int s_mm[6];
    block_minmax_box(any_s, smn, smx, mm_raw, s_mm);
    block_minmax_box(any_h, hmn, hmx, mm_xy, s_mm);
}

// insert_training_data (src/bgkoctomap/bgkoctomap.cpp:82-94): the caller's labelled points ARE the training set.  Copies
// them to xy, accumulates the bounding box (bbox(), :464-484) and sets the counters the binning stage reads.
__global__ void k_td_load(const ScanArgs *__restrict__ A, ScanCounters *c, float4 *xy, unsigned int train_cap,
                          unsigned int *mm_xy) {
    const unsigned int n = A->n;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->n_hits = 0; c->n_frees = 0; c->n_train = n;
        if (n > train_cap) atomicOr(&c->overflow, OVF_RAW);
    }
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-mn[0], -mn[0], -mn[0]};
    bool any = false;
    if (n <= train_cap)
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const float *p = A->xyz + (size_t) i * A->stride_f;
            const float4 v = make_float4(p[0], p[1], p[2], p[3]);
            xy[i] = v;
            mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
            mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
            any = true;
        }
    __shared__ unsigned int s_mm[6];
    block_minmax_box(any, mn, mx, mm_xy, s_mm);
}

// ---- ingest: the server-side steps in front of insert_pointcloud (src/bgkoctomap/bgkoctomap_server.cpp:70-86) --------
// pcl_ros::transformPointCloud -> pcl::transformPointCloud(cloud, Eigen::Matrix4f): p' = T p in fp32, evaluated like
// pcl::detail::Transformer (PCL >= 1.10): (m0 x + m1 y) + (m2 z + m3) per row, products and sums rounded separately.
// (PCL is a third-party dependency absent from /root/reference: this seam is "parity unpinned", see DESIGN.md.)
__global__ void k_tf_points(const ScanArgs *__restrict__ A, float4 *out) {
    const unsigned int n = A->n;
    const float *m = A->tf;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *p = A->raw_xyz + (size_t) i * A->raw_stride_f;
        const float x = p[0], y = p[1], z = p[2];
        out[i] = make_float4((m[0] * x + m[1] * y) + (m[2] * z + m[3]), (m[4] * x + m[5] * y) + (m[6] * z + m[7]),
                             (m[8] * x + m[9] * y) + (m[10] * z + m[11]), 0.f);
    }
}

// after the prefilter's voxel grid: the filtered cloud (hits_ds) becomes the scan's cloud if more than min_points
// points are left (bgkoctomap_server.cpp:84: `if (filtered_cloud.size() > 5)`), else the scan is empty; the counters the
// prefilter used are reset for the scan proper (overflow bits are kept)
__global__ void k_ingest_commit(ScanArgs *A, ScanCounters *c, unsigned int *mm, const float4 *__restrict__ filtered,
                                int prefiltered) {
    __shared__ unsigned int s_n;
    if (threadIdx.x == 0) s_n = prefiltered ? c->n_ds_hits : A->n;
    __syncthreads();
    const unsigned int n = s_n;
    if (prefiltered)
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            A->stage_cloud[i] = filtered[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned int ovf = c->overflow;
        *c = ScanCounters();
        c->overflow = ovf;
        A->n = n > (unsigned int) max(A->min_points, 0) ? n : 0u;
        A->ds = A->scan_ds; A->inv_ds = A->scan_inv_ds;
    }
    if (blockIdx.x == 0 && threadIdx.x < 18) mm[threadIdx.x] = (threadIdx.x % 6) < 3 ? 0xFFFFFFFFu : 0u;
}

inline int bits_for(unsigned int n) {   // radix-sort end bit for keys < n
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long) n) ++b;
    return b;
}

}  // namespace

// pcl::VoxelGrid, pass `which` (0: cloud -> hits_ds, 1: raw frees -> xy behind the hits)
void Map::enqueue_voxel_grid(int which) {
    const unsigned int cap = which == 0 ? caps.points : caps.raw;
    unsigned int *mm = d_mm + 6 * which;
    cub::DoubleBuffer<unsigned int> dk(sort_keys[0].as<unsigned int>(), sort_keys[1].as<unsigned int>());
    cub::DoubleBuffer<unsigned int> dv(sort_vals[0].as<unsigned int>(), sort_vals[1].as<unsigned int>());
    unsigned int *tile_sums = tiles.as<unsigned int>();
    unsigned int *llist = long_list.as<unsigned int>() + (size_t) which * (kMaxLongRuns + kMaxMidRuns);
    unsigned int *runs = run_start.as<unsigned int>();
    const float4 *fr = frees_raw.as<float4>();
    const int grid = ceil_div(cap, kThreads);
    const int n_tiles = ceil_div(cap, kTile);
    const int mm_grid = std::max(1, std::min(grid, num_sms * 4));
    size_t tmp = cub_tmp_bytes;
    const int end_bit = bits_for(caps.vg_cells);
    if (which == 0) {
        k_vg_minmax<0><<<mm_grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, mm);
        k_vg_keys<0><<<grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, mm, dk.Current(), dv.Current(), cap,
                                                    caps.vg_cells);
    } else {   // the bounding box of the free samples was accumulated by k_hit_fill
        k_vg_keys<1><<<grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, mm, dk.Current(), dv.Current(), cap,
                                                    caps.vg_cells);
        --launches;
    }
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, dk, dv, (int) cap, 0, end_bit, stream));
    const unsigned int *ks = dk.Current(), *vs = dv.Current();
    const unsigned int *d_n = which == 0 ? &d_args->n : &d_cnt->n_raw_frees;
    unsigned int *d_total = which == 0 ? &d_cnt->n_ds_hits : &d_cnt->n_frees;
    k_run_count<<<n_tiles, kTileThreads, 0, stream>>>(ks, d_n, cap, tile_sums, d_cnt);
    k_vg_run_place<<<n_tiles, kTileThreads, 0, stream>>>(ks, d_n, cap, tile_sums, (unsigned int) n_tiles, runs,
                                                         d_total, d_cnt);
    const int long_grid = num_sms;
    const unsigned int n_sub_cap = cap / kSub + 1;
    float *lf_first = long_flags.as<float>();
    unsigned char *lf_same = reinterpret_cast<unsigned char *>(lf_first + 3 * (size_t) n_sub_cap);
    unsigned int *mm_xy = d_mm + 12;
    if (which == 0) {
        k_vg_centroid<0><<<grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, d_total, hits_ds.as<float4>(),
                                                        llist, mm_xy);
        k_vg_long_flags<0><<<long_grid, kLongThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, llist, lf_first,
                                                                   lf_same, n_sub_cap);
        k_vg_long<0><<<long_grid, kLongThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, hits_ds.as<float4>(), llist,
                                                             lf_first, lf_same, n_sub_cap, mm_xy);
    } else {
        k_vg_centroid<1><<<grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, d_total, xy.as<float4>(), llist,
                                                        mm_xy);
        k_vg_long_flags<1><<<long_grid, kLongThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, llist, lf_first,
                                                                   lf_same, n_sub_cap);
        k_vg_long<1><<<long_grid, kLongThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, xy.as<float4>(), llist,
                                                             lf_first, lf_same, n_sub_cap, mm_xy);
    }
    launches += 7 + 2 + (end_bit + 7) / 8;   // ours + CUB radix sort (histogram, exclusive sum, onesweep passes)
}

// BGK / GP front-end.  On completion (stream-ordered): xy[0..n_train) = hits (label 1) then free centroids (label
// 0 / -1) and d_cnt->{n_ds_hits, n_hits, n_raw_frees, n_frees, n_train} are set.
void Map::enqueue_frontend_bgk() {
    enqueue_voxel_grid(0);
    const int n_tiles = ceil_div(caps.points, kHitTile);
    unsigned long long *tile_sums = tiles.as<unsigned long long>();
    k_hit_count<<<n_tiles, kHitTile, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, hit_cnt.as<unsigned int>(),
                                                  tile_sums);
    k_hit_fill<<<n_tiles, kFillThreads, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, hit_cnt.as<unsigned int>(),
                                                 tile_sums, (unsigned int) n_tiles, xy.as<float4>(),
                                                 frees_raw.as<float4>(), caps.raw, d_mm + 6, d_mm + 12);
    launches += 2;
    enqueue_voxel_grid(1);
}

// The whole scan, stream-ordered, no host synchronisation (this is what the CUDA graph captures).
void Map::enqueue_training_data() {
    const int grid = std::max(1, std::min(ceil_div(caps.train, kThreads), num_sms * 8));
    k_td_load<<<grid, kThreads, 0, stream>>>(d_args, d_cnt, xy.as<float4>(), caps.train, d_mm + 12);
    ++launches;
}

// transform + prefilter; leaves d_args describing the cloud insert_pointcloud proper starts from
void Map::enqueue_ingest() {
    const int grid = std::max(1, std::min(ceil_div(caps.points, kThreads), num_sms * 8));
    k_tf_points<<<grid, kThreads, 0, stream>>>(d_args, stage_cloud.as<float4>());
    ++launches;
    const int pre = ingest_pre_ds > 0 ? 1 : 0;
    if (pre) enqueue_voxel_grid(0);               // stage_cloud --(leaf = pre_ds)--> hits_ds, count in n_ds_hits
    // (single CTA: the copy must not race with the counter reset, and a prefiltered cloud is small)
    k_ingest_commit<<<1, 1024, 0, stream>>>(d_args, d_cnt, d_mm, hits_ds.as<float4>(), pre);
    ++launches;
}

void Map::enqueue_scan(int mode) {
    const bool frontend_only = mode == 1;
    launches = 0;
    // BGK / GP: the sort-free cooperative pipeline (frontend_fused.cu) unless it was switched off
    const bool fused = fused_applicable(mode);
    if (fused) enqueue_fused_begin(1);
    else { k_scan_begin<<<1, 32, 0, stream>>>(d_cnt, d_mm); ++launches; }
    if (mode == 3) enqueue_ingest();
    if (mode == 2) enqueue_training_data();
    else if (hp.method == LA3DM_BGKL) enqueue_frontend_bgkl();
    else if (hp.method == LA3DM_BGKLV) enqueue_frontend_lv();
    else if (fused) enqueue_fused(0);
    else enqueue_frontend_bgk();
    if (!frontend_only && hp.method == LA3DM_BGKLV) {
        enqueue_lv();
    } else if (!frontend_only) {
        if (fused) {
            enqueue_fused(1);
            if (hp.method == LA3DM_GP) enqueue_gp_sizes();   // capacity checks before the plan touches the map
            enqueue_fused(2);
        } else enqueue_binning();
        if (hp.method == LA3DM_GP) enqueue_gp();
        else if (hp.method == LA3DM_BGKL) enqueue_predict_bgkl();
        else enqueue_predict();
        if (hp.method == LA3DM_BGK) enqueue_peer_wait();
    }
    // a launch that was refused (bad configuration, missing function attribute) must surface as LA3DM_ERR_CUDA, not as
    // a scan that silently did nothing
    LA3DM_CUDA(cudaGetLastError());
}

// temp storage of the largest radix sort the scan issues
size_t radix_sort_temp_bytes(unsigned int items) {
    cub::DoubleBuffer<unsigned int> dk(nullptr, nullptr), dv(nullptr, nullptr);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int) items, 0, 32, nullptr);
    return tmp;
}

}  // namespace la3dm_b200
