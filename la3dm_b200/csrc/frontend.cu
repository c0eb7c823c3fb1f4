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
    if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
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
// runs longer than kLongRun are handed to k_vg_long
template <int W>
__global__ void k_vg_centroid(const ScanArgs *__restrict__ A, ScanCounters *c, const float4 *__restrict__ frees_raw,
                              const unsigned int *__restrict__ vals, const unsigned int *__restrict__ run_start,
                              const unsigned int *__restrict__ d_total, float4 *out, unsigned int *long_list) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (c->overflow || r >= *d_total) return;
    const float *in; int stride; unsigned int n;
    vg_source<W>(A, c, frees_raw, in, stride, n);
    const unsigned int first = run_start[r], last = run_start[r + 1];
    if (last - first > (unsigned int) kLongRun) {
        const unsigned int slot = atomicAdd(&c->n_long_runs[W], 1u);
        if (slot < (unsigned int) kMaxLongRuns) { long_list[slot] = r; return; }
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (unsigned int li = first; li < last; ++li) {
        const float *p = in + (size_t) vals[li] * stride;
        sx += p[0]; sy += p[1]; sz += p[2];
    }
    const float cnt = (float) (last - first);
    const unsigned int off = W == 0 ? 0u : c->n_hits;
    const float label = W == 0 ? 1.0f : A->free_label;
    out[off + r] = make_float4(sx / cnt, sy / cnt, sz / cnt, label);
}

// long runs: a CTA stages the run through shared memory chunk by chunk; three lanes (x, y, z) add the chunk in
// order.  A chunk made of one repeated point (the sensor origin, pushed once per hit) is added with add_repeat.
constexpr int kLongThreads = 512;
constexpr int kLongChunk = 2048;

template <int W>
__global__ void __launch_bounds__(kLongThreads)
k_vg_long(const ScanArgs *__restrict__ A, const ScanCounters *__restrict__ c, const float4 *__restrict__ frees_raw,
          const unsigned int *__restrict__ vals, const unsigned int *__restrict__ run_start, float4 *out,
          const unsigned int *__restrict__ long_list) {
    __shared__ float sc[3][kLongChunk];
    if (c->overflow) return;
    const unsigned int nl = min(c->n_long_runs[W], (unsigned int) kMaxLongRuns);
    const float *in; int stride; unsigned int n;
    vg_source<W>(A, c, frees_raw, in, stride, n);
    const unsigned int off = W == 0 ? 0u : c->n_hits;
    const float label = W == 0 ? 1.0f : A->free_label;
    for (unsigned int idx = blockIdx.x; idx < nl; idx += gridDim.x) {
        const unsigned int r = long_list[idx];
        const unsigned int first = run_start[r], last = run_start[r + 1];
        float acc = 0.f;
        for (unsigned int c0 = first; c0 < last; c0 += kLongChunk) {
            const unsigned int m = min((unsigned int) kLongChunk, last - c0);
            const float *p0 = in + (size_t) vals[c0] * stride;
            const unsigned int r0 = __float_as_uint(p0[0]), r1 = __float_as_uint(p0[1]), r2 = __float_as_uint(p0[2]);
            int same = 1;
            for (unsigned int j = threadIdx.x; j < m; j += kLongThreads) {
                const float *p = in + (size_t) vals[c0 + j] * stride;
                const float x = p[0], y = p[1], z = p[2];
                sc[0][j] = x; sc[1][j] = y; sc[2][j] = z;
                same &= (__float_as_uint(x) == r0) & (__float_as_uint(y) == r1) & (__float_as_uint(z) == r2);
            }
            const int all_same = __syncthreads_and(same);
            if (threadIdx.x < 3) {
                const float *v = sc[threadIdx.x];
                if (all_same) acc = add_repeat(acc, v[0], m);
                else
                    for (unsigned int j = 0; j < m; ++j) acc += v[j];
            }
            __syncthreads();
        }
        float *o = reinterpret_cast<float *>(out + off + r);
        if (threadIdx.x < 3) o[threadIdx.x] = acc / (float) (last - first);
        else if (threadIdx.x == 3) o[3] = label;
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
    unsigned int cnt = 1;                                   // the origin
    float d = fr;
    while (d < l) { ++cnt; d += fr; }
    if (l > fr) ++cnt;
    return cnt;
}

// tile_sums[tile] = (kept hits << 32) | free points of the tile; hit_cnt[i] = per-hit count (0 = dropped)
__global__ void k_hit_count(const float4 *__restrict__ hits, const ScanCounters *__restrict__ c,
                            const ScanArgs *__restrict__ A, unsigned int *hit_cnt, unsigned long long *tile_sums) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const unsigned int n = c->overflow ? 0u : c->n_ds_hits;
    unsigned long long acc = 0;
    // strided over the tile so that the float4 loads coalesce; hit_cnt keeps hit order
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int i = blockIdx.x * kTile + k * kTileThreads + threadIdx.x;
        if (i < n) {
            const unsigned int cnt = hit_free_count(hits[i], A);
            hit_cnt[i] = cnt;
            if (cnt) acc += (1ull << 32) | (unsigned long long) cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_sum;
}

// writes the kept hits (label 1) to xy[0 .. n_hits) and every free point to frees_raw, in the reference's push order
__global__ void k_hit_fill(const float4 *__restrict__ hits, ScanCounters *c, const ScanArgs *__restrict__ A,
                           const unsigned int *__restrict__ hit_cnt, const unsigned long long *__restrict__ tile_sums,
                           unsigned int n_tiles, float4 *xy, float4 *frees, unsigned int raw_cap) {
    __shared__ unsigned long long smem[66];
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
    const unsigned int base = blockIdx.x * kTile + threadIdx.x * kTileItems;
    unsigned long long mine = 0;
    unsigned int cnts[kTileItems];
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int i = base + k;
        cnts[k] = i < n ? hit_cnt[i] : 0u;
        if (cnts[k]) mine += (1ull << 32) | (unsigned long long) cnts[k];
    }
    unsigned long long cta_total;
    unsigned long long pos = prefix + block_exclusive_scan(mine, smem, cta_total);
    const float ox = A->ox, oy = A->oy, oz = A->oz, fr = A->fr;
#pragma unroll 1
    for (int k = 0; k < kTileItems; ++k) {
        if (!cnts[k]) continue;
        const float4 h = hits[base + k];
        xy[(unsigned int) (pos >> 32)] = make_float4(h.x, h.y, h.z, 1.0f);                 // :399
        float4 *out = frees + (unsigned int) (pos & 0xFFFFFFFFull);
        pos += (1ull << 32) | (unsigned long long) cnts[k];
        *out++ = make_float4(ox, oy, oz, 0.f);                                             // :404
        // beam_sample preamble (:437-449)
        const float dx = h.x - ox, dy = h.y - oy, dz = h.z - oz;
        const float l = (float) sqrt((double) (dx * dx + dy * dy + dz * dz));
        const float nx = dx / l, ny = dy / l, nz = dz / l;
        float d = fr;
        while (d < l) {
            *out++ = make_float4(ox + nx * d, oy + ny * d, oz + nz * d, 0.f);               // :453
            d += fr;
        }
        if (l > fr) {
            const float e = l - fr;
            *out++ = make_float4(ox + nx * e, oy + ny * e, oz + nz * e, 0.f);               // :457
        }
    }
}

__global__ void k_finish_train(ScanCounters *c) {
    if (c->overflow) return;
    c->n_train = c->n_hits + c->n_frees;
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
    unsigned int *llist = long_list.as<unsigned int>() + (size_t) which * kMaxLongRuns;
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
    } else {
        k_vg_minmax<1><<<mm_grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, mm);
        k_vg_keys<1><<<grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, mm, dk.Current(), dv.Current(), cap,
                                                    caps.vg_cells);
    }
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, dk, dv, (int) cap, 0, end_bit, stream));
    const unsigned int *ks = dk.Current(), *vs = dv.Current();
    const unsigned int *d_n = which == 0 ? &d_args->n : &d_cnt->n_raw_frees;
    unsigned int *d_total = which == 0 ? &d_cnt->n_ds_hits : &d_cnt->n_frees;
    k_run_count<<<n_tiles, kTileThreads, 0, stream>>>(ks, d_n, cap, tile_sums, d_cnt);
    k_vg_run_place<<<n_tiles, kTileThreads, 0, stream>>>(ks, d_n, cap, tile_sums, (unsigned int) n_tiles, runs,
                                                         d_total, d_cnt);
    const int long_grid = 64;
    if (which == 0) {
        k_vg_centroid<0><<<grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, d_total, hits_ds.as<float4>(),
                                                        llist);
        k_vg_long<0><<<long_grid, kLongThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, hits_ds.as<float4>(), llist);
    } else {
        k_vg_centroid<1><<<grid, kThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, d_total, xy.as<float4>(), llist);
        k_vg_long<1><<<long_grid, kLongThreads, 0, stream>>>(d_args, d_cnt, fr, vs, runs, xy.as<float4>(), llist);
    }
    launches += 6 + 2 + (end_bit + 7) / 8;   // ours + CUB radix sort (histogram, exclusive sum, onesweep passes)
}

// BGK / GP front-end.  On completion (stream-ordered): xy[0..n_train) = hits (label 1) then free centroids (label
// 0 / -1) and d_cnt->{n_ds_hits, n_hits, n_raw_frees, n_frees, n_train} are set.
void Map::enqueue_frontend_bgk() {
    enqueue_voxel_grid(0);
    const int n_tiles = ceil_div(caps.points, kTile);
    unsigned long long *tile_sums = tiles.as<unsigned long long>();
    k_hit_count<<<n_tiles, kTileThreads, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, hit_cnt.as<unsigned int>(),
                                                      tile_sums);
    k_hit_fill<<<n_tiles, kTileThreads, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, hit_cnt.as<unsigned int>(),
                                                     tile_sums, (unsigned int) n_tiles, xy.as<float4>(),
                                                     frees_raw.as<float4>(), caps.raw);
    launches += 2;
    enqueue_voxel_grid(1);
    k_finish_train<<<1, 1, 0, stream>>>(d_cnt);
    ++launches;
}

// The whole scan, stream-ordered, no host synchronisation (this is what the CUDA graph captures).
void Map::enqueue_scan(bool frontend_only) {
    launches = 0;
    k_scan_begin<<<1, 32, 0, stream>>>(d_cnt, d_mm);
    ++launches;
    enqueue_frontend_bgk();
    if (!frontend_only) {
        enqueue_binning();
        enqueue_predict();
    }
}

// temp storage of the largest radix sort the scan issues
size_t radix_sort_temp_bytes(unsigned int items) {
    cub::DoubleBuffer<unsigned int> dk(nullptr, nullptr), dv(nullptr, nullptr);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int) items, 0, 32, nullptr);
    return tmp;
}

}  // namespace la3dm_b200
