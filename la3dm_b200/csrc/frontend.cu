// la3dm_b200 -- GPU front-end of insert_pointcloud: voxel-grid downsampling, range filter and free-space beam
// sampling.  Replaces get_training_data / downsample / beam_sample
// (src/bgkoctomap/bgkoctomap.cpp:383-458; PCL's VoxelGrid is restated from its published algorithm).
//
// All arithmetic mirrors the reference's fp32 evaluation order (library built with -fmad=false).
#include <cub/cub.cuh>

#include "engine.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kThreads = 256;

// ---- voxel grid ----------------------------------------------------------------------------------------------------
__global__ void k_mm_init(unsigned int *mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = 0xFFFFFFFFu;          // min (flipped)
    else if (threadIdx.x < 6) mm[threadIdx.x] = 0u;              // max (flipped)
}

// min / max over n points (getMinMax3D); n may live on the device (d_n) for the training-set bbox
__global__ void k_minmax(const float *__restrict__ in, int stride_f, unsigned int n_host,
                         const unsigned int *__restrict__ d_n, unsigned int *mm) {
    const unsigned int n = d_n ? *d_n : n_host;
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *p = in + (size_t) i * stride_f;
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = p[a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
    }
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

struct VGFrame {
    int min_b[3];
    int mul1, mul2;
    bool passthrough;
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
    return f;
}

__global__ void k_vg_keys(const float *__restrict__ in, int stride_f, unsigned int n, float inv,
                          const unsigned int *__restrict__ mm, unsigned int *keys, unsigned int *vals,
                          unsigned int *d_passthrough) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const VGFrame f = vg_frame(mm, inv);
    if (i == 0) *d_passthrough = f.passthrough ? 1u : 0u;
    unsigned int key;
    if (f.passthrough) key = i;
    else {
        const float *p = in + (size_t) i * stride_f;
        const int i0 = (int) (floorf(p[0] * inv) - (float) f.min_b[0]);
        const int i1 = (int) (floorf(p[1] * inv) - (float) f.min_b[1]);
        const int i2 = (int) (floorf(p[2] * inv) - (float) f.min_b[2]);
        key = (unsigned int) (i0 + i1 * f.mul1 + i2 * f.mul2);
    }
    keys[i] = key;
    vals[i] = i;
}

__global__ void k_heads(const unsigned int *__restrict__ keys, unsigned int n, unsigned int *flags) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// run_start[rank] = i for every head; run_start[total] = n; *d_total = total
__global__ void k_runs(const unsigned int *__restrict__ flags, const unsigned int *__restrict__ ranks, unsigned int n,
                       unsigned int *run_start, unsigned int *d_total) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) run_start[ranks[i]] = i;
    if (i == n - 1) {
        const unsigned int total = ranks[i] + flags[i];
        run_start[total] = n;
        *d_total = total;
    }
}

// one thread per voxel: sequential fp32 sum in ascending input order, then / float(count)  (CentroidPoint)
__global__ void k_vg_centroid(const float *__restrict__ in, int stride_f, const unsigned int *__restrict__ vals,
                              const unsigned int *__restrict__ run_start, const unsigned int *__restrict__ d_total,
                              float4 *out, const unsigned int *__restrict__ d_out_off, float label) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *d_total) return;
    const unsigned int first = run_start[r], last = run_start[r + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (unsigned int li = first; li < last; ++li) {
        const float *p = in + (size_t) vals[li] * stride_f;
        sx += p[0]; sy += p[1]; sz += p[2];
    }
    const float cnt = (float) (last - first);
    const unsigned int off = d_out_off ? *d_out_off : 0u;
    out[off + r] = make_float4(sx / cnt, sy / cnt, sz / cnt, label);
}

// ds_resolution < 0: downsample() is the identity (src/bgkoctomap/bgkoctomap.cpp:420-423)
__global__ void k_copy_points(const float *__restrict__ in, int stride_f, unsigned int n, float4 *out,
                              const unsigned int *__restrict__ d_out_off, float label, unsigned int *d_count) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *d_count = n;
    if (i >= n) return;
    const float *p = in + (size_t) i * stride_f;
    const unsigned int off = d_out_off ? *d_out_off : 0u;
    out[off + i] = make_float4(p[0], p[1], p[2], label);
}

// ---- range filter + beam sampling ---------------------------------------------------------------------------------
struct Beam {
    float l, nx, ny, nz;
};

// beam_sample preamble (src/bgkoctomap/bgkoctomap.cpp:437-449)
__device__ inline Beam beam_of(const float4 h, const float3 o) {
    Beam b;
    const float dx = h.x - o.x, dy = h.y - o.y, dz = h.z - o.z;
    const float s = dx * dx + dy * dy + dz * dz;
    b.l = (float) sqrt((double) s);
    b.nx = dx / b.l; b.ny = dy / b.l; b.nz = dz / b.l;
    return b;
}

// per downsampled hit: keep flag (range filter :394-398) and number of free points it emits
// (origin once per kept hit :404, samples d = fr, 2fr.. < l with fp32 accumulation :451-455, tail sample :456-457)
__global__ void k_hit_count(const float4 *__restrict__ hits, const unsigned int *__restrict__ d_n, float3 o, float fr,
                            float max_range, unsigned long long *packed) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n = *d_n;
    if (i >= n) return;
    const float4 h = hits[i];
    const float dx = h.x - o.x, dy = h.y - o.y, dz = h.z - o.z;
    const float s = dx * dx + dy * dy + dz * dz;
    if (max_range > 0) {
        const double l = sqrt((double) s);                 // point3f::norm() (point3f.h:207-214)
        if (l > (double) max_range) { packed[i] = 0ull; return; }
    }
    const float l = (float) sqrt((double) s);
    unsigned int cnt = 1;                                   // the origin
    float d = fr;
    while (d < l) { ++cnt; d += fr; }
    if (l > fr) ++cnt;
    packed[i] = (1ull << 32) | (unsigned long long) cnt;
}

__global__ void k_hit_fill(const float4 *__restrict__ hits, const unsigned int *__restrict__ d_n, float3 o, float fr,
                           const unsigned long long *__restrict__ packed,
                           const unsigned long long *__restrict__ offs, float4 *xy, float4 *frees,
                           ScanCounters *cnt) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n = *d_n;
    if (i >= n) return;
    const unsigned long long pk = packed[i], of = offs[i];
    if (i == n - 1) {
        const unsigned long long tot = of + pk;
        cnt->n_hits = (unsigned int) (tot >> 32);
        cnt->n_raw_frees = (unsigned int) (tot & 0xFFFFFFFFull);
    }
    if (pk == 0ull) return;
    const float4 h = hits[i];
    xy[(unsigned int) (of >> 32)] = make_float4(h.x, h.y, h.z, 1.0f);                  // :399
    float4 *out = frees + (unsigned int) (of & 0xFFFFFFFFull);
    *out++ = make_float4(o.x, o.y, o.z, 0.f);                                          // :404
    const Beam b = beam_of(h, o);
    float d = fr;
    while (d < b.l) {
        *out++ = make_float4(o.x + b.nx * d, o.y + b.ny * d, o.z + b.nz * d, 0.f);      // :453
        d += fr;
    }
    if (b.l > fr) {
        const float e = b.l - fr;
        *out++ = make_float4(o.x + b.nx * e, o.y + b.ny * e, o.z + b.nz * e, 0.f);      // :457
    }
}

__global__ void k_hit_empty(ScanCounters *cnt) { cnt->n_hits = 0; cnt->n_raw_frees = 0; }

__global__ void k_finish_train(ScanCounters *cnt) { cnt->n_train = cnt->n_hits + cnt->n_frees; }

}  // namespace

// min/max of a point array into mm[0..2] / mm[3..5] (order-preserving uint encoding)
void Map::minmax_points(const float *d_in, int stride_f, unsigned int n_host, const unsigned int *d_n,
                        unsigned int *mm, unsigned int n_upper) {
    k_mm_init<<<1, 32, 0, stream>>>(mm);
    const int grid = std::max(1, std::min(ceil_div(n_upper, kThreads), num_sms * 8));
    k_minmax<<<grid, kThreads, 0, stream>>>(d_in, stride_f, n_host, d_n, mm);
    launches += 2;
}

// pcl::VoxelGrid on n points (host-known upper bound n; all n are valid).  Returns nothing on the host; the output
// count goes to *d_count.  `which` selects the passthrough flag slot.
unsigned int Map::voxel_grid(const float *d_in, int stride_f, unsigned int n, float leaf, float4 *d_out,
                             const unsigned int *d_out_off, float label, unsigned int *d_count, int which) {
    if (n == 0) {
        LA3DM_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned int), stream));
        return 0;
    }
    const int grid = ceil_div(n, kThreads);
    if (leaf < 0) {
        k_copy_points<<<grid, kThreads, 0, stream>>>(d_in, stride_f, n, d_out, d_out_off, label, d_count);
        ++launches;
        return n;
    }
    const float inv = 1.0f / leaf;
    unsigned int *mm = d_mm + 6 * which;
    minmax_points(d_in, stride_f, n, nullptr, mm, n);
    for (int i = 0; i < 2; ++i) { sort_keys[i].reserve((size_t) n * 4, stream); sort_vals[i].reserve((size_t) n * 4, stream); }
    flags.reserve((size_t) n * 4, stream);
    ranks.reserve((size_t) n * 4, stream);
    run_start.reserve((size_t) (n + 1) * 4, stream);
    k_vg_keys<<<grid, kThreads, 0, stream>>>(d_in, stride_f, n, inv, mm, sort_keys[0].as<unsigned int>(),
                                             sort_vals[0].as<unsigned int>(), &d_cnt->vg_passthrough[which]);
    cub::DoubleBuffer<unsigned int> dk(sort_keys[0].as<unsigned int>(), sort_keys[1].as<unsigned int>());
    cub::DoubleBuffer<unsigned int> dv(sort_vals[0].as<unsigned int>(), sort_vals[1].as<unsigned int>());
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int) n, 0, 32, stream);
    size_t tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, flags.as<unsigned int>(), ranks.as<unsigned int>(), (int) n, stream);
    cub_tmp.reserve(std::max(tmp, tmp2), stream);
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, dk, dv, (int) n, 0, 32, stream));
    k_heads<<<grid, kThreads, 0, stream>>>(dk.Current(), n, flags.as<unsigned int>());
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp2, flags.as<unsigned int>(), ranks.as<unsigned int>(),
                                             (int) n, stream));
    k_runs<<<grid, kThreads, 0, stream>>>(flags.as<unsigned int>(), ranks.as<unsigned int>(), n,
                                          run_start.as<unsigned int>(), d_count);
    k_vg_centroid<<<grid, kThreads, 0, stream>>>(d_in, stride_f, dv.Current(), run_start.as<unsigned int>(), d_count,
                                                 d_out, d_out_off, label);
    launches += 5 + 5;   // ours + CUB radix sort (histogram, scan, 4 onesweep passes ~ 4) + scan
    return n;
}

// BGK / GP front-end.  On return (stream-ordered): xy[0..n_train) = hits (label 1) then free centroids (label 0 / -1),
// d_cnt->{n_ds_hits, n_hits, n_raw_frees, n_frees, n_train} are set and h_cnt holds n_hits / n_raw_frees.
void Map::frontend_bgk(const float *d_xyz, unsigned int n, int stride_f, float3 origin, float ds, float fr,
                       float max_range) {
    const float free_label = hp.method == LA3DM_GP ? -1.0f : 0.0f;   // src/gpoctomap/gpoctomap.cpp:399
    hits_ds.reserve((size_t) std::max(n, 1u) * sizeof(float4), stream);
    voxel_grid(d_xyz, stride_f, n, ds, hits_ds.as<float4>(), nullptr, 1.0f, &d_cnt->n_ds_hits, 0);

    if (n == 0) {
        k_hit_empty<<<1, 1, 0, stream>>>(d_cnt);
        ++launches;
    } else {
        scan64.reserve((size_t) n * 16, stream);
        unsigned long long *packed = scan64.as<unsigned long long>();
        unsigned long long *offs = packed + n;
        const int grid = ceil_div(n, kThreads);
        // entries >= n_ds_hits must scan as zero
        LA3DM_CUDA(cudaMemsetAsync(packed, 0, (size_t) n * 8, stream));
        k_hit_count<<<grid, kThreads, 0, stream>>>(hits_ds.as<float4>(), &d_cnt->n_ds_hits, origin, fr, max_range,
                                                   packed);
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, packed, offs, (int) n, stream);
        cub_tmp.reserve(tmp, stream);
        LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp, packed, offs, (int) n, stream));
        launches += 3;
        // sync #1: sizes of the free cloud
        // (count kernel wrote totals? no: totals come from the fill kernel, which needs the buffers.  Read the last
        //  scan element + last packed value instead.)
        unsigned long long last[2];
        d2h_bytes += 4;
        LA3DM_CUDA(cudaMemcpyAsync(&h_cnt->n_ds_hits, &d_cnt->n_ds_hits, sizeof(unsigned int), cudaMemcpyDeviceToHost,
                                   stream));
        LA3DM_CUDA(cudaStreamSynchronize(stream));
        const unsigned int m = h_cnt->n_ds_hits;
        if (m == 0) {
            k_hit_empty<<<1, 1, 0, stream>>>(d_cnt);
            ++launches;
            h_cnt->n_hits = 0; h_cnt->n_raw_frees = 0;
        } else {
            d2h_bytes += 16;
            LA3DM_CUDA(cudaMemcpyAsync(&last[0], packed + (m - 1), 8, cudaMemcpyDeviceToHost, stream));
            LA3DM_CUDA(cudaMemcpyAsync(&last[1], offs + (m - 1), 8, cudaMemcpyDeviceToHost, stream));
            LA3DM_CUDA(cudaStreamSynchronize(stream));
            const unsigned long long tot = last[0] + last[1];
            h_cnt->n_hits = (unsigned int) (tot >> 32);
            h_cnt->n_raw_frees = (unsigned int) (tot & 0xFFFFFFFFull);
            xy.reserve(((size_t) h_cnt->n_hits + h_cnt->n_raw_frees + 1) * sizeof(float4), stream);
            frees_raw.reserve(((size_t) h_cnt->n_raw_frees + 1) * sizeof(float4), stream);
            k_hit_fill<<<ceil_div(m, kThreads), kThreads, 0, stream>>>(hits_ds.as<float4>(), &d_cnt->n_ds_hits, origin,
                                                                      fr, packed, offs, xy.as<float4>(),
                                                                      frees_raw.as<float4>(), d_cnt);
            ++launches;
        }
    }
    // second voxel grid over the free cloud; centroids land behind the hits in xy
    voxel_grid(frees_raw.as<float>(), 4, h_cnt->n_raw_frees, ds, xy.as<float4>(), &d_cnt->n_hits, free_label,
               &d_cnt->n_frees, 1);
    k_finish_train<<<1, 1, 0, stream>>>(d_cnt);
    ++launches;
}

}  // namespace la3dm_b200
