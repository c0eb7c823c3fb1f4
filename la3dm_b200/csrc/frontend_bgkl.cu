// la3dm_b200 -- BGKLOctoMap front-end: get_training_data / beam_sample of the line-segment variant
// (src/bgkloctomap/bgkloctomap.cpp:285-344, 360-383).  Per downsampled, in-range hit, in this order:
//   the re-projected hit  origin + n * l          (label 1, ray_idx -1),
//   the origin and the beam samples origin + n' * d, d = l' - fr, l' - 2 fr, ... > 0   (label 0, markers of ray idx),
//   and one ray segment  origin -> origin + n * (l - fr).
// There is no second voxel grid.  Like the BGK front-end it runs without host synchronisation.
#include "engine.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kHitTile = 256;

struct BeamL {
    float l, nx, ny, nz;        // from the downsampled hit (:310-314)
    float ox_, oy_, oz_;        // occ_endpt (:316)
    float l2, mx, my, mz;       // beam_sample's own length / direction, from occ_endpt (:372-376)
};

__device__ inline BeamL beam_l(const float4 h, float ox, float oy, float oz) {
    BeamL b;
    const float dx = h.x - ox, dy = h.y - oy, dz = h.z - oz;
    b.l = (float) sqrt((double) (dx * dx + dy * dy + dz * dz));
    b.nx = dx / b.l; b.ny = dy / b.l; b.nz = dz / b.l;
    b.ox_ = ox + b.nx * b.l; b.oy_ = oy + b.ny * b.l; b.oz_ = oz + b.nz * b.l;
    const float ex = b.ox_ - ox, ey = b.oy_ - oy, ez = b.oz_ - oz;
    b.l2 = (float) sqrt((double) (ex * ex + ey * ey + ez * ez));
    b.mx = ex / b.l2; b.my = ey / b.l2; b.mz = ez / b.l2;
    return b;
}

// training entries a hit emits (0 = dropped by the range filter :304-308): hit + origin + samples
__device__ inline unsigned int hitl_count(const float4 h, const ScanArgs *A) {
    const float dx = h.x - A->ox, dy = h.y - A->oy, dz = h.z - A->oz;
    if (A->max_range > 0) {
        const double l = sqrt((double) (dx * dx + dy * dy + dz * dz));
        if (l > (double) A->max_range) return 0u;
    }
    const BeamL b = beam_l(h, A->ox, A->oy, A->oz);
    unsigned int cnt = 2;
    float d = b.l2 - A->fr;
    while (d > 0.0) { ++cnt; const float nd = d - A->fr; if (nd == d) break; d = nd; }
    return cnt;
}

__global__ void __launch_bounds__(kHitTile)
k_hitl_count(const float4 *__restrict__ hits, const ScanCounters *__restrict__ c, const ScanArgs *__restrict__ A,
             unsigned int *hit_cnt, unsigned long long *tile_sums) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const unsigned int n = c->overflow ? 0u : c->n_ds_hits;
    unsigned long long acc = 0;
    const unsigned int i = blockIdx.x * kHitTile + threadIdx.x;
    if (i < n) {
        const unsigned int cnt = hitl_count(hits[i], A);
        hit_cnt[i] = cnt;
        if (cnt) acc = (1ull << 32) | (unsigned long long) cnt;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_sum;
}

// xy[pos ..] = hit, origin, samples; ray_of[pos ..] = -1, idx, idx, ...; rays[2 idx] = origin, rays[2 idx + 1] = end
__global__ void __launch_bounds__(kHitTile)
k_hitl_fill(const float4 *__restrict__ hits, ScanCounters *c, const ScanArgs *__restrict__ A,
            const unsigned int *__restrict__ hit_cnt, const unsigned long long *__restrict__ tile_sums,
            unsigned int n_tiles, float4 *xy, int *ray_of, float4 *rays, unsigned int train_cap, unsigned int *mm_xy) {
    __shared__ unsigned long long smem[66];
    __shared__ unsigned long long s_pos[kHitTile];
    __shared__ unsigned int s_cnt[kHitTile];
    const unsigned int n = c->overflow ? 0u : c->n_ds_hits;
    unsigned long long prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    const unsigned int n_hits = (unsigned int) (total >> 32), n_train = (unsigned int) (total & 0xFFFFFFFFull);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->n_hits = n_hits;
        c->n_raw_frees = n_train;          // reported back so that the host can size the training set
        c->n_frees = n_train - n_hits;
        c->n_train = n_train;
        if (n_train > train_cap) atomicOr(&c->overflow, OVF_RAW);
    }
    if (n_train > train_cap) return;
    const unsigned int i = blockIdx.x * kHitTile + threadIdx.x;
    const unsigned int cnt = i < n ? hit_cnt[i] : 0u;
    unsigned long long cta_total;
    const unsigned long long mine = cnt ? ((1ull << 32) | (unsigned long long) cnt) : 0ull;
    s_pos[threadIdx.x] = prefix + block_exclusive_scan(mine, smem, cta_total);
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    const float ox = A->ox, oy = A->oy, oz = A->oz, fr = A->fr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-mn[0], -mn[0], -mn[0]};
    bool any = false;
    for (int h = warp; h < kHitTile; h += kHitTile / 32) {
        const unsigned int ch = s_cnt[h];
        if (!ch) continue;
        const unsigned long long pos = s_pos[h];
        const unsigned int ray = (unsigned int) (pos >> 32), at = (unsigned int) (pos & 0xFFFFFFFFull);
        const BeamL b = beam_l(hits[blockIdx.x * kHitTile + h], ox, oy, oz);
        for (unsigned int e = lane; e < ch; e += 32) {
            float4 p;
            int rid = (int) ray;
            if (e == 0) { p = make_float4(b.ox_, b.oy_, b.oz_, 1.0f); rid = -1; }            // :316-318
            else if (e == 1) p = make_float4(ox, oy, oz, 0.0f);                               // :328-330
            else {
                float d = b.l2 - fr;                                                          // :378-382
                for (unsigned int q = 2; q < e; ++q) d -= fr;
                p = make_float4(ox + b.mx * d, oy + b.my * d, oz + b.mz * d, 0.0f);
            }
            xy[at + e] = p;
            ray_of[at + e] = rid;
            mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
            mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
            mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
            any = true;
        }
        if (lane == 0) {
            const float le = b.l - fr;                                                        // :336-339
            rays[2 * (size_t) ray] = make_float4(ox, oy, oz, 0.0f);
            rays[2 * (size_t) ray + 1] = make_float4(ox + b.nx * le, oy + b.ny * le, oz + b.nz * le, 0.0f);
        }
    }
    // bounding box of the training entries (src/bgkloctomap/bgkloctomap.cpp:385-407; markers are points)
    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            for (int o = 16; o > 0; o >>= 1) {
                mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
                mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                atomicMin(&mm_xy[a], float_flip(mn[a]));
                atomicMax(&mm_xy[3 + a], float_flip(mx[a]));
            }
        }
    }
}

}  // namespace

// On completion: xy[0..n_train) in the reference's push order, ray_of[], rays[2 * n_hits], counters set.
void Map::enqueue_frontend_bgkl() {
    enqueue_voxel_grid(0);
    const int n_tiles = ceil_div(caps.points, kHitTile);
    unsigned long long *tile_sums = tiles.as<unsigned long long>();
    k_hitl_count<<<n_tiles, kHitTile, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, hit_cnt.as<unsigned int>(),
                                                   tile_sums);
    k_hitl_fill<<<n_tiles, kHitTile, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, hit_cnt.as<unsigned int>(),
                                                  tile_sums, (unsigned int) n_tiles, xy.as<float4>(), ray_of.as<int>(),
                                                  rays.as<float4>(), caps.raw, d_mm + 12);
    launches += 2;
}

}  // namespace la3dm_b200
