// la3dm_b200 -- BGKLVOctoMap front-end: get_training_data / beam_sample of the "-LV" variant
// (src/bgklvoctomap/bgklvoctomap.cpp:303-423, 439-462).  Per downsampled hit p:
//   * the hit itself as a training point (label 1) only when max_range > 0 and |p - o| < max_range (:323-329);
//   * a free ray o -> o + n l, l = |p - o| - ell sqrt(2) (or max_range - ell sqrt(2) beyond range), SHORTENED against
//     every other hit that lies near it (:341-386: one pass over all hits in order, l updated as it goes -- O(hits^2)),
//     dropped if it is a short downward ray (:389-391), its start moved ell away from the sensor (:394-403);
//   * markers along the ray: its start and samples from its end back towards the start every free_resolution (:405-414).
// The arithmetic mixes float and double exactly like the reference (double l / offset / influence, float points).
#include "engine.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kHitTile = 256;

// point3f::norm() of (a - b): float differences, float sum of squares left to right, double sqrt (point3f.h:207-214)
__device__ __forceinline__ double norm_diff(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return sqrt((double) (dx * dx + dy * dy + dz * dz));
}

// |p_j - o| for every downsampled hit (used by both the range test :345-348 and dist2 :357 -- the squares are the same)
__global__ void k_lv_ranges(const float4 *__restrict__ hits, const ScanCounters *__restrict__ c,
                            const ScanArgs *__restrict__ A, double *range) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (c->overflow || i >= c->n_ds_hits) return;
    const float4 p = hits[i];
    range[i] = norm_diff(p.x, p.y, p.z, A->ox, A->oy, A->oz);
}

struct LvRay {
    float fo[3];            // free_origin
    float fe[3];            // free_endpt
    unsigned int n_samples; // beam samples between them
    unsigned int flags;     // bit 0: the hit is a training point, bit 1: the ray is kept
};

// One WARP per hit: the whole of the per-hit body of get_training_data except the emission.  The pass over all other
// hits (:341-386) is O(hits^2); its candidate test and the projection onto the ray only depend on the UNshortened ray,
// so the 32 lanes evaluate 32 hits at a time and only the hits that would shorten the ray (within `influence` of it)
// are then applied one after the other in hit order -- the same sequence of updates of `l` as upstream's loop.
// CTAs of 8 warps = 8 hits each, so that even a 2 000-hit scan covers every SM; the per-tile sums k_lv_fill expects
// come from k_lv_tile_sums.
constexpr int kRayThreads = 256;

__global__ void __launch_bounds__(kRayThreads)
k_lv_rays(const float4 *__restrict__ hits, const double *__restrict__ range, const ScanCounters *__restrict__ c,
          const ScanArgs *__restrict__ A, const DevParams *__restrict__ P, LvRay *out) {
    const unsigned int H = c->overflow ? 0u : c->n_ds_hits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kPerWarp = 1;
    const float ox = A->ox, oy = A->oy, oz = A->oz;
    const float max_range = A->max_range;
    const double offset = (double) P->ell * pow(2.0, 0.5);       // :314
    const double influence = (double) P->ell;                    // :315
    for (int hh = 0; hh < kPerWarp; ++hh) {
        const unsigned int i = blockIdx.x * (kRayThreads / 32) + warp * kPerWarp + hh;
        if (i >= H) break;                                           // warp-uniform
        const float4 p = hits[i];
        double l = range[i];                                         // :318
        const float nx = (float) ((double) (p.x - ox) / l), ny = (float) ((double) (p.y - oy) / l),
                    nz = (float) ((double) (p.z - oz) / l);          // :319-321
        LvRay r;
        r.flags = 0;
        r.n_samples = 0;
        if (max_range > 0) {
            if (l < (double) max_range) {
                const float dx = p.x - ox, dy = p.y - oy, dz = p.z - oz;
                l = (double) (float) sqrt((double) (dx * dx + dy * dy + dz * dz));   // :326
                l = l - offset;
                r.flags |= 1u;                                       // :328-329
            } else {
                l = (double) max_range - offset;                     // :333
            }
        }
        float npz = p.z;                                             // nearest_point (only z is used)
        // free_endpt (:338): float + float * double -> double, narrowed by the point3f constructor
        const float ex0 = (float) ((double) ox + (double) nx * l), ey0 = (float) ((double) oy + (double) ny * l),
                    ez0 = (float) ((double) oz + (double) nz * l);
        const double l0 = l;                                         // the nearby test uses the unshortened length
        const float lvx = ex0 - ox, lvy = ey0 - oy, lvz = ez0 - oz;   // line_vec (:371)
        const double lv_norm = sqrt((double) (lvx * lvx + lvy * lvy + lvz * lvz));
        const double lv_norm2 = pow(lv_norm, 2);
        const bool high = (double) p.z > (offset + (double) oz);      // first half of the floor test (:351)
        for (unsigned int j0 = 0; j0 < H; j0 += 32) {
            const unsigned int j = j0 + (unsigned int) lane;
            bool shortens = false;
            double b = 0.0;
            float qz_hit = 0.f;
            if (j < H) {
                const float4 q = hits[j];
                const double rj = range[j];
                bool cand = !(max_range > 0 && rj > (double) max_range);          // :345-348
                cand = cand && !(high && (double) q.z < (double) oz + influence);   // :351-353
                if (cand) {
                    const double dist1 = norm_diff(ex0, ey0, ez0, q.x, q.y, q.z);   // :355
                    cand = dist1 < influence || (dist1 < l0 && rj < l0);            // :359-365 (dist2 == range of q)
                }
                if (cand) {
                    // projection of the hit onto the ray (:372-384); whether it is applied depends on the current l
                    const float vx = q.x - ox, vy = q.y - oy, vz = q.z - oz;
                    b = (double) (vx * lvx + vy * lvy + vz * lvz);
                    const float s = (float) (b / lv_norm2);                   // operator*(float)
                    const float qx = ox + lvx * s, qy = oy + lvy * s, qz = oz + lvz * s;
                    shortens = norm_diff(q.x, q.y, q.z, qx, qy, qz) < influence;
                    qz_hit = q.z;
                }
            }
            unsigned int todo = __ballot_sync(0xffffffffu, shortens);
            while (todo) {                                           // in hit order, like the loop upstream
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const double bj = __shfl_sync(0xffffffffu, b, src);
                const float zj = __shfl_sync(0xffffffffu, qz_hit, src);
                if (bj > pow(l, 2)) continue;                        // :376
                npz = zj;
                l = bj / lv_norm;                                    // :385
            }
        }
        const bool drop = l < (double) max_range / 5.0 && l / (offset - (double) npz) > 0;   // :389-391
        if (!drop) {
            r.flags |= 2u;
            r.fe[0] = (float) ((double) ox + (double) nx * l);
            r.fe[1] = (float) ((double) oy + (double) ny * l);
            r.fe[2] = (float) ((double) oz + (double) nz * l);
            const double mu = 1.0;
            if (l > influence * mu) {                                // :398-400
                r.fo[0] = (float) ((double) ox + (double) nx * influence * mu);
                r.fo[1] = (float) ((double) oy + (double) ny * influence * mu);
                r.fo[2] = (float) ((double) oz + (double) nz * influence * mu);
            } else {
                r.fo[0] = r.fe[0]; r.fo[1] = r.fe[1]; r.fo[2] = r.fe[2];
            }
            // beam_sample(free_endpt, free_origin) (:439-462): d = l'; while (d > 0) { emit; d -= fr }
            const float bx = r.fe[0] - r.fo[0], by = r.fe[1] - r.fo[1], bz = r.fe[2] - r.fo[2];
            const float lb = (float) sqrt((double) (bx * bx + by * by + bz * bz));
            float d = lb;
            unsigned int ns = 0;
            while (d > 0.0) { ++ns; const float nd = d - A->fr; if (nd == d) break; d = nd; }
            r.n_samples = ns;
        } else {
            r.fo[0] = r.fo[1] = r.fo[2] = r.fe[0] = r.fe[1] = r.fe[2] = 0.f;
        }
        if (lane == 0) out[i] = r;
    }
}

// tile_sums[tile] = (rays kept << 32) | training entries of the tile's kHitTile hits
__global__ void __launch_bounds__(kHitTile)
k_lv_tile_sums(const LvRay *__restrict__ info, const ScanCounters *__restrict__ c, unsigned long long *tile_sums) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const unsigned int H = c->overflow ? 0u : c->n_ds_hits;
    const unsigned int i = blockIdx.x * kHitTile + threadIdx.x;
    unsigned long long acc = 0;
    if (i < H) {
        const unsigned int flags = info[i].flags, ns = info[i].n_samples;
        const unsigned int entries = (flags & 1u) + ((flags & 2u) ? 1u + ns : 0u);
        acc = ((unsigned long long) ((flags >> 1) & 1u) << 32) | (unsigned long long) entries;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_sum;
}

// xy / ray_of / rays / ray_first in the reference's push order
__global__ void __launch_bounds__(kHitTile)
k_lv_fill(const float4 *__restrict__ hits, ScanCounters *c, const ScanArgs *__restrict__ A,
          const LvRay *__restrict__ info, const unsigned long long *__restrict__ tile_sums, unsigned int n_tiles,
          float4 *xy, int *ray_of, float4 *rays, unsigned int *ray_first, unsigned int train_cap,
          unsigned int *mm_xy) {
    __shared__ unsigned long long smem[66];
    __shared__ unsigned long long s_pos[kHitTile];
    const unsigned int H = c->overflow ? 0u : c->n_ds_hits;
    unsigned long long prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    const unsigned int n_rays = (unsigned int) (total >> 32), n_train = (unsigned int) (total & 0xFFFFFFFFull);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->n_raw_frees = n_train;          // reported back so that the host can size the training set
        c->n_train = n_train;
        c->n_frees = n_rays;               // LV: number of rays
        if (n_train > train_cap) atomicOr(&c->overflow, OVF_RAW);
    }
    if (n_train > train_cap) return;
    const unsigned int i = blockIdx.x * kHitTile + threadIdx.x;
    LvRay r;
    r.flags = 0; r.n_samples = 0;
    if (i < H) r = info[i];
    const unsigned int entries = (r.flags & 1u) + ((r.flags & 2u) ? 1u + r.n_samples : 0u);
    unsigned long long cta_total;
    const unsigned long long mine = ((unsigned long long) ((r.flags >> 1) & 1u) << 32) | (unsigned long long) entries;
    s_pos[threadIdx.x] = prefix + block_exclusive_scan(mine, smem, cta_total);
    __syncthreads();
    // hits kept as training points are counted by the first thread of the grid afterwards (n_hits is informational)
    const float fr = A->fr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-mn[0], -mn[0], -mn[0]};
    bool any = false;
    unsigned int my_hits = 0;
    for (int h = warp; h < kHitTile; h += kHitTile / 32) {
        const unsigned int gi = blockIdx.x * kHitTile + h;
        if (gi >= H) break;
        const LvRay q = info[gi];
        const unsigned int ne = (q.flags & 1u) + ((q.flags & 2u) ? 1u + q.n_samples : 0u);
        if (!ne) continue;
        const unsigned long long pos = s_pos[h];
        const unsigned int ray = (unsigned int) (pos >> 32);
        unsigned int at = (unsigned int) (pos & 0xFFFFFFFFull);
        if (q.flags & 1u) {
            if (lane == 0) {
                const float4 p = hits[gi];
                xy[at] = make_float4(p.x, p.y, p.z, 1.0f);            // :328
                ray_of[at] = -1;
                mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
                mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
                mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
                any = true;
                ++my_hits;
            }
            ++at;
        }
        if (!(q.flags & 2u)) continue;
        const float bx = q.fe[0] - q.fo[0], by = q.fe[1] - q.fo[1], bz = q.fe[2] - q.fo[2];
        const float lb = (float) sqrt((double) (bx * bx + by * by + bz * bz));
        const float ux = bx / lb, uy = by / lb, uz = bz / lb;
        for (unsigned int e = lane; e < 1u + q.n_samples; e += 32) {
            float4 pt;
            if (e == 0) pt = make_float4(q.fo[0], q.fo[1], q.fo[2], 0.0f);                     // :408
            else {
                float d = lb;                                                                  // :457-461
                for (unsigned int k = 1; k < e; ++k) d -= fr;
                pt = make_float4(q.fo[0] + ux * d, q.fo[1] + uy * d, q.fo[2] + uz * d, 0.0f);
            }
            xy[at + e] = pt;
            ray_of[at + e] = (int) ray;
            mn[0] = fminf(mn[0], pt.x); mx[0] = fmaxf(mx[0], pt.x);
            mn[1] = fminf(mn[1], pt.y); mx[1] = fmaxf(mx[1], pt.y);
            mn[2] = fminf(mn[2], pt.z); mx[2] = fmaxf(mx[2], pt.z);
            any = true;
        }
        if (lane == 0) {
            rays[2 * (size_t) ray] = make_float4(q.fo[0], q.fo[1], q.fo[2], 0.0f);              // :416-417
            rays[2 * (size_t) ray + 1] = make_float4(q.fe[0], q.fe[1], q.fe[2], 0.0f);
            ray_first[ray] = at;
        }
    }
    if (my_hits) atomicAdd(&c->n_hits, my_hits);
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

// On completion: xy[0..n_train) in the reference's push order, ray_of[], rays[2 * n_rays], ray_first[], counters set.
void Map::enqueue_frontend_lv() {
    enqueue_voxel_grid(0);
    const int grid = ceil_div(caps.points, 256);
    k_lv_ranges<<<grid, 256, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, lv_range.as<double>());
    const int n_tiles = ceil_div(caps.points, kHitTile);
    unsigned long long *tile_sums = tiles.as<unsigned long long>();
    k_lv_rays<<<ceil_div(caps.points, kRayThreads / 32), kRayThreads, 0, stream>>>(
        hits_ds.as<float4>(), lv_range.as<double>(), d_cnt, d_args, d_params, lv_info.as<LvRay>());
    k_lv_tile_sums<<<n_tiles, kHitTile, 0, stream>>>(lv_info.as<LvRay>(), d_cnt, tile_sums);
    ++launches;
    k_lv_fill<<<n_tiles, kHitTile, 0, stream>>>(hits_ds.as<float4>(), d_cnt, d_args, lv_info.as<LvRay>(), tile_sums,
                                                (unsigned int) n_tiles, xy.as<float4>(), ray_of.as<int>(),
                                                rays.as<float4>(), ray_first.as<unsigned int>(), caps.raw, d_mm + 12);
    launches += 3;
}

size_t lv_ray_info_bytes() { return sizeof(LvRay); }

}  // namespace la3dm_b200
