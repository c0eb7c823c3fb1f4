// la3dm_b200 -- BGKLOctoMap: per-block training lists of points and ray segments, and the fused
// predict -> Occupancy::update -> prune.
//
// Replaces the TRAIN bookkeeping and the PREDICT / PRUNE loops of BGKLOctoMap::insert_pointcloud
// (src/bgkloctomap/bgkloctomap.cpp:125-182, 190-262):
//   a block's training set = its hits as degenerate segments + every ray that has a marker in the block, once
//   (ray_keys de-duplication :145-171);
//   BGKLInference::predict / point_to_line_dist / covSparseLine (include/bgkloctomap/bgklinference.h:106-141, 183-197):
//   distance from the voxel centre to the segment, divided by ell AFTER the distance, same sparse kernel;
//   node.update only if kbar > 0.001f (:231).
#include <algorithm>

#include "block_common.cuh"
#include "runs.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kSegTile = 32;

// keep flag of sorted membership i: hits always; a marker only if it is the first of its ray in this block (markers of
// one ray are contiguous in training order, hence contiguous inside the block's sorted range)
__device__ __forceinline__ bool seg_keep(const unsigned int *keys, const unsigned int *vals, const int *ray_of,
                                         unsigned int i) {
    const int rid = ray_of[vals[i]];
    if (rid < 0 || i == 0) return true;
    return keys[i] != keys[i - 1] || ray_of[vals[i - 1]] != rid;
}

__global__ void k_segl_count(const unsigned int *__restrict__ keys, const unsigned int *__restrict__ vals,
                             const int *__restrict__ ray_of, const ScanCounters *__restrict__ c, unsigned int cap,
                             unsigned int *tile_sums) {
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned int n = c->overflow ? 0u : min(c->n_members, cap);
    unsigned int cnt = 0;
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int i = blockIdx.x * kTile + k * kTileThreads + threadIdx.x;
        if (i < n && seg_keep(keys, vals, ray_of, i)) ++cnt;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_cnt;
}

// segs[2 r], segs[2 r + 1] = end points (w of the first = label); seg_start[d] = first segment of data block d
__global__ void k_segl_place(const unsigned int *__restrict__ keys, const unsigned int *__restrict__ vals,
                             const int *__restrict__ ray_of, ScanCounters *c, unsigned int cap,
                             const unsigned int *__restrict__ tile_sums, unsigned int n_tiles,
                             const float4 *__restrict__ xy, const float4 *__restrict__ rays,
                             const unsigned int *__restrict__ cell_db, float4 *segs, unsigned int *seg_start) {
    __shared__ unsigned int smem[66];
    const unsigned int n = c->overflow ? 0u : min(c->n_members, cap);
    unsigned int prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    if (blockIdx.x == 0 && threadIdx.x == 0 && !c->overflow) seg_start[c->n_data_blocks] = total;
    const unsigned int base = blockIdx.x * kTile + threadIdx.x * kTileItems;
    unsigned int flags = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int i = base + k;
        if (i < n && seg_keep(keys, vals, ray_of, i)) { flags |= 1u << k; ++cnt; }
    }
    unsigned int cta_total;
    unsigned int pos = prefix + block_exclusive_scan(cnt, smem, cta_total);
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
        if (!(flags & (1u << k))) continue;
        const unsigned int i = base + k;
        const unsigned int e = vals[i];
        const int rid = ray_of[e];
        if (rid < 0) {                    // a hit: the point twice, label 1 (bgkloctomap.cpp:139-143)
            const float4 p = xy[e];
            segs[2 * (size_t) pos] = make_float4(p.x, p.y, p.z, 1.0f);
            segs[2 * (size_t) pos + 1] = make_float4(p.x, p.y, p.z, 0.0f);
        } else {                          // the marker's ray, label 0 (:160-166)
            segs[2 * (size_t) pos] = rays[2 * (size_t) rid];
            segs[2 * (size_t) pos + 1] = rays[2 * (size_t) rid + 1];
        }
        if (i == 0 || keys[i] != keys[i - 1]) seg_start[cell_db[keys[i]] - 1] = pos;
        ++pos;
    }
}

// covSparse element (bgklinference.h:188-191), d already divided by ell; caller guarantees d <= 1
__device__ __forceinline__ float sparse_kernel(float d, float sf2) {
    const float t = d * 2.0f * 3.1415926f;
    float s, c;
    sincosf(t, &s, &c);
    float k = (((2.0f + c) * (1.0f - d) / 3.0f) + s / (2.0f * 3.1415926f)) * sf2;
    return k < 0.0f ? 0.0f : k;
}

struct UpdateParams {
    float var_thresh, occupied_thresh, free_thresh;
};

__device__ __forceinline__ unsigned char bgk_classify(float a, float b, const UpdateParams &P) {
    const float var = (a * b) / ((a + b) * (a + b) * (a + b + 1.0f));
    if (var > P.var_thresh) return LA3DM_UNKNOWN;
    const float p = a / (a + b);
    return p > P.occupied_thresh ? LA3DM_OCCUPIED : (p < P.free_thresh ? LA3DM_FREE : LA3DM_UNKNOWN);
}

struct SegSmem {
    float4 a[kSegTile];       // p0.xyz, label
    float4 v[kSegTile];       // p1 - p0, w = |p1 - p0|^2 (fp32, summed left to right) or -1 for a degenerate segment
    float4 b[kSegTile];       // p1
};

// point_to_line_dist for one (voxel centre, segment) pair, divided by ell.  c1, c2 are float dot products widened to
// double upstream; comparing and dividing them in fp32 gives the same results (IEEE division: the double quotient
// rounded to float equals the float quotient).  point3f::norm(): squares summed left to right, double sqrt, narrowed.
__device__ __forceinline__ float seg_dist_scaled(float qx, float qy, float qz, const float4 a, const float4 v,
                                                 const float4 p1, float ell) {
    const float px = qx - a.x, py = qy - a.y, pz = qz - a.z;
    float ex = px, ey = py, ez = pz;                       // q - p0
    if (v.w >= 0.f) {                                      // line_len >= EPSILON
        const float c1 = px * v.x + py * v.y + pz * v.z;
        if (c1 > 0) {
            if (v.w <= c1) { ex = qx - p1.x; ey = qy - p1.y; ez = qz - p1.z; }
            else {
                const float b = c1 / v.w;
                const float nx = a.x + v.x * b, ny = a.y + v.y * b, nz = a.z + v.z * b;
                ex = qx - nx; ey = qy - ny; ez = qz - nz;
            }
        }
    }
    const float d = sqrtf(ex * ex + ey * ey + ez * ez);
    return d / ell;
}

constexpr unsigned int kLongSegs = 192;     // neighbour lists longer than this are cut into chunks of kSegChunk segments
constexpr unsigned int kSegChunk = 128;

// segments [begin, end) of one neighbour list added to this lane's (ybar, kbar) in list order: tiles of 32 staged in shared
// memory, culled against the box of the block's leaf centres
__device__ __forceinline__ void seg_accumulate(SegSmem &S, const float4 *__restrict__ src, unsigned int begin, unsigned int end,
                                               int node, float qx, float qy, float qz, float cx, float cy, float cz,
                                               float reach, float cull2, float ell, float sf2, int lane, float &yb, float &kb) {
    for (unsigned int base = begin; base < end; base += kSegTile) {
        const unsigned int m = min((unsigned int) kSegTile, end - base);
        bool keep = false;
        float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sv = sa, sb = sa;
        if ((unsigned int) lane < m) {
            sa = src[2 * (size_t) (base + lane)];
            sb = src[2 * (size_t) (base + lane) + 1];
            sv = make_float4(sb.x - sa.x, sb.y - sa.y, sb.z - sa.z, 0.f);
            const float c2 = sv.x * sv.x + sv.y * sv.y + sv.z * sv.z;
            const float len = (float) sqrt((double) c2);
            sv.w = len < 0.0001f ? -1.0f : c2;                     // EPSILON (bgklinference.h:14)
            // cull: distance between the segment's bounding box and the box of the block's leaf centres
            const float gx = fmaxf(fmaxf(fminf(sa.x, sb.x) - (cx + reach), (cx - reach) - fmaxf(sa.x, sb.x)), 0.f);
            const float gy = fmaxf(fmaxf(fminf(sa.y, sb.y) - (cy + reach), (cy - reach) - fmaxf(sa.y, sb.y)), 0.f);
            const float gz = fmaxf(fmaxf(fminf(sa.z, sb.z) - (cz + reach), (cz - reach) - fmaxf(sa.z, sb.z)), 0.f);
            keep = gx * gx + gy * gy + gz * gz < cull2;
        }
        unsigned int todo = __ballot_sync(0xffffffffu, keep);
        if (!todo) continue;
        __syncwarp();
        S.a[lane] = sa;
        S.v[lane] = sv;
        S.b[lane] = sb;
        __syncwarp();
        while (todo) {
            const int q = __ffs(todo) - 1;
            todo &= todo - 1;
            if (node >= 0) {
                const float4 za = S.a[q], zv = S.v[q], zb = S.b[q];
                const float d = seg_dist_scaled(qx, qy, qz, za, zv, zb, ell);
                if (d < 1.0f) { const float k = sparse_kernel(d, sf2); yb += k * za.w; kb += k; }
            }
        }
    }
}

// test blocks per (k_bgkl_yk, k_bgkl_apply) pair: bounds the (ybar, kbar) buffer at 7 x leaves x 8 B per block
inline unsigned int bgkl_chunk(const DevParams &P) {
    const unsigned int groups = (unsigned int) (P.finest + 31) / 32;
    return std::max(1u, 65536u / groups);
}

// BGKLInference::predict for one (test block, neighbour, group of 32 leaves): a warp per unit, a lane per leaf; the
// neighbour's segments are streamed in tiles of 32 (staged in shared memory, culled against the block's leaf box) and
// every lane adds them to its own (ybar, kbar) in list order.  The 7 x groups units of a test block are independent --
// the block around the sensor, which every ray crosses, is spread over 14 warps at block_depth 3 instead of one -- and
// the sequential part, Occupancy::update neighbour after neighbour, is k_bgkl_apply.
// yk: [block - t0][7][32 groups] (ybar, kbar)
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_bgkl_yk(const NeighbourPlan *__restrict__ plan, const float4 *__restrict__ segs, const long long *__restrict__ keys,
          const unsigned char *__restrict__ pool, const float3 *__restrict__ lut, const DevParams *__restrict__ Pg,
          const ScanArgs *__restrict__ A, const ScanCounters *__restrict__ cnt, unsigned int t0, unsigned int chunk,
          float2 *yk, unsigned int *long_cnt, uint4 *long_units, unsigned int *chunk_unit, unsigned int long_cap,
          unsigned int chunk_cap) {
    __shared__ SegSmem sm[kWarpsPerCta];
    __shared__ DevParams Ps;
    load_params(Ps, Pg);
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SegSmem &S = sm[warp];
    const unsigned int T = cnt->n_test_blocks;
    if (t0 >= T) return;
    const unsigned int t1 = min(T, t0 + chunk);
    const unsigned int gw = blockIdx.x * kWarpsPerCta + warp, n_w = gridDim.x * kWarpsPerCta;
    const float ell = P.ell, sf2 = P.sf2, bs = P.block_size;
    const unsigned int shard_world = (unsigned int) A->shard_world, shard_rank = (unsigned int) A->shard_rank;
    const unsigned int groups = (unsigned int) (P.finest + 31) / 32, per_block = 7u * groups;
    const unsigned int units = (t1 - t0) * per_block;
    const int pruned = P.pruned_state;
    const float reach = 0.5f * (bs - P.resolution) * 1.001f;       // leaf centres lie within this of the block centre
    const float cull2 = ell * ell * (1.0f + 1e-3f);
    for (unsigned int u = gw; u < units; u += n_w) {
        const unsigned int t = t0 + u / per_block, r = u % per_block;
        const int nb = (int) (r / groups), s = (int) (r % groups);
        if (t % shard_world != shard_rank) continue;
        const NeighbourPlan *pl = plan + t;
        const unsigned int n = pl->count[nb];
        if (n == 0) continue;
        const unsigned int slot = pl->slot;
        // this lane's leaf of the group (a fresh block has no record yet: all finest voxels)
        int node = -1;
        const int j = lane + 32 * s;
        if (j < P.finest) {
            if (pl->is_new) node = P.layer_off[P.depth - 1] + j;
            else {
                const unsigned char *rst = pool + (size_t) slot * (size_t) P.rec_bytes + P.st_off;
                int d = P.depth - 1, i = j, shift = 0;
                while (d > 0 && (rst[P.layer_off[d] + i] & 7) == pruned) { --d; i >>= 3; shift += 3; }
                if (((i << shift) == j) && ((rst[P.layer_off[d] + i] & 7) != pruned)) node = P.layer_off[d] + i;
            }
        }
        if (!__any_sync(0xffffffffu, node >= 0)) continue;
        const long long key = keys[slot];
        const float cx = axis_center(key >> 40, bs), cy = axis_center((key >> 20) & 0xFFFFF, bs),
                    cz = axis_center(key & 0xFFFFF, bs);
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (node >= 0) {
            const float3 o = lut[node];
            qx = o.x + cx; qy = o.y + cy; qz = o.z + cz;           // Block::get_loc
        }
        const float4 *src = segs + 2 * (size_t) pl->start[nb];
        if (n > kLongSegs && long_cnt) {
            // a list this long (the block around the sensor holds a marker of EVERY ray) is cut into chunks that other
            // warps sum (k_bgkl_yk_long); the partial sums are added in chunk order (k_bgkl_yk_sum)
            const unsigned int nch = (n + kSegChunk - 1u) / kSegChunk;
            unsigned int first = 0, slot = 0;
            if (lane == 0) {
                slot = atomicAdd(&long_cnt[0], 1u);
                first = atomicAdd(&long_cnt[1], nch);
            }
            slot = __shfl_sync(0xffffffffu, slot, 0);
            first = __shfl_sync(0xffffffffu, first, 0);
            const bool fits = slot < long_cap && first + nch <= chunk_cap;
            if (slot < long_cap && lane == 0) long_units[slot] = fits ? make_uint4(u, first, nch, 0u) : make_uint4(u, 0u, 0u, 0u);
            for (unsigned int q = lane; q < nch; q += 32)
                if (first + q < chunk_cap) chunk_unit[first + q] = fits ? slot : 0xFFFFFFFFu;
            if (fits) continue;
            // (no room in the lists -- they are sized for 7 x groups x memberships: summed right here, in one piece)
        }
        float yb = 0.f, kb = 0.f;
        seg_accumulate(S, src, 0u, n, node, qx, qy, qz, cx, cy, cz, reach, cull2, ell, sf2, lane, yb, kb);
        if (node >= 0) yk[((size_t) (t - t0) * 7 + nb) * (size_t) (groups * 32) + j] = make_float2(yb, kb);
    }
}

// one warp per chunk of a long neighbour list: partial (ybar, kbar) of the unit's 32 leaves over kSegChunk segments
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_bgkl_yk_long(const NeighbourPlan *__restrict__ plan, const float4 *__restrict__ segs, const long long *__restrict__ keys,
               const unsigned char *__restrict__ pool, const float3 *__restrict__ lut, const DevParams *__restrict__ Pg,
               const ScanCounters *__restrict__ cnt, unsigned int t0, const unsigned int *__restrict__ long_cnt,
               const uint4 *__restrict__ long_units, const unsigned int *__restrict__ chunk_unit, unsigned int long_cap,
               unsigned int chunk_cap, float2 *partial) {
    __shared__ SegSmem sm[kWarpsPerCta];
    __shared__ DevParams Ps;
    load_params(Ps, Pg);
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SegSmem &S = sm[warp];
    const unsigned int n_chunks = min(long_cnt[1], chunk_cap);
    const unsigned int gw = blockIdx.x * kWarpsPerCta + warp, n_w = gridDim.x * kWarpsPerCta;
    const float ell = P.ell, sf2 = P.sf2, bs = P.block_size;
    const unsigned int groups = (unsigned int) (P.finest + 31) / 32, per_block = 7u * groups;
    const int pruned = P.pruned_state;
    const float reach = 0.5f * (bs - P.resolution) * 1.001f;
    const float cull2 = ell * ell * (1.0f + 1e-3f);
    for (unsigned int ci = gw; ci < n_chunks; ci += n_w) {
        const unsigned int slot_u = chunk_unit[ci];
        if (slot_u >= long_cap) continue;
        const uint4 lu = long_units[slot_u];
        const unsigned int u = lu.x, c = ci - lu.y;
        const unsigned int t = t0 + u / per_block, r = u % per_block;
        const int nb = (int) (r / groups), s = (int) (r % groups);
        const NeighbourPlan *pl = plan + t;
        const unsigned int n = pl->count[nb];
        const unsigned int slot = pl->slot;
        int node = -1;
        const int j = lane + 32 * s;
        if (j < P.finest) {
            if (pl->is_new) node = P.layer_off[P.depth - 1] + j;
            else {
                const unsigned char *rst = pool + (size_t) slot * (size_t) P.rec_bytes + P.st_off;
                int d = P.depth - 1, i = j, shift = 0;
                while (d > 0 && (rst[P.layer_off[d] + i] & 7) == pruned) { --d; i >>= 3; shift += 3; }
                if (((i << shift) == j) && ((rst[P.layer_off[d] + i] & 7) != pruned)) node = P.layer_off[d] + i;
            }
        }
        const long long key = keys[slot];
        const float cx = axis_center(key >> 40, bs), cy = axis_center((key >> 20) & 0xFFFFF, bs),
                    cz = axis_center(key & 0xFFFFF, bs);
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (node >= 0) {
            const float3 o = lut[node];
            qx = o.x + cx; qy = o.y + cy; qz = o.z + cz;           // Block::get_loc
        }
        float yb = 0.f, kb = 0.f;
        if (__any_sync(0xffffffffu, node >= 0))
            seg_accumulate(S, segs + 2 * (size_t) pl->start[nb], c * kSegChunk, min(n, (c + 1u) * kSegChunk), node, qx, qy,
                           qz, cx, cy, cz, reach, cull2, ell, sf2, lane, yb, kb);
        partial[(size_t) ci * 32 + lane] = make_float2(yb, kb);
    }
}

// (ybar, kbar) of a long unit = its chunks' partial sums added in chunk order; a lane per leaf, a warp per unit
__global__ void k_bgkl_yk_sum(const DevParams *__restrict__ Pg, const ScanCounters *__restrict__ cnt, unsigned int t0,
                              const unsigned int *__restrict__ long_cnt, const uint4 *__restrict__ long_units,
                              unsigned int long_cap, const float2 *__restrict__ partial, float2 *yk) {
    if (cnt->overflow) return;
    const int lane = threadIdx.x & 31;
    const unsigned int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_w = (gridDim.x * blockDim.x) >> 5;
    const unsigned int n_units = min(long_cnt[0], long_cap);
    const unsigned int groups = (unsigned int) (Pg->finest + 31) / 32, per_block = 7u * groups;
    for (unsigned int q = gw; q < n_units; q += n_w) {
        const uint4 lu = long_units[q];
        if (lu.z == 0u) continue;
        float yb = 0.f, kb = 0.f;
        for (unsigned int c = 0; c < lu.z; ++c) {
            const float2 p = partial[(size_t) (lu.y + c) * 32 + lane];
            yb += p.x; kb += p.y;
        }
        const unsigned int tl = lu.x / per_block, r = lu.x % per_block;
        const unsigned int nb = r / groups, s = r % groups;
        const unsigned int j = (unsigned int) lane + 32u * s;
        if (j < (unsigned int) Pg->finest) yk[((size_t) tl * 7 + nb) * (size_t) (groups * 32) + j] = make_float2(yb, kb);
    }
}

// one warp per test block; the record staged in (dynamic) shared memory when it fits (block_depth <= 4), updated in
// place in global memory otherwise; leaf after leaf (32 at a time, a lane each): Occupancy::update's accumulation with
// the (ybar, kbar) of every neighbour in ExtendedBlock order, guarded by kbar > 0.001f (bgkloctomap.cpp:231),
// classification once per touched leaf; then prune and write back
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_bgkl_apply(const NeighbourPlan *__restrict__ plan, unsigned char *__restrict__ pool, const DevParams *__restrict__ Pg,
             const ScanArgs *__restrict__ A, ScanCounters *cnt, unsigned int t0, unsigned int chunk,
             const float2 *__restrict__ yk, int staged) {
    extern __shared__ __align__(16) unsigned char bgkl_smem_raw[];
    __shared__ DevParams Ps;
    load_params(Ps, Pg);
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int T = cnt->n_test_blocks;
    if (t0 >= T) return;
    const unsigned int t1 = min(T, t0 + chunk);
    const unsigned int gw = blockIdx.x * kWarpsPerCta + warp, n_w = gridDim.x * kWarpsPerCta;
    const unsigned int shard_world = (unsigned int) A->shard_world, shard_rank = (unsigned int) A->shard_rank;
    const UpdateParams U{P.var_thresh, P.occupied_thresh, P.free_thresh};
    const int groups = (P.finest + 31) / 32;
    const int pruned = P.pruned_state;
    unsigned long long visits = 0, updates = 0, pairs = 0;

    for (unsigned int t = t0 + gw; t < t1; t += n_w) {
        if (t % shard_world != shard_rank) continue;
        const NeighbourPlan pl = plan[t];
        uint4 *grec = reinterpret_cast<uint4 *>(pool + (size_t) pl.slot * (size_t) P.rec_bytes);
        uint4 *wrec = staged ? reinterpret_cast<uint4 *>(bgkl_smem_raw + (size_t) warp * P.rec_bytes) : grec;
        __syncwarp();
        if (staged || pl.is_new) stage_record(wrec, grec, pl.is_new != 0, P, lane);
        __syncwarp();
        float2 *rab = reinterpret_cast<float2 *>(wrec);
        unsigned char *rst = reinterpret_cast<unsigned char *>(wrec) + P.st_off;
        unsigned int n_total = 0;
        for (int nb = 0; nb < 7; ++nb) n_total += pl.count[nb];
        bool any = false;
        for (int s = 0; s < groups; ++s) {
            const int j = lane + 32 * s;
            int node = -1;
            if (j < P.finest) {
                int d = P.depth - 1, i = j, shift = 0;
                while (d > 0 && (rst[P.layer_off[d] + i] & 7) == pruned) { --d; i >>= 3; shift += 3; }
                if (((i << shift) == j) && ((rst[P.layer_off[d] + i] & 7) != pruned)) node = P.layer_off[d] + i;
            }
            if (node < 0) continue;
            ++visits;
            pairs += n_total;
            if (n_total == 0) continue;
            float2 v = rab[node];
            bool touched = false;
            for (int nb = 0; nb < 7; ++nb) {
                if (pl.count[nb] == 0) continue;
                const float2 m = yk[((size_t) (t - t0) * 7 + nb) * (size_t) (groups * 32) + j];
                if (m.y > 0.001f) { v.x += m.x; v.y += m.y - m.x; touched = true; }
            }
            if (touched) {
                rab[node] = v;
                rst[node] = bgk_classify(v.x, v.y, U) | 0x80;
                ++updates;
                any = true;
            }
        }
        const bool dirty = __any_sync(0xffffffffu, any) || pl.is_new;
        __syncwarp();
        if (dirty) {
            prune_record(rab, rst, P, lane);
            if (staged) for (int w = lane; w < (P.rec_bytes >> 4); w += 32) grec[w] = wrec[w];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        visits += __shfl_xor_sync(0xffffffffu, visits, o);
        updates += __shfl_xor_sync(0xffffffffu, updates, o);
        pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    }
    if (lane == 0 && visits) {
        atomicAdd(&cnt->visits, visits);
        atomicAdd(&cnt->updates, updates);
        atomicAdd(&cnt->pairs, pairs);
    }
}

}  // namespace

// long-list buffers of the predict stage for the current capacities (ensure_workspace)
bool Map::ensure_bgkl_workspace() {
    const size_t groups = (size_t) (hp.finest + 31) / 32;
    const unsigned int chunk = bgkl_chunk(hp);
    const size_t long_cap = 7 * groups * (size_t) caps.members / kLongSegs + 64;
    const size_t chunk_cap = 7 * groups * (size_t) caps.members / kSegChunk + long_cap + 64;
    const size_t iters = (size_t) caps.tests / chunk + 2;
    bool moved = bgkl_long_units.reserve(long_cap * sizeof(uint4), stream);
    moved |= bgkl_chunk_unit.reserve(chunk_cap * 4, stream);
    moved |= bgkl_partial.reserve(chunk_cap * 32 * sizeof(float2), stream);
    moved |= bgkl_long_cnt.reserve(iters * 8, stream);
    return moved;
}

// per-block training lists (after the memberships were sorted by block: engine's binning stage)
void Map::enqueue_bgkl_lists(const unsigned int *sorted_keys, const unsigned int *sorted_vals) {
    const int m_tiles = ceil_div(caps.members, kTile);
    unsigned int *tile_sums = tiles.as<unsigned int>();
    k_segl_count<<<m_tiles, kTileThreads, 0, stream>>>(sorted_keys, sorted_vals, ray_of.as<int>(), d_cnt, caps.members,
                                                       tile_sums);
    k_segl_place<<<m_tiles, kTileThreads, 0, stream>>>(sorted_keys, sorted_vals, ray_of.as<int>(), d_cnt, caps.members,
                                                       tile_sums, (unsigned int) m_tiles, xy.as<float4>(),
                                                       rays.as<float4>(), cell_db.as<unsigned int>(), segs.as<float4>(),
                                                       seg_start.as<unsigned int>());
    launches += 2;
}

void Map::enqueue_predict_bgkl() {
    const int ctas = num_sms * 4;
    const unsigned int chunk = bgkl_chunk(hp);
    const size_t rec_smem = (size_t) hp.rec_bytes * kWarpsPerCta;
    const int staged = rec_smem <= 40 * 1024 ? 1 : 0;          // block_depth <= 4 (5.3 KB per record)
    record_event(ev_p0);
    // long neighbour lists: units, chunks and partial sums (sized for every membership in 7 x groups lists)
    const size_t groups = (size_t) (hp.finest + 31) / 32;
    const size_t long_cap = 7 * groups * (size_t) caps.members / kLongSegs + 64;
    const size_t chunk_cap = 7 * groups * (size_t) caps.members / kSegChunk + long_cap + 64;
    const size_t iters = (size_t) caps.tests / chunk + 2;
    if (bgkl_long_units.reserve(long_cap * sizeof(uint4), stream) | bgkl_chunk_unit.reserve(chunk_cap * 4, stream) |
        bgkl_partial.reserve(chunk_cap * 32 * sizeof(float2), stream) | bgkl_long_cnt.reserve(iters * 8, stream))
        throw StatusError{LA3DM_ERR_CUDA, "BGKL long-list buffers must be sized by ensure_workspace"};
    LA3DM_CUDA(cudaMemsetAsync(bgkl_long_cnt.p, 0, iters * 8, stream));
    unsigned int it = 0;
    for (unsigned int t0 = 0; t0 < caps.tests; t0 += chunk, ++it) {
        unsigned int *lc = bgkl_long_cnt.as<unsigned int>() + 2 * it;
        k_bgkl_yk<<<ctas, kWarpsPerCta * 32, 0, stream>>>(plan.as<NeighbourPlan>(), segs.as<float4>(),
                                                          keys.as<long long>(), pool.as<unsigned char>(), d_lut, d_params,
                                                          d_args, d_cnt, t0, chunk, gp_mv.as<float2>(), lc,
                                                          bgkl_long_units.as<uint4>(), bgkl_chunk_unit.as<unsigned int>(),
                                                          (unsigned int) long_cap, (unsigned int) chunk_cap);
        k_bgkl_yk_long<<<ctas, kWarpsPerCta * 32, 0, stream>>>(plan.as<NeighbourPlan>(), segs.as<float4>(),
                                                               keys.as<long long>(), pool.as<unsigned char>(), d_lut,
                                                               d_params, d_cnt, t0, lc, bgkl_long_units.as<uint4>(),
                                                               bgkl_chunk_unit.as<unsigned int>(), (unsigned int) long_cap,
                                                               (unsigned int) chunk_cap, bgkl_partial.as<float2>());
        k_bgkl_yk_sum<<<num_sms, 256, 0, stream>>>(d_params, d_cnt, t0, lc, bgkl_long_units.as<uint4>(),
                                                   (unsigned int) long_cap, bgkl_partial.as<float2>(), gp_mv.as<float2>());
        k_bgkl_apply<<<ctas, kWarpsPerCta * 32, staged ? rec_smem : 0, stream>>>(
            plan.as<NeighbourPlan>(), pool.as<unsigned char>(), d_params, d_args, d_cnt, t0, chunk, gp_mv.as<float2>(),
            staged);
        launches += 4;
    }
    record_event(ev_p1);
}

}  // namespace la3dm_b200
