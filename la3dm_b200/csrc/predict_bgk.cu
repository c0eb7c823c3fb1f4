// la3dm_b200 -- the hot loop: fused  predict -> Occupancy::update -> OcTree::prune  for BGKOctoMap.
//
// Replaces the PREDICT and PRUNE loops of BGKOctoMap::insert_pointcloud (src/bgkoctomap/bgkoctomap.cpp:293-353):
//   BGKInference::predict / covSparse  (include/bgkoctomap/bgkinference.h:73-79, 113-126),
//   Occupancy::update / get_var / get_prob (src/bgkoctomap/bgkoctree_node.cpp:27-44, bgkoctree_node.h:60),
//   OcTree::is_leaf / prune (src/bgkoctomap/bgkoctree.cpp:72-82, 101-148), Block::get_loc (bgkblock.h:64-66).
//
// k_predict_bgk (block_depth <= 3, i.e. <= 64 finest voxels per block): one warp per test block, persistent grid.
//   * the training points of the 7 neighbour blocks (ExtendedBlock order) are streamed as ONE sequence in tiles of 32;
//     a point that cannot reach the hull of a slot's leaves is culled per tile (warp-uniform); a block none of whose
//     points survives only contributes its leaf count to the statistics (kept in a spare byte of the record);
//   * otherwise the block's record (alpha/beta + state bytes, one contiguous 16-byte aligned span) is staged in shared
//     memory with 16-byte accesses, updated and pruned there, and written back the same way;
//   * a lane owns the finest octree slots lane and lane + 32; a slot whose ancestors were pruned resolves to the
//     coarser leaf, handled by the lane that owns the leaf's first finest descendant;
//   * the compact-support test (d < 1) runs per (point, leaf) in registers; pairs inside the support are appended to a
//     per-warp queue (ballot + popc) and their queue position to the lane's private index list; the kernel function
//     (sqrt, sin, cos) is then evaluated DENSELY over the queue -- all 32 lanes busy instead of the few that are in
//     range -- and each lane adds its own pairs to (ybar, kbar) in training-array order, i.e. the same order of fp32
//     additions as the CPU oracle; the first pair of a new neighbour closes the previous one:
//     Occupancy::update's accumulation if kbar > 0, in ExtendedBlock order; classification once per leaf at the end.
// k_predict_bgk_deep (block_depth 4): plain formulation, 16 slots per lane, works on the record in global memory.
//
// Bound: issue slots / FP32 pipe (~24 flop per pair vs 17 B per voxel visit).
#include <stdlib.h>

#include "block_common.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kPtTile = 32;
constexpr int kQCap = 128;            // queue entries per warp; flushed when fewer than 64 are free

// covSparse element (bgkinference.h:115-116), d already scaled by 1/ell; caller guarantees d <= 1
__device__ __forceinline__ float sparse_kernel(float d, float sf2) {
    const float t = d * 2.0f * 3.1415926f;
    float s, c;
    sincosf_libm(t, s, c);
    float k = (((2.0f + c) * (1.0f - d) / 3.0f) + s / (2.0f * 3.1415926f)) * sf2;
    return k < 0.0f ? 0.0f : k;      // bgkinference.h:120-125
}

struct UpdateParams {
    float var_thresh, occupied_thresh, free_thresh;
};

// state of a node from (m_A, m_B): the tail of Occupancy::update (bgkoctree_node.cpp:36-43, get_var: .h:60)
__device__ __forceinline__ unsigned char bgk_classify(float a, float b, const UpdateParams &P) {
    const float var = (a * b) / ((a + b) * (a + b) * (a + b + 1.0f));
    if (var > P.var_thresh) return LA3DM_UNKNOWN;
    const float p = a / (a + b);
    return p > P.occupied_thresh ? LA3DM_OCCUPIED : (p < P.free_thresh ? LA3DM_FREE : LA3DM_UNKNOWN);
}

// Occupancy::update (bgkoctree_node.cpp:31-44); returns the new state
__device__ __forceinline__ unsigned char bgk_update(float &a, float &b, float ybar, float kbar, const UpdateParams &P) {
    a += ybar;
    b += kbar - ybar;
    return bgk_classify(a, b, P);
}

constexpr int kMine = 24;                  // points between flushes = private index-list entries per (lane, slot)
constexpr unsigned char kFirst = 0x80u;    // index-list flag: the lane's first pair of a neighbour

struct WarpSmem {
    uint4 rec[kRecMax / 16];               // the block record
    float4 pts[kPtTile];                   // current tile of training points (x/ell, y/ell, z/ell, label)
    float qd[kQCap];                       // shared queue: squared distance of an in-support pair in, kernel value out
    float qw[kQCap];                       //               label of the pair's training point
    unsigned char mine[2][kMine][32];      // per (slot, lane): queue positions of the lane's own pairs, in order
};

// leaves of a staged record: nodes that are not PRUNED and are either at the finest layer or have PRUNED children
// (is_leaf, bgkoctree.cpp:72-82); every lane gets the total
__device__ __forceinline__ int count_leaves(const unsigned char *rst, const DevParams &P, int lane) {
    int c = 0;
    for (int d = 0; d < P.depth; ++d) {
        const int off = P.layer_off[d], cnt = P.layer_off[d + 1] - off;
        for (int i = lane; i < cnt; i += 32) {
            if ((rst[off + i] & 7) == kStPRUNED) continue;
            if (d == P.depth - 1 || (rst[P.layer_off[d + 1] + 8 * i] & 7) == kStPRUNED) ++c;
        }
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

#ifndef LA3DM_PREDICT_MIN_CTAS
#define LA3DM_PREDICT_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kWarpsPerCta * 32, LA3DM_PREDICT_MIN_CTAS)
k_predict_bgk(const NeighbourPlan *__restrict__ plan, const float4 *__restrict__ pts,
              const long long *__restrict__ keys, unsigned char *__restrict__ pool, const float3 *__restrict__ lut,
              const DevParams *__restrict__ Pg, const ScanArgs *__restrict__ A, ScanCounters *cnt) {
    __shared__ WarpSmem sm[kWarpsPerCta];
    __shared__ DevParams Ps;
    if (threadIdx.x < sizeof(DevParams) / 4)
        reinterpret_cast<int *>(&Ps)[threadIdx.x] = reinterpret_cast<const int *>(Pg)[threadIdx.x];
    __syncthreads();
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt = (1u << lane) - 1u;
    WarpSmem &S = sm[warp];
    const unsigned int T = cnt->n_test_blocks;
    const unsigned int warps_total = gridDim.x * kWarpsPerCta;
    const int D = P.depth, finest = P.finest, nodes = P.nodes, st_off = P.st_off;
    const int rec_words = P.rec_bytes >> 4;
    const float ell = P.ell, sf2 = P.sf2, bs = P.block_size;
    const UpdateParams U{P.var_thresh, P.occupied_thresh, P.free_thresh};
    const int shard_world = A->shard_world, shard_rank = A->shard_rank;
    float2 *rab = reinterpret_cast<float2 *>(S.rec);
    unsigned char *rst = reinterpret_cast<unsigned char *>(S.rec) + st_off;

    // hull of the leaf centres of a slot, relative to the block centre, in units of ell (conservative): every leaf
    // centre lies within (block_size - resolution) / 2 of the block centre; with 64 finest voxels slot 0 / 1 hold the
    // lower / upper half in x (bit 5 of the finest index is the x bit of the depth-1 child, bgkblock.cpp:23-27)
    const float reach = 0.5f * (bs - P.resolution) * 1.001f / ell;
    const float hx0_hi = finest > 32 ? 1e-3f * reach : reach, hx1_lo = finest > 32 ? 0.0f : -reach;
    const float cull2 = 1.0f + 1e-4f;

    unsigned long long visits = 0, updates = 0, pairs = 0;

    // test block t belongs to rank t % world: this rank walks t = u * world + rank, u dealt over its warps
    for (unsigned int u = blockIdx.x * kWarpsPerCta + warp;; u += warps_total) {
        const unsigned int t = u * (unsigned int) shard_world + (unsigned int) shard_rank;
        if (t >= T) break;
        // ---- plan: lanes 0..6 hold start/count of one neighbour each
        const NeighbourPlan *pl = plan + t;
        const unsigned int slot = pl->slot, is_new = pl->is_new;
        const unsigned int my_start = lane < 7 ? pl->start[lane] : 0u, my_count = lane < 7 ? pl->count[lane] : 0u;
        unsigned int pre = my_count;                        // inclusive prefix over lanes 0..6
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += up;
        }
        const unsigned int tot = __shfl_sync(0xffffffffu, pre, 6);
        pre -= my_count;                                    // exclusive
        uint4 *grec = reinterpret_cast<uint4 *>(pool + (size_t) slot * (size_t) P.rec_bytes);
        // block centre from its key (hash_key_to_block, bgkblock.cpp:79-83)
        const long long key = keys[slot];
        const float cx = axis_center(key >> 40, bs), cy = axis_center((key >> 20) & 0xFFFFF, bs),
                    cz = axis_center(key & 0xFFFFF, bs);
        const float ccx = cx / ell, ccy = cy / ell, ccz = cz / ell;

        // ---- tile of the neighbours' points (ranges concatenated in ExtendedBlock order) + cull against the slot hulls
        unsigned int base = 0, bnd, c0, c1;
        float4 z;
        auto load_tile = [&]() {
            const unsigned int gi = base + lane;
            const bool valid = gi < tot;
            int nb = 0;
#pragma unroll
            for (int k = 1; k < 7; ++k) nb += (gi >= __shfl_sync(0xffffffffu, pre, k)) ? 1 : 0;
            const unsigned int nb_start = __shfl_sync(0xffffffffu, my_start, nb);
            const unsigned int nb_pre = __shfl_sync(0xffffffffu, pre, nb);
            z = make_float4(0.f, 0.f, 0.f, 0.f);
            bool keep0 = false, keep1 = false;
            if (valid) {
                z = pts[nb_start + (gi - nb_pre)];
                const float ry = fmaxf(fabsf(z.y - ccy) - reach, 0.f), rz = fmaxf(fabsf(z.z - ccz) - reach, 0.f);
                const float ryz = ry * ry + rz * rz;
                const float dxc = z.x - ccx;
                const float r0 = fmaxf(fmaxf(-reach - dxc, dxc - hx0_hi), 0.f);
                const float r1 = fmaxf(fmaxf(hx1_lo - dxc, dxc - reach), 0.f);
                keep0 = (r0 * r0 + ryz) < cull2;
                keep1 = (r1 * r1 + ryz) < cull2;
            }
            // first point of a neighbour's range: the previous neighbour's sums are complete
            bnd = __ballot_sync(0xffffffffu, valid && gi == nb_pre);
            c0 = __ballot_sync(0xffffffffu, keep0);
            c1 = __ballot_sync(0xffffffffu, keep1);
        };
        load_tile();

        if (tot <= (unsigned int) kPtTile && !(c0 | c1)) {
            // ---- no training point can reach a leaf of this block: only the statistics (and a fresh block's defaults)
            int n_leaves;
            if (is_new) {
                __syncwarp();
                stage_default_record(S.rec, P, lane);
                __syncwarp();
                for (int w = lane; w < rec_words; w += 32) grec[w] = S.rec[w];
                n_leaves = finest;
            } else {
                n_leaves = reinterpret_cast<const unsigned char *>(grec)[st_off + nodes];   // kept by the prune step
            }
            if (lane == 0) { visits += n_leaves; pairs += (unsigned long long) n_leaves * tot; }
            continue;
        }

        // ---- record -> shared memory
        __syncwarp();
        if (is_new) stage_default_record(S.rec, P, lane);
        else for (int w = lane; w < rec_words; w += 32) S.rec[w] = grec[w];
        __syncwarp();

        // ---- resolve this lane's leaves
        int node[2];
        float px[2], py[2], pz[2], a[2], b[2], yb[2], kb[2];
        unsigned char touched[2];
        bool pend[2];
        unsigned int mc[2];
        int owned = 0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int j = lane + 32 * s;
            node[s] = -1;
            touched[s] = 0;
            pend[s] = false;
            mc[s] = 0;
            a[s] = b[s] = px[s] = py[s] = pz[s] = 0.f;
            yb[s] = kb[s] = 0.f;
            if (j < finest) {
                // walk up while PRUNED: leaf (d, i) is owned by the lane of its first finest descendant
                int d = D - 1, i = j, shift = 0;
                while (d > 0 && (rst[P.layer_off[d] + i] & 7) == kStPRUNED) { --d; i >>= 3; shift += 3; }
                if (((i << shift) == j) && ((rst[P.layer_off[d] + i] & 7) != kStPRUNED)) {
                    const int n = P.layer_off[d] + i;
                    node[s] = n;
                    const float2 v = rab[n];
                    a[s] = v.x; b[s] = v.y;
                    const float3 off = lut[n];
                    // Block::get_loc: LUT offset + centre, then covSparse's  xs / ell
                    px[s] = (off.x + cx) / ell; py[s] = (off.y + cy) / ell; pz[s] = (off.z + cz) / ell;
                    ++owned;
                }
            }
        }
        visits += owned;
        pairs += (unsigned long long) owned * tot;
        const unsigned int have0 = __ballot_sync(0xffffffffu, node[0] >= 0) ? 0xffffffffu : 0u;
        const unsigned int have1 = __ballot_sync(0xffffffffu, node[1] >= 0) ? 0xffffffffu : 0u;

        // Drains the shared queue: kernel value of every queued pair (all lanes busy), then every lane adds its own
        // pairs to (ybar, kbar) in training order; a mark closes a neighbour: Occupancy::update's accumulation
        // (bgkoctree_node.cpp:31-35) if kbar > 0 (bgkoctomap.cpp:332).  The classification that follows it upstream only
        // survives for the last update of a scan, so it is done once at the end of the block.
        unsigned int nq = 0, since = 0;      // queue fill; points walked since the last flush (bounds every mc[])
        auto flush = [&]() {
            __syncwarp();
            for (unsigned int i = lane; i < nq; i += 32) S.qd[i] = sparse_kernel(sqrtf(S.qd[i]), sf2);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                for (unsigned int r = 0; r < mc[s]; ++r) {
                    const unsigned int e = S.mine[s][r][lane];
                    if (e & kFirst) {      // first pair of a new neighbour: the previous neighbour's sums are complete
                        if (kb[s] > 0.0f) { a[s] += yb[s]; b[s] += kb[s] - yb[s]; touched[s] = 1; }
                        yb[s] = kb[s] = 0.f;
                    }
                    const float k = S.qd[e & 0x7Fu];
                    yb[s] += k * S.qw[e & 0x7Fu];
                    kb[s] += k;
                }
                mc[s] = 0;
            }
            nq = 0;
            since = 0;
            __syncwarp();
        };

        // ---- stream the tiles
        while (true) {
            const unsigned int m0 = c0 & have0, m1 = c1 & have1;
            __syncwarp();
            S.pts[lane] = z;
            __syncwarp();
            unsigned int todo = m0 | m1 | bnd;
            while (todo) {
                const int q = __ffs(todo) - 1;
                todo &= todo - 1;
                if ((bnd >> q) & 1u) pend[0] = pend[1] = false;      // a new neighbour starts
                const unsigned int t0 = (m0 >> q) & 1u, t1 = (m1 >> q) & 1u;
                if (!(t0 | t1)) continue;
                const float4 zq = S.pts[q];
                if (t0) {
                    const float dx = zq.x - px[0], dy = zq.y - py[0], dz = zq.z - pz[0];
                    const float d2 = dx * dx + (dy * dy + dz * dz);    // Eigen rowwise().norm() of a 3-vector, squared
                    const bool in = node[0] >= 0 && d2 < 1.0f;        // k <= 0 for d >= 1 (clamped upstream)
                    const unsigned int mk = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const unsigned int e = nq + __popc(mk & lt);
                        S.qd[e] = d2;
                        S.qw[e] = zq.w;
                        S.mine[0][mc[0]++][lane] = (unsigned char) (pend[0] ? e : (e | kFirst));
                        pend[0] = true;
                    }
                    nq += __popc(mk);
                }
                if (t1) {
                    const float dx = zq.x - px[1], dy = zq.y - py[1], dz = zq.z - pz[1];
                    const float d2 = dx * dx + (dy * dy + dz * dz);
                    const bool in = node[1] >= 0 && d2 < 1.0f;
                    const unsigned int mk = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const unsigned int e = nq + __popc(mk & lt);
                        S.qd[e] = d2;
                        S.qw[e] = zq.w;
                        S.mine[1][mc[1]++][lane] = (unsigned char) (pend[1] ? e : (e | kFirst));
                        pend[1] = true;
                    }
                    nq += __popc(mk);
                }
                if (++since >= (unsigned int) kMine || nq > (unsigned int) (kQCap - 64)) flush();
            }
            base += kPtTile;
            if (base >= tot) break;
            load_tile();
        }
        flush();
#pragma unroll
        for (int s = 0; s < 2; ++s)      // the last neighbour
            if (kb[s] > 0.0f) { a[s] += yb[s]; b[s] += kb[s] - yb[s]; touched[s] = 1; }

        // ---- classify the touched leaves (the rest of Occupancy::update, bgkoctree_node.cpp:36-43) and write them
        // into the staged record
        bool any = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (node[s] >= 0 && touched[s]) {
                rab[node[s]] = make_float2(a[s], b[s]);
                rst[node[s]] = bgk_classify(a[s], b[s], U) | 0x80;   // classified = true
                ++updates;
                any = true;
            }
        }
        const bool dirty = __any_sync(0xffffffffu, any) || is_new;
        __syncwarp();
        if (dirty) {
            prune_record(rab, rst, P, lane);
            const int n_leaves = count_leaves(rst, P, lane);
            if (lane == 0) rst[nodes] = (unsigned char) n_leaves;     // spare byte behind the states (early-out above)
            __syncwarp();
            for (int w = lane; w < rec_words; w += 32) grec[w] = S.rec[w];
        }
    }

    // stats: one atomic per warp
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

// ---- k_predict_bgk_oct: block_depth == 3.  EIGHT lanes per test block, FOUR test blocks per warp. -----------------------
// Lane o of a block owns child o of every depth-1 octant: slot s = finest voxel 9 + 8s + o (so the pairs of a training
// point, which cluster in one or two octants, spread evenly over the 8 lanes).  If octant o is pruned, lane o's slot o
// holds the depth-1 leaf 1 + o instead (no lane has a voxel there); if the whole block is pruned, lane 0's slot 0 holds
// the root.
//   * the record never goes through a staging copy: a lane loads the (m_A, m_B) pairs and state bytes of its own nodes
//     (8 lanes x 8 bytes = one 64-byte line per load), keeps the floats in a per-lane column of shared memory (indexed
//     by slot at run time) and writes back only what changed;
//   * the per-axis centre coordinates of a lane's voxels are separable (init_key_loc_map, bgkblock.cpp:7-32: bits
//     4 / 2 / 1 of a child index pick x / y / z at every level), so one point costs 6 differences and 6 squares for 8
//     voxels;
//   * a tile = up to 8 points per block; points that cannot reach the hull of the block's leaf centres are dropped when
//     the tile is loaded (order preserved); the support test leaves one bit per (point, slot) in a 64-bit register;
//   * the pairs inside the support are numbered by one warp scan, their squared distances go to a shared queue, the
//     kernel function (sqrt, sin, cos) is evaluated DENSELY over the queue by all 32 lanes, and every lane adds its own
//     pairs to (ybar, kbar) in training-array order -- the order of fp32 additions of the CPU path; a pair from a new
//     neighbour closes the previous neighbour's sums (Occupancy::update if kbar > 0, ExtendedBlock order);
//     classification once per touched leaf; OcTree::prune by votes of the block's 8 lanes.
constexpr int kOctStride = 9;         // float4 per block in the point tile: 8 + 1 so the 4 broadcasts hit distinct banks
constexpr int kStripStride = 65;      // floats per lane in the pair strips: 8 points x 8 slots (+1: bank spread)

struct OctSmem {
    float a[8][32], b[8][32];         // [slot][lane]: m_A, m_B
    float yb[8][32], kb[8][32];       // [slot][lane]: sums over the current neighbour
    float strip[32 * kStripStride];   // per lane: squared distances of its in-support pairs in, kernel values out
    float4 pts[4 * kOctStride];       // surviving points of the current tile, per block
    unsigned int inc[32];             // inclusive scan of the lanes' pair counts
    unsigned char nb[32];             // neighbour (0..6) of each tile point
    unsigned char lnb[8][32];         // [slot][lane]: neighbour whose sums are being accumulated (0xFF: none yet)
};
constexpr size_t kOctSmemBytes = sizeof(OctSmem) * kWarpsPerCta + sizeof(DevParams);

template <int kMinCtas>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kMinCtas)
k_predict_bgk_oct(const NeighbourPlan *__restrict__ plan, const float4 *__restrict__ pts,
                  const long long *__restrict__ keys, unsigned char *__restrict__ pool, const float3 *__restrict__ lut,
                  const DevParams *__restrict__ Pg, const ScanArgs *__restrict__ A, ScanCounters *cnt,
                  const unsigned int *__restrict__ heavy_list) {
    extern __shared__ __align__(16) unsigned char oct_smem_raw[];
    OctSmem *sm = reinterpret_cast<OctSmem *>(oct_smem_raw);
    DevParams &Ps = *reinterpret_cast<DevParams *>(oct_smem_raw + sizeof(OctSmem) * kWarpsPerCta);
    if (threadIdx.x < sizeof(DevParams) / 4)
        reinterpret_cast<int *>(&Ps)[threadIdx.x] = reinterpret_cast<const int *>(Pg)[threadIdx.x];
    __syncthreads();
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, o = lane & 7, gbase = lane & 24;
    const unsigned int full = 0xffffffffu;
    OctSmem &S = sm[warp];
    float4 *tile = S.pts + g * kOctStride;
    unsigned char *tnb = S.nb + g * 8;
    const unsigned int T = cnt->n_test_blocks;
    const int nodes = P.nodes, st_off = P.st_off;
    const float ell = P.ell, sf2 = P.sf2, bs = P.block_size;
    const UpdateParams U{P.var_thresh, P.occupied_thresh, P.free_thresh};
    const unsigned int shard_world = (unsigned int) A->shard_world, shard_rank = (unsigned int) A->shard_rank;
    // every leaf centre lies within (block_size - resolution) / 2 of the block centre (conservative, in units of ell)
    const float reach = 0.5f * (bs - P.resolution) * 1.001f / ell;
    const float cull2 = 1.0f + 1e-4f;
    const float def_a = P.def_a, def_b = P.def_b;
    // centre offsets (block-relative) of child o of octants 0 / 4 / 2 / 1, of node 1 + o and of the root
    const float3 lf0 = lut[9 + o], lfx = lut[9 + 32 + o], lfy = lut[9 + 16 + o], lfz = lut[9 + 8 + o];
    const float3 l1 = lut[1 + o], l0 = lut[0];

    unsigned long long visits = 0, updates = 0, pairs = 0;

    // Work units of 4 test blocks are handed out through one atomic counter.  The heavy blocks (more than kHeavyTot
    // neighbourhood points, listed by k_plan) go first, four of similar weight to a warp; then all test blocks in
    // cell order, the heavy ones skipped.  Test block t belongs to rank t % world.
    const unsigned int heavy_tot = A->heavy_tot;
    const unsigned int n_heavy = heavy_list ? cnt->n_heavy : 0u;
    const unsigned int heavy_units = (n_heavy + 3u) >> 2;
    const unsigned int T_mine = T > shard_rank ? (T - shard_rank + shard_world - 1u) / shard_world : 0u;
    const unsigned int units = heavy_units + ((T_mine + 3u) >> 2);
    unsigned int w_next = 0;
    if (lane == 0) w_next = atomicAdd(&cnt->work_next, 1u);
    for (;;) {
        const unsigned int w = __shfl_sync(full, w_next, 0);
        if (w >= units) break;
        if (lane == 0) w_next = atomicAdd(&cnt->work_next, 1u);          // in flight while this unit is processed
        unsigned int t = 0;
        bool have;
        if (w < heavy_units) {
            const unsigned int idx = 4u * w + (unsigned int) g;
            have = idx < n_heavy;
            if (have) { t = heavy_list[idx]; have = t % shard_world == shard_rank; }
        } else {
            t = (4u * (w - heavy_units) + (unsigned int) g) * shard_world + shard_rank;
            have = t < T;
        }
        // ---- plan: lanes o = 0..6 of a group hold start / count of neighbour o
        unsigned int my_start = 0, my_count = 0, slot = 0, is_new = 0;
        if (have) {
            const NeighbourPlan *pl = plan + t;
            if (o < 7) { my_start = pl->start[o]; my_count = pl->count[o]; }
            slot = pl->slot;
            is_new = pl->is_new;
        }
        if (heavy_list && w >= heavy_units) {          // a heavy block met in cell order was done in the first phase
            unsigned int sum = my_count;
            sum += __shfl_xor_sync(full, sum, 1); sum += __shfl_xor_sync(full, sum, 2); sum += __shfl_xor_sync(full, sum, 4);
            if (sum > heavy_tot) { have = false; my_count = 0; is_new = 0; }
        }
        unsigned int pre = my_count;                                      // inclusive prefix inside the group
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const unsigned int up = __shfl_up_sync(full, pre, d, 8);
            if (o >= d) pre += up;
        }
        const unsigned int tot = __shfl_sync(full, pre, 7, 8);
        pre -= my_count;                                                  // exclusive
        unsigned int max_tot = tot;
        max_tot = max(max_tot, __shfl_xor_sync(full, max_tot, 8));
        max_tot = max(max_tot, __shfl_xor_sync(full, max_tot, 16));

        unsigned char *rec = pool + (size_t) slot * (size_t) P.rec_bytes;
        float2 *gab = reinterpret_cast<float2 *>(rec);
        unsigned char *gst = rec + st_off;

        // ---- this lane's part of the record: floats -> its shared-memory column, states -> a packed register pair
        unsigned int stlo = 0x02020202u, sthi = 0x02020202u;             // LA3DM_UNKNOWN x 8 (slot s: byte s)
        float2 ab1 = make_float2(def_a, def_b), ab0 = make_float2(def_a, def_b);
        unsigned int st1 = LA3DM_UNKNOWN, st0 = LA3DM_UNKNOWN, st_oct = LA3DM_UNKNOWN;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        __syncwarp();
        if (have && !is_new) {
            unsigned int sb[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float2 v = gab[9 + 8 * s + o];
                S.a[s][lane] = v.x; S.b[s][lane] = v.y;
                sb[s] = gst[9 + 8 * s + o];
            }
            stlo = sb[0] | (sb[1] << 8) | (sb[2] << 16) | (sb[3] << 24);
            sthi = sb[4] | (sb[5] << 8) | (sb[6] << 16) | (sb[7] << 24);
            ab1 = gab[1 + o]; st1 = gst[1 + o];
            ab0 = gab[0]; st0 = gst[0];
            st_oct = gst[9 + 8 * o];                                     // child 0 of octant o: PRUNED <=> octant o is pruned
        } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) { S.a[s][lane] = def_a; S.b[s][lane] = def_b; }
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) { S.yb[s][lane] = 0.f; S.kb[s][lane] = 0.f; S.lnb[s][lane] = 0xFFu; }
        if (have) {
            // block centre from its key (hash_key_to_block, bgkblock.cpp:79-83)
            const long long key = keys[slot];
            cx = axis_center(key >> 40, bs); cy = axis_center((key >> 20) & 0xFFFFF, bs); cz = axis_center(key & 0xFFFFF, bs);
        }
        // ---- leaves owned by this lane (is_leaf, bgkoctree.cpp:72-82).  A finest voxel is a leaf unless PRUNED; node
        // 1 + o is a leaf if it is not PRUNED and its children are; the root is a leaf if its children are PRUNED.
        const unsigned int st1_first = __shfl_sync(full, st1, 0, 8);
        const bool root_leaf = have && (st1_first & 7u) == kStPRUNED;
        const bool d1_leaf = have && !root_leaf && (st_oct & 7u) == kStPRUNED;
        const bool coarse = d1_leaf || (root_leaf && o == 0);            // lane's slot `cs` holds a coarse leaf
        const int cs = root_leaf ? 0 : o;
        // existing regular slots: byte s of the state words != PRUNED
        unsigned int vm = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const unsigned int sv = ((s < 4 ? stlo : sthi) >> (8 * (s & 3))) & 7u;
            if (have && sv != (unsigned int) kStPRUNED) vm |= 1u << s;
        }
        float ccx_ = 0.f, ccy_ = 0.f, ccz_ = 0.f;                        // centre of the coarse leaf, / ell
        if (coarse) {
            const float2 v = root_leaf ? ab0 : ab1;
            const unsigned int sv = root_leaf ? st0 : st1;
            S.a[cs][lane] = v.x; S.b[cs][lane] = v.y;
            if (cs < 4) stlo = (stlo & ~(0xFFu << (8 * cs))) | (sv << (8 * cs));
            else sthi = (sthi & ~(0xFFu << (8 * (cs - 4)))) | (sv << (8 * (cs - 4)));
            const float3 off = root_leaf ? l0 : l1;
            ccx_ = (off.x + cx) / ell; ccy_ = (off.y + cy) / ell; ccz_ = (off.z + cz) / ell;
            vm |= 1u << cs;
        }
        // Block::get_loc (bgkblock.h:64-66): LUT offset + centre, then covSparse's  xs / ell  (bgkinference.h:114)
        const float x0 = (lf0.x + cx) / ell, y0 = (lf0.y + cy) / ell, z0 = (lf0.z + cz) / ell;
        const float x1 = (lfx.x + cx) / ell, y1 = (lfy.y + cy) / ell, z1 = (lfz.z + cz) / ell;
        const float ccx = cx / ell, ccy = cy / ell, ccz = cz / ell;
        const unsigned int cbit = coarse ? (1u << cs) : 0u;
        const unsigned int regular = vm & ~cbit;
        const bool any_coarse = __ballot_sync(full, coarse) != 0u;
        float *strip = S.strip + lane * kStripStride;
        const int nleaf = __popc(vm);
        visits += (unsigned long long) nleaf;
        pairs += (unsigned long long) nleaf * tot;

        unsigned int touched = 0;                  // per slot: Occupancy::update ran

        // ---- stream the points of the 7 neighbours (ranges concatenated in ExtendedBlock order), 8 per block at a time
        // (the next tile's points are requested before the current tile is processed: the load latency hides behind the
        // support test instead of stalling the warp at the head of every tile)
        float4 z_next = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned int nbi_next = 0;
        auto fetch_tile = [&](unsigned int base) {
            const unsigned int gi = base + (unsigned int) o;
            unsigned int nbi = 0;
#pragma unroll
            for (int k = 1; k < 7; ++k) nbi += (gi >= __shfl_sync(full, pre, k, 8)) ? 1u : 0u;
            const unsigned int nb_start = __shfl_sync(full, my_start, (int) nbi, 8);
            const unsigned int nb_pre = __shfl_sync(full, pre, (int) nbi, 8);
            nbi_next = nbi;
            if (gi < tot) z_next = pts[nb_start + (gi - nb_pre)];
        };
        if (max_tot) fetch_tile(0);
        for (unsigned int base = 0; base < max_tot; base += 8) {
            const unsigned int gi = base + (unsigned int) o;
            const bool valid = gi < tot;
            const unsigned int nbi = nbi_next;
            float4 z = z_next;
            if (base + 8 < max_tot) fetch_tile(base + 8);
            bool keep = false;
            if (valid) {
                const float rx = fmaxf(fabsf(z.x - ccx) - reach, 0.f), ry = fmaxf(fabsf(z.y - ccy) - reach, 0.f),
                            rz = fmaxf(fabsf(z.z - ccz) - reach, 0.f);
                keep = (rx * rx + (ry * ry + rz * rz)) < cull2;
            }
            const unsigned int kept = __ballot_sync(full, keep);
            const unsigned int gk = (kept >> gbase) & 0xFFu;
            const int nsurv = __popc(gk);
            const int max_surv = max(max(__popc(kept & 0xFFu), __popc(kept & 0xFF00u)),
                                     max(__popc(kept & 0xFF0000u), __popc(kept & 0xFF000000u)));
            if (max_surv == 0) continue;
            __syncwarp();
            if (keep) {
                const int pos = __popc(gk & ((1u << o) - 1u));
                tile[pos] = z;
                tnb[pos] = (unsigned char) nbi;
            }
            __syncwarp();

            // ---- support test: bit 8 (q & 3) + s of mlo (q < 4) / mhi (q >= 4)  <=>  point q is within ell of this lane's
            // regular slot s; the squared distance of every such pair goes to the lane's strip (point-major order)
            unsigned int mlo = 0, mhi = 0, mco = 0;
            unsigned int n = 0;
            for (int q = 0; q < max_surv; ++q) {
                const float4 zq = tile[q];
                const float dx0 = zq.x - x0, dx1 = zq.x - x1, dy0 = zq.y - y0, dy1 = zq.y - y1, dz0 = zq.z - z0,
                            dz1 = zq.z - z1;
                const float xx0 = dx0 * dx0, xx1 = dx1 * dx1;
                const float yy0 = dy0 * dy0, yy1 = dy1 * dy1, zz0 = dz0 * dz0, zz1 = dz1 * dz1;
                const float s00 = yy0 + zz0, s01 = yy0 + zz1, s10 = yy1 + zz0, s11 = yy1 + zz1;   // [y bit][z bit]
                // d2 = dx*dx + (dy*dy + dz*dz): Eigen rowwise().norm() of a 3-vector, squared; k <= 0 for d >= 1
                const float d0 = xx0 + s00, d1 = xx0 + s01, d2 = xx0 + s10, d3 = xx0 + s11, d4 = xx1 + s00,
                            d5 = xx1 + s01, d6 = xx1 + s10, d7 = xx1 + s11;
                unsigned int in = (d0 < 1.0f ? 1u : 0u) | (d1 < 1.0f ? 2u : 0u) | (d2 < 1.0f ? 4u : 0u) |
                                  (d3 < 1.0f ? 8u : 0u) | (d4 < 1.0f ? 16u : 0u) | (d5 < 1.0f ? 32u : 0u) |
                                  (d6 < 1.0f ? 64u : 0u) | (d7 < 1.0f ? 128u : 0u);
                in = q < nsurv ? (in & regular) : 0u;
                if (in & 1u) strip[n++] = d0;
                if (in & 2u) strip[n++] = d1;
                if (in & 4u) strip[n++] = d2;
                if (in & 8u) strip[n++] = d3;
                if (in & 16u) strip[n++] = d4;
                if (in & 32u) strip[n++] = d5;
                if (in & 64u) strip[n++] = d6;
                if (in & 128u) strip[n++] = d7;
                const unsigned int sh = in << (8 * (q & 3));
                if (q < 4) mlo |= sh; else mhi |= sh;
            }
            if (any_coarse) {                       // warp-uniform: the coarse leaves' pairs follow the regular ones
                for (int q = 0; q < max_surv; ++q) {
                    const float4 zq = tile[q];
                    const float dx = zq.x - ccx_, dy = zq.y - ccy_, dz = zq.z - ccz_;
                    const float dc = dx * dx + (dy * dy + dz * dz);
                    if (coarse && q < nsurv && dc < 1.0f) { strip[n++] = dc; mco |= 1u << q; }
                }
            }
            unsigned int inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int up = __shfl_up_sync(full, inc, d);
                if (lane >= d) inc += up;
            }
            const unsigned int total = __shfl_sync(full, inc, 31);
            if (total == 0) continue;
            S.inc[lane] = inc;
            __syncwarp();
            // ---- kernel value of every pair of the tile, all lanes busy: pair i belongs to the first lane L with inc[L] > i
            for (unsigned int i = lane; i < total; i += 32) {
                int L = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (S.inc[L + step - 1] <= i) L += step;
                const unsigned int excl = L ? S.inc[L - 1] : 0u;
                float *pd = S.strip + L * kStripStride + (i - excl);
                *pd = sparse_kernel(sqrtf(*pd), sf2);
            }
            __syncwarp();
            // ---- every lane adds its own pairs in the order it wrote them; the first pair of a new neighbour closes the
            // previous neighbour: Occupancy::update's accumulation (bgkoctree_node.cpp:31-35) if kbar > 0
            // (bgkoctomap.cpp:332)
            n = 0;
            auto add_pair = [&](int q, int sl) {
                const unsigned int nbq = tnb[q];
                const float wq = tile[q].w;
                const float k = strip[n++];
                float *col = &S.a[0][0] + sl * 32 + lane;                // a, b, yb, kb of slot sl: 256 floats apart
                unsigned char *ln = &S.lnb[0][0] + sl * 32 + lane;
                float ybv = col[512], kbv = col[768];
                if (nbq != (unsigned int) *ln) {
                    if (kbv > 0.0f) {
                        col[0] += ybv;
                        col[256] += kbv - ybv;
                        touched |= 1u << sl;
                    }
                    ybv = 0.f; kbv = 0.f;
                    *ln = (unsigned char) nbq;
                }
                col[512] = ybv + k * wq;
                col[768] = kbv + k;
            };
            while (mlo) { const int bit = __ffs(mlo) - 1; mlo &= mlo - 1; add_pair(bit >> 3, bit & 7); }
            while (mhi) { const int bit = __ffs(mhi) - 1; mhi &= mhi - 1; add_pair(4 + (bit >> 3), bit & 7); }
            while (mco) { const int q = __ffs(mco) - 1; mco &= mco - 1; add_pair(q, cs); }
            __syncwarp();
        }
        // ---- the last neighbour, then the rest of Occupancy::update (bgkoctree_node.cpp:36-43) once per touched leaf
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float kbv = S.kb[s][lane];
            float av = S.a[s][lane], bv = S.b[s][lane];
            if (kbv > 0.0f) {
                const float ybv = S.yb[s][lane];
                av += ybv; bv += kbv - ybv;
                S.a[s][lane] = av; S.b[s][lane] = bv;
                touched |= 1u << s;
            }
            if ((touched >> s) & 1u) {
                const unsigned int ns = (unsigned int) bgk_classify(av, bv, U) | 0x80u;
                if (s < 4) stlo = (stlo & ~(0xFFu << (8 * s))) | (ns << (8 * s));
                else sthi = (sthi & ~(0xFFu << (8 * (s - 4)))) | (ns << (8 * (s - 4)));
                ++updates;
            }
        }
        const unsigned int dirty_lanes = __ballot_sync(full, touched != 0u);
        const bool dirty = have && (((dirty_lanes >> gbase) & 0xFFu) != 0u || is_new != 0u);
        __syncwarp();                               // lane 0's column is read by the other lanes below

        // ---- OcTree::prune (bgkoctree.cpp:101-148).  Layer 2 -> 1: octant s collapses if its 8 voxels (slot s of the
        // block's 8 lanes) share FREE or OCCUPIED; node 1 + s takes child 0's floats and state (`classified` is not
        // copied, bgkoctree_node.h:40-45), the voxels become PRUNED
        unsigned int pr2 = 0, pr2_occ = 0;          // octants pruned now / as OCCUPIED (same in the block's 8 lanes)
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const unsigned int sv = ((s < 4 ? stlo : sthi) >> (8 * (s & 3))) & 7u;
            const bool reg = (regular >> s) & 1u;
            const unsigned int fv = (__ballot_sync(full, reg && sv == LA3DM_FREE) >> gbase) & 0xFFu;
            const unsigned int ov = (__ballot_sync(full, reg && sv == LA3DM_OCCUPIED) >> gbase) & 0xFFu;
            if (fv == 0xFFu || ov == 0xFFu) {
                pr2 |= 1u << s;
                if (ov == 0xFFu) pr2_occ |= 1u << s;
                if (s < 4) stlo = (stlo & ~(7u << (8 * s))) | ((unsigned int) kStPRUNED << (8 * s));
                else sthi = (sthi & ~(7u << (8 * (s - 4)))) | ((unsigned int) kStPRUNED << (8 * (s - 4)));
            }
        }
        // current depth-1 node 1 + o of this lane
        float2 c1ab = ab1;
        unsigned int c1st = st1;
        const unsigned int my_byte = ((o < 4 ? stlo : sthi) >> (8 * (o & 3))) & 0xFFu;    // state byte of slot o
        if (d1_leaf) { c1ab = make_float2(S.a[o][lane], S.b[o][lane]); c1st = my_byte; }
        const bool pr2_mine = (pr2 >> o) & 1u;
        if (pr2_mine) {          // child 0 of octant o is slot o of the block's first lane
            c1ab = make_float2(S.a[o][gbase], S.b[o][gbase]);
            c1st = (st1 & 0x80u) | (((pr2_occ >> o) & 1u) ? (unsigned int) LA3DM_OCCUPIED : (unsigned int) LA3DM_FREE);
        }
        // layer 1 -> 0: the 8 depth-1 nodes vote
        const unsigned int f1 = (__ballot_sync(full, (c1st & 7u) == LA3DM_FREE) >> gbase) & 0xFFu;
        const unsigned int o1 = (__ballot_sync(full, (c1st & 7u) == LA3DM_OCCUPIED) >> gbase) & 0xFFu;
        const bool pr1 = have && (f1 == 0xFFu || o1 == 0xFFu);
        const float c1a_first = __shfl_sync(full, c1ab.x, 0, 8), c1b_first = __shfl_sync(full, c1ab.y, 0, 8);
        float2 c0ab = ab0;
        unsigned int c0st = st0;
        if (root_leaf && o == 0) { c0ab = make_float2(S.a[0][lane], S.b[0][lane]); c0st = stlo & 0xFFu; }
        if (pr1) {
            c0ab = make_float2(c1a_first, c1b_first);
            c0st = (st0 & 0x80u) | (c1st & 7u);
            c1st = (c1st & 0x80u) | (unsigned int) kStPRUNED;
        }
        // ---- write back what changed (a fresh block: everything)
        if (dirty) {
            const bool all = is_new != 0u;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (!((regular >> s) & 1u)) continue;
                const bool tch = (touched >> s) & 1u;
                if (all || tch) gab[9 + 8 * s + o] = make_float2(S.a[s][lane], S.b[s][lane]);
                if (all || tch || ((pr2 >> s) & 1u))
                    gst[9 + 8 * s + o] = (unsigned char) (((s < 4 ? stlo : sthi) >> (8 * (s & 3))) & 0xFFu);
            }
            const bool d1_touched = d1_leaf && ((touched >> o) & 1u);
            if (all || pr2_mine || d1_touched) gab[1 + o] = c1ab;
            if (all || pr2_mine || d1_touched || pr1) gst[1 + o] = (unsigned char) c1st;
            if (o == 0) {
                if (all || pr1 || (root_leaf && (touched & 1u))) { gab[0] = c0ab; gst[0] = (unsigned char) c0st; }
                // leaf count behind the states (read by k_predict_bgk's early-out)
                gst[nodes] = (unsigned char) ((pr1 || root_leaf) ? 1u : 8u + 7u * (unsigned int) __popc(regular & ~pr2));
            } else if (all) {
                for (int n = st_off + nodes + 2 * o - 1; n < P.rec_bytes && n <= st_off + nodes + 2 * o; ++n) rec[n] = 0;
            }
        }
    }

    // stats: one atomic per warp
    for (int d = 16; d > 0; d >>= 1) {
        visits += __shfl_xor_sync(full, visits, d);
        updates += __shfl_xor_sync(full, updates, d);
        pairs += __shfl_xor_sync(full, pairs, d);
    }
    if (lane == 0 && visits) {
        atomicAdd(&cnt->visits, visits);
        atomicAdd(&cnt->updates, updates);
        atomicAdd(&cnt->pairs, pairs);
    }
}

// ---- block_depth 4: 512 finest voxels per block, 16 slots per lane, record updated in global memory ----------------
constexpr int kDeepSlots = 16;

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_predict_bgk_deep(const NeighbourPlan *__restrict__ plan, const float4 *__restrict__ pts,
                   const long long *__restrict__ keys, unsigned char *__restrict__ pool,
                   const float3 *__restrict__ lut, const DevParams *__restrict__ Pg, const ScanArgs *__restrict__ A,
                   ScanCounters *cnt) {
    __shared__ float4 tile[kWarpsPerCta][kPtTile];
    __shared__ DevParams Ps;
    if (threadIdx.x < sizeof(DevParams) / 4)
        reinterpret_cast<int *>(&Ps)[threadIdx.x] = reinterpret_cast<const int *>(Pg)[threadIdx.x];
    __syncthreads();
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const UpdateParams U{P.var_thresh, P.occupied_thresh, P.free_thresh};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int T = cnt->n_test_blocks;
    const unsigned int warps_total = gridDim.x * kWarpsPerCta;
    const int D = P.depth, finest = P.finest;
    const float ell = P.ell, sf2 = P.sf2;
    const int shard_world = A->shard_world, shard_rank = A->shard_rank;
    unsigned long long visits = 0, updates = 0, pairs = 0;

    // test block t belongs to rank t % world: this rank walks t = u * world + rank, u dealt over its warps
    for (unsigned int u = blockIdx.x * kWarpsPerCta + warp;; u += warps_total) {
        const unsigned int t = u * (unsigned int) shard_world + (unsigned int) shard_rank;
        if (t >= T) break;
        const NeighbourPlan pl = plan[t];
        unsigned char *rec = pool + (size_t) pl.slot * (size_t) P.rec_bytes;
        float2 *bab = reinterpret_cast<float2 *>(rec);
        unsigned char *bst = rec + P.st_off;
        if (pl.is_new) {
            for (int n = lane; n < P.nodes; n += 32) { bab[n] = make_float2(P.def_a, P.def_b); bst[n] = LA3DM_UNKNOWN; }
            __syncwarp();
        }
        const long long key = keys[pl.slot];
        const float cx = axis_center(key >> 40, P.block_size), cy = axis_center((key >> 20) & 0xFFFFF, P.block_size),
                    cz = axis_center(key & 0xFFFFF, P.block_size);
        int node[kDeepSlots];
        float px[kDeepSlots], py[kDeepSlots], pz[kDeepSlots], a[kDeepSlots], b[kDeepSlots];
        unsigned char state[kDeepSlots], touched[kDeepSlots];
#pragma unroll
        for (int s = 0; s < kDeepSlots; ++s) {
            const int j = lane + 32 * s;
            node[s] = -1;
            touched[s] = 0;
            state[s] = LA3DM_UNKNOWN;
            a[s] = b[s] = px[s] = py[s] = pz[s] = 0.f;
            if (j < finest) {
                int d = D - 1, i = j, shift = 0;
                while (d > 0 && (bst[P.layer_off[d] + i] & 7) == kStPRUNED) { --d; i >>= 3; shift += 3; }
                const unsigned char sb = bst[P.layer_off[d] + i];
                if (((i << shift) == j) && ((sb & 7) != kStPRUNED)) {
                    const int n = P.layer_off[d] + i;
                    node[s] = n;
                    state[s] = sb;
                    const float2 v = bab[n];
                    a[s] = v.x; b[s] = v.y;
                    const float3 off = lut[n];
                    px[s] = (off.x + cx) / ell; py[s] = (off.y + cy) / ell; pz[s] = (off.z + cz) / ell;
                    ++visits;
                }
            }
        }
        for (int nb = 0; nb < 7; ++nb) {
            const unsigned int cntp = pl.count[nb];
            if (cntp == 0) continue;
            const float4 *src = pts + pl.start[nb];
            float yb[kDeepSlots], kb[kDeepSlots];
#pragma unroll
            for (int s = 0; s < kDeepSlots; ++s) { yb[s] = 0.f; kb[s] = 0.f; }
            for (unsigned int base = 0; base < cntp; base += kPtTile) {
                const unsigned int m = min((unsigned int) kPtTile, cntp - base);
                __syncwarp();
                if ((unsigned int) lane < m) tile[warp][lane] = src[base + lane];
                __syncwarp();
                for (unsigned int q = 0; q < m; ++q) {
                    const float4 z = tile[warp][q];
#pragma unroll
                    for (int s = 0; s < kDeepSlots; ++s) {
                        if (node[s] < 0) continue;
                        const float dx = z.x - px[s], dy = z.y - py[s], dz = z.z - pz[s];
                        const float d2 = dx * dx + (dy * dy + dz * dz);
                        if (d2 < 1.0f) {
                            const float k = sparse_kernel(sqrtf(d2), sf2);
                            yb[s] += k * z.w;
                            kb[s] += k;
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < kDeepSlots; ++s) {
                if (node[s] >= 0) {
                    pairs += cntp;
                    if (kb[s] > 0.0f) {
                        state[s] = bgk_update(a[s], b[s], yb[s], kb[s], U) | 0x80;
                        touched[s] = 1;
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < kDeepSlots; ++s) {
            if (node[s] >= 0 && touched[s]) {
                bab[node[s]] = make_float2(a[s], b[s]);
                bst[node[s]] = state[s];
                ++updates;
            }
        }
        __syncwarp();
        for (int d = D - 1; d > 0; --d) {
            const int off = P.layer_off[d], poff = P.layer_off[d - 1];
            const int groups = 1 << (3 * (d - 1));
            for (int g = lane; g < groups; g += 32) {
                const unsigned char s0 = bst[off + 8 * g] & 7;
                if (s0 == LA3DM_FREE || s0 == LA3DM_OCCUPIED) {
                    bool same = true;
#pragma unroll
                    for (int i = 1; i < 8; ++i) same = same && ((bst[off + 8 * g + i] & 7) == s0);
                    if (same) {
                        bab[poff + g] = bab[off + 8 * g];
                        bst[poff + g] = (bst[poff + g] & 0x80) | s0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) bst[off + 8 * g + i] = (bst[off + 8 * g + i] & 0x80) | kStPRUNED;
                    }
                }
            }
            __syncwarp();
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


// ---- k_predict_bgk_flat: block_depth <= 3, ONE warp per test block, the (leaf x point) pair space FLATTENED over the
// lanes.  On the headline workload only 6 % of the (leaf, training point) pairs of a test block are inside the kernel's
// support, a block has 33 leaves and 8.5 points that can reach one of them on average, and a quarter of the blocks has
// none: the cost of a block is its fixed overhead and the idle lanes of whatever is mapped to "one lane per leaf" or
// "one lane per point".  So:
//   * the points of the 7 neighbour ranges are read 32 at a time, culled against the hull of the block's leaf centres and
//     compacted into shared memory; a block without survivors only adds its leaf count to the statistics;
//   * the block's leaves (is_leaf, bgkoctree.cpp:72-82) are listed once, with their centres already divided by ell
//     (Block::get_loc + covSparse's  xs / ell, rounded like the reference; block_depth 3: 21 divisions per block through
//     the per-axis tables, the LUT being separable);
//   * pair i = leaf (i / S) x survivor (i % S) -- every lane has a pair whatever the shape of the block (1 leaf x 130
//     points, 64 leaves x 3 points); pairs with d < 1 are appended to a ring in leaf-major order;
//   * whenever 32 pairs wait, all lanes evaluate the kernel function, one warp-shuffle SEGMENTED scan keyed by the leaf
//     sums k y and k per leaf, and the last lane of every segment adds the two sums to the leaf's accumulators;
//   * Occupancy::update (bgkoctree_node.cpp:31-44) runs once per leaf with sum > 0; (m_A, m_B) are read and written in
//     global memory for those leaves only (a quarter of the visits), the state bytes as 19 words per block.
// Summation order.  BGKInference::predict returns per neighbour block ybar = Ks y, kbar = rowsum(Ks), and
// insert_pointcloud applies  if (kbar > 0) update(ybar, kbar)  neighbour after neighbour (bgkoctomap.cpp:314-335).  Every
// kernel value is >= 0 (bgkinference.h:120-125), so "some neighbour has kbar > 0" is "the sum over all neighbours is
// > 0", and a neighbour with kbar == 0 adds nothing: m_A += sum k y, m_B += sum k - sum k y over the whole
// ExtendedBlock differs from the reference only in the association of the fp32 additions (the reference itself sums in
// R-tree order).  The order used here is fixed by the input alone (ring position), so results are reproducible and
// identical on every replica; parity against the compiled reference is checked at 1e-4 on the probability.
constexpr int kFlatWarps = 8;
constexpr int kFlatPts = 64;          // survivors staged per chunk
constexpr int kFlatQ = 128;           // ring capacity (in-support pairs waiting for evaluation; <= 31 + 64 at a time)
#ifndef LA3DM_FLAT_MIN_CTAS
#define LA3DM_FLAT_MIN_CTAS 4
#endif

struct FlatSmem {
    float4 leaf[64];                  // centre / ell of the block's leaves; .w = node index (bits)
    float4 pt[kFlatPts];              // survivors of the hull cull: (x / ell, y / ell, z / ell, label)
    float2 acc[64];                   // per leaf: (sum k y, sum k)
    float2 q[kFlatQ];                 // ring: (squared distance, label)
    unsigned int st[20];              // state bytes of the record and the spare bytes behind them, as words
    float tab[3][8];                  // per axis: (LUT offset + block centre) / ell of the 4 + 2 + 1 coordinates
    unsigned char ql[kFlatQ];         // ring: leaf position of the pair
};
static_assert(sizeof(FlatSmem) % 16 == 0, "per-warp slices stay 16-byte aligned");

// covSparse element for a squared distance in [0, 1) (already in units of ell): bgkinference.h:115-126 with the
// reference's constants, order of operations and roundings (IEEE sqrt and divisions, libm's sin / cos)
__device__ __forceinline__ float sparse_kernel_d2(float d2, float sf2) {
    const float d = sqrtf(d2);
    const float t = d * 2.0f * 3.1415926f;
    float s, c;
    sincosf_libm(t, s, c);
    const float k = (((2.0f + c) * (1.0f - d) / 3.0f) + s / (2.0f * 3.1415926f)) * sf2;
    return k < 0.0f ? 0.0f : k;      // bgkinference.h:120-125
}

template <bool kD3>
__global__ void __launch_bounds__(kFlatWarps * 32, LA3DM_FLAT_MIN_CTAS)
k_predict_bgk_flat(const NeighbourPlan *__restrict__ plan, const float4 *__restrict__ pts,
                   const long long *__restrict__ keys, unsigned char *__restrict__ pool,
                   const float3 *__restrict__ lut, const DevParams *__restrict__ Pg, const ScanArgs *__restrict__ A,
                   ScanCounters *cnt, const unsigned int *__restrict__ heavy_list) {
    __shared__ __align__(16) FlatSmem sm[kFlatWarps];
    __shared__ DevParams Ps;
    if (threadIdx.x < sizeof(DevParams) / 4)
        reinterpret_cast<int *>(&Ps)[threadIdx.x] = reinterpret_cast<const int *>(Pg)[threadIdx.x];
    __syncthreads();
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int full = 0xffffffffu, lt = (1u << lane) - 1u;
    FlatSmem &S = sm[warp];
    unsigned char *sst = reinterpret_cast<unsigned char *>(S.st);
    const unsigned int T = cnt->n_test_blocks;
    const int D = kD3 ? 3 : P.depth;
    const int nodes = kD3 ? 73 : P.nodes, st_off = kD3 ? 584 : P.st_off, rec_bytes = kD3 ? 672 : P.rec_bytes;
    const int nst_words = (nodes + 4) >> 2;                  // state bytes + the leaf-count byte behind them
    const int l1 = kD3 ? 1 : P.layer_off[1], l2 = kD3 ? 9 : (D > 2 ? P.layer_off[2] : 0x7fffffff);
    const float ell = P.ell, sf2 = P.sf2, bs = P.block_size;
    const float inv_ell = 1.0f / ell;
    const float occ_t = P.occupied_thresh, free_t = P.free_thresh, var_t = P.var_thresh;
    const unsigned int shard_world = (unsigned int) A->shard_world, shard_rank = (unsigned int) A->shard_rank;
    // every leaf centre lies within (block_size - resolution) / 2 of the block centre (conservative, in units of ell)
    const float reach = 0.5f * (bs - P.resolution) * 1.001f / ell;
    const float cull2 = 1.0f + 1e-4f;

    unsigned int visits = 0, updates = 0;
    unsigned long long pairs = 0;
#ifdef LA3DM_FLAT_STATS
    unsigned int dbg_it = 0, dbg_chunks = 0, dbg_drains = 0, dbg_surv = 0, dbg_in = 0, dbg_blocks = 0, dbg_np = 0;
#endif

    // Work units come from one atomic counter: first the heavy blocks (more than heavy_tot neighbourhood points, listed
    // by k_plan), one per unit, then units of kUnit consecutive test blocks of this rank (t % world == rank), the heavy
    // ones skipped.
    constexpr unsigned int kUnit = 4;
    const unsigned int heavy_tot = A->heavy_tot;
    const unsigned int n_heavy = heavy_list ? cnt->n_heavy : 0u;
    const unsigned int T_mine = T > shard_rank ? (T - shard_rank + shard_world - 1u) / shard_world : 0u;
    const unsigned int units = n_heavy + (T_mine + kUnit - 1u) / kUnit;
    unsigned int w_next = 0;
    if (lane == 0) w_next = atomicAdd(&cnt->work_next, 1u);
    for (;;) {
        const unsigned int w = __shfl_sync(full, w_next, 0);
        if (w >= units) break;
        if (lane == 0) w_next = atomicAdd(&cnt->work_next, 1u);          // in flight while this unit is processed
        const bool heavy_unit = w < n_heavy;
        const unsigned int n_in_unit = heavy_unit ? 1u : kUnit;
#pragma unroll 1
        for (unsigned int j = 0; j < n_in_unit; ++j) {
            unsigned int t;
            if (heavy_unit) t = heavy_list[w];
            else {
                t = (kUnit * (w - n_heavy) + j) * shard_world + shard_rank;
                if (t >= T) break;
            }
            // ---- plan: lanes 0..6 hold start / count of one neighbour each
            const unsigned int plw = lane < 16 ? reinterpret_cast<const unsigned int *>(plan + t)[lane] : 0u;
            const unsigned int my_start = plw;
            unsigned int my_count = __shfl_down_sync(full, plw, 7);
            if (lane >= 7) my_count = 0u;
            const unsigned int slot = __shfl_sync(full, plw, 14), is_new = __shfl_sync(full, plw, 15);
            unsigned int pre = my_count;                                  // inclusive prefix over lanes 0..6
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const unsigned int up = __shfl_up_sync(full, pre, o);
                if (lane >= o) pre += up;
            }
            const unsigned int tot = __shfl_sync(full, pre, 6);
            if (!heavy_unit && heavy_list && tot > heavy_tot) continue;   // done in the first phase
            pre -= my_count;                                              // exclusive
            const unsigned int delta = my_start - pre;                    // point gi of neighbour k sits at gi + delta_k
#ifdef LA3DM_FLAT_STATS
            ++dbg_blocks;
#endif
            unsigned char *rec = pool + (size_t) slot * (size_t) rec_bytes;
            float2 *gab = reinterpret_cast<float2 *>(rec);
            unsigned int *gst = reinterpret_cast<unsigned int *>(rec + st_off);
            // ---- state bytes (a fresh Block: the default node everywhere, bgkoctree_node.h:34; its record is written now)
            unsigned int stw = 0;
            if (is_new) {
                for (int n = lane; n < nodes; n += 32) gab[n] = make_float2(P.def_a, P.def_b);
                if (lane < nst_words) {
                    stw = 0x02020202u;                                    // LA3DM_UNKNOWN x 4
                    const int b0 = 4 * lane;                              // bytes b0 .. b0 + 3 of the state area
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (b0 + b >= nodes)
                            stw = (stw & ~(0xFFu << (8 * b))) | ((b0 + b == nodes ? (unsigned int) (P.finest & 0xFF) : 0u) << (8 * b));
                    gst[lane] = stw;
                }
            } else if (lane < nst_words) stw = gst[lane];
            const long long key = keys[slot];
            // ---- 32 points of the neighbourhood (ranges concatenated in ExtendedBlock order)
            auto fetch = [&](unsigned int base, float4 &z) -> bool {
                const unsigned int gi = base + (unsigned int) lane;
                unsigned int nbi = 0;
#pragma unroll
                for (int k = 1; k < 7; ++k) nbi += (gi >= __shfl_sync(full, pre, k)) ? 1u : 0u;
                const unsigned int d = __shfl_sync(full, delta, (int) nbi);
                if (gi < tot) { z = pts[gi + d]; return true; }
                return false;
            };
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            bool valid = fetch(0, z);
            // block centre from its key (hash_key_to_block, bgkblock.cpp:79-83); the hull test may round differently
            const float cx = axis_center(key >> 40, bs), cy = axis_center((key >> 20) & 0xFFFFF, bs),
                        cz = axis_center(key & 0xFFFFF, bs);
            const float ccx = cx * inv_ell, ccy = cy * inv_ell, ccz = cz * inv_ell;
            __syncwarp();
            if (lane < nst_words) S.st[lane] = stw;
            int Lf = -1;                                                  // leaves listed (-1: not yet)
            unsigned int qh = 0, qt = 0;                                  // ring head / tail (monotonic; position = & 63)

            // evaluates `c` (<= 32) waiting pairs: kernel value by all lanes, segmented sums per leaf, accumulate
            auto drain = [&](unsigned int c) {
                __syncwarp();
#ifdef LA3DM_FLAT_STATS
                ++dbg_drains;
#endif
                const unsigned int pos = (qh + (unsigned int) lane) & (kFlatQ - 1);
                const bool e = (unsigned int) lane < c;
                float vy = 0.f, vk = 0.f;
                unsigned int lp = 0xFFu;
                if (e) {
                    const float2 dq = S.q[pos];
                    lp = S.ql[pos];
                    vk = sparse_kernel_d2(dq.x, sf2);
                    vy = vk * dq.y;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float ty = __shfl_up_sync(full, vy, o), tk = __shfl_up_sync(full, vk, o);
                    const unsigned int tl = __shfl_up_sync(full, lp, o);
                    if (lane >= o && tl == lp) { vy += ty; vk += tk; }
                }
                const unsigned int nl = __shfl_down_sync(full, lp, 1);
                if (e && (lane == 31 || nl != lp)) {
                    float2 a = S.acc[lp];
                    a.x += vy; a.y += vk;
                    S.acc[lp] = a;
                }
                qh += c;
                __syncwarp();
            };

            unsigned int ns = 0;                                          // survivors staged in S.pt
#pragma unroll 1
            for (unsigned int base = 0; base < tot; base += 32) {
                if (base) valid = fetch(base, z);
                bool keep = false;
                if (valid) {
                    const float rx = fmaxf(fabsf(z.x - ccx) - reach, 0.f), ry = fmaxf(fabsf(z.y - ccy) - reach, 0.f),
                                rz = fmaxf(fabsf(z.z - ccz) - reach, 0.f);
                    keep = (rx * rx + (ry * ry + rz * rz)) < cull2;
                }
                const unsigned int kept = __ballot_sync(full, keep);
                if (keep) S.pt[ns + __popc(kept & lt)] = z;
                ns += __popc(kept);
                const bool last = base + 32 >= tot;
                if (ns == 0 || (!last && ns <= (unsigned int) (kFlatPts - 32))) continue;
                // ---- list the block's leaves once: centre / ell and node index; zero their accumulators
                if (Lf < 0) {
                    if (kD3) {
                        if (lane < 21) {
                            const int a = lane / 7, jj = lane - 7 * a;
                            const float c = a == 0 ? cx : (a == 1 ? cy : cz);
                            S.tab[a][jj] = (P.ax_off[a][jj] + c) / ell;
                        }
                    }
                    __syncwarp();
                    Lf = 0;
                    if (kD3) {
                        // lane: finest voxels 9 + lane and 9 + 32 + lane, then (lanes 0..8) the nine coarse nodes
#pragma unroll
                        for (int h = 0; h < 3; ++h) {
                            const int n = h < 2 ? 9 + 32 * h + lane : lane;
                            bool leaf = false;
                            if (h < 2) leaf = (sst[n] & 7) != kStPRUNED;
                            else if (lane < 9) leaf = (sst[n] & 7) != kStPRUNED && (sst[lane == 0 ? 1 : 1 + 8 * lane] & 7) == kStPRUNED;
                            const unsigned int m = __ballot_sync(full, leaf);
                            if (leaf) {
                                const int pos = Lf + __popc(m & lt);
                                int xj, yj, zj;
                                if (h < 2) {
                                    xj = 2 * h + ((lane >> 2) & 1); yj = ((lane >> 3) & 2) | ((lane >> 1) & 1);
                                    zj = ((lane >> 2) & 2) | (lane & 1);
                                } else if (lane > 0) { const int i = lane - 1; xj = 4 + ((i >> 2) & 1); yj = 4 + ((i >> 1) & 1); zj = 4 + (i & 1); }
                                else { xj = yj = zj = 6; }
                                S.leaf[pos] = make_float4(S.tab[0][xj], S.tab[1][yj], S.tab[2][zj], __int_as_float(n));
                                S.acc[pos] = make_float2(0.f, 0.f);
                            }
                            Lf += __popc(m);
                        }
                    } else {
                        for (int n0 = 0; n0 < nodes; n0 += 32) {
                            const int n = n0 + lane;
                            bool leaf = false;
                            if (n < nodes) {
                                const int d = n >= l2 ? 2 : (n >= l1 ? 1 : 0);
                                const int i = n - P.layer_off[d];
                                leaf = (sst[n] & 7) != kStPRUNED &&
                                       (d == D - 1 || (sst[P.layer_off[d + 1] + 8 * i] & 7) == kStPRUNED);
                            }
                            const unsigned int m = __ballot_sync(full, leaf);
                            if (leaf) {
                                const int pos = Lf + __popc(m & lt);
                                // Block::get_loc: LUT offset + centre, then covSparse's  xs / ell
                                const float3 off = lut[n];
                                S.leaf[pos] = make_float4((off.x + cx) / ell, (off.y + cy) / ell, (off.z + cz) / ell,
                                                          __int_as_float(n));
                                S.acc[pos] = make_float2(0.f, 0.f);
                            }
                            Lf += __popc(m);
                        }
                    }
                }
                __syncwarp();
                // ---- the chunk's pairs, leaf-major: pair i = leaf i / ns, survivor i % ns; each lane walks i = lane,
                // lane + 32, ... keeping (leaf, survivor) incrementally
                const unsigned int np = (unsigned int) Lf * ns;
#ifdef LA3DM_FLAT_STATS
                ++dbg_chunks; dbg_surv += ns; dbg_np += np;
#endif
                // (two pairs per lane and trip, i and i + 32: independent distance computations behind one loop overhead)
                const unsigned int q32 = 32u / ns, r32 = 32u - q32 * ns;
                const unsigned int w64 = 2u * r32 >= ns ? 1u : 0u, q64 = 2u * q32 + w64, r64 = 2u * r32 - w64 * ns;
                unsigned int lpA = (unsigned int) lane / ns, piA = (unsigned int) lane - lpA * ns;
                unsigned int lpB = lpA + q32, piB = piA + r32;
                if (piB >= ns) { piB -= ns; ++lpB; }
#pragma unroll 1
                for (unsigned int i0 = (unsigned int) lane; i0 < np + (unsigned int) lane; i0 += 64) {
                    float dA = 2.f, dB = 2.f, yA = 0.f, yB = 0.f;
                    if (i0 < np) {
                        const float4 L = S.leaf[lpA], pq = S.pt[piA];
                        const float dx = pq.x - L.x, dy = pq.y - L.y, dz = pq.z - L.z;
                        dA = dx * dx + (dy * dy + dz * dz);      // Eigen rowwise().norm() of a 3-vector, squared
                        yA = pq.w;
                    }
                    if (i0 + 32u < np) {
                        const float4 L = S.leaf[lpB], pq = S.pt[piB];
                        const float dx = pq.x - L.x, dy = pq.y - L.y, dz = pq.z - L.z;
                        dB = dx * dx + (dy * dy + dz * dz);
                        yB = pq.w;
                    }
                    const bool inA = dA < 1.0f, inB = dB < 1.0f;     // k <= 0 for d >= 1 (clamped upstream)
                    const unsigned int mA = __ballot_sync(full, inA), mB = __ballot_sync(full, inB);
#ifdef LA3DM_FLAT_STATS
                    ++dbg_it; dbg_in += __popc(mA) + __popc(mB);
#endif
                    if ((mA | mB) != 0u) {
                        const unsigned int cA = __popc(mA);
                        if (inA) {
                            const unsigned int pos = (qt + __popc(mA & lt)) & (kFlatQ - 1);
                            S.q[pos] = make_float2(dA, yA);
                            S.ql[pos] = (unsigned char) lpA;
                        }
                        if (inB) {
                            const unsigned int pos = (qt + cA + __popc(mB & lt)) & (kFlatQ - 1);
                            S.q[pos] = make_float2(dB, yB);
                            S.ql[pos] = (unsigned char) lpB;
                        }
                        qt += cA + __popc(mB);
                        while (qt - qh >= 32u) drain(32u);
                    }
                    lpA += q64; piA += r64;
                    if (piA >= ns) { piA -= ns; ++lpA; }
                    lpB += q64; piB += r64;
                    if (piB >= ns) { piB -= ns; ++lpB; }
                }
                if (qt != qh) drain(qt - qh);
                ns = 0;
                __syncwarp();
            }
            // ---- statistics; a block no training point can reach is done (its leaf count sits behind the states)
            if (Lf < 0) {
                const unsigned int wl = __shfl_sync(full, stw, nodes >> 2);
                const unsigned int n_leaves = (wl >> (8 * (nodes & 3))) & 0xFFu;
                if (lane == 0) { visits += n_leaves; pairs += (unsigned long long) n_leaves * tot; }
                continue;
            }
            if (lane == 0) { visits += (unsigned int) Lf; pairs += (unsigned long long) Lf * tot; }
            // ---- Occupancy::update (bgkoctree_node.cpp:31-44) for the leaves with kbar > 0 (bgkoctomap.cpp:332)
            bool changed = false, touched_any = false;
#pragma unroll 1
            for (int lp0 = 0; lp0 < Lf; lp0 += 32) {
                const int lq = lp0 + lane;
                if (lq < Lf) {
                    const float2 s = S.acc[lq];
                    if (s.y > 0.0f) {
                        const int n = __float_as_int(S.leaf[lq].w);
                        float2 ab = gab[n];
                        ab.x += s.x;
                        ab.y += s.y - s.x;
                        gab[n] = ab;
                        // get_var (bgkoctree_node.h:60) is below 1/4 for any (m_A, m_B) > 0: only evaluated if it can matter
                        unsigned int ns_ = LA3DM_UNKNOWN;
                        bool known = true;
                        if (var_t < 0.25f) known = !((ab.x * ab.y) / ((ab.x + ab.y) * (ab.x + ab.y) * (ab.x + ab.y + 1.0f)) > var_t);
                        if (known) {
                            const float p = ab.x / (ab.x + ab.y);
                            ns_ = p > occ_t ? LA3DM_OCCUPIED : (p < free_t ? LA3DM_FREE : LA3DM_UNKNOWN);
                        }
                        const unsigned int old = sst[n];
                        sst[n] = (unsigned char) (ns_ | 0x80u);   // classified = true
                        changed = changed || ((old & 7u) != ns_);
                        touched_any = true;
                        ++updates;
                    }
                }
            }
            if (!__any_sync(full, touched_any)) continue;
            __syncwarp();
            // ---- OcTree::prune (bgkoctree.cpp:101-148): only a state that changed can complete a group of 8 equal siblings
            if (__any_sync(full, changed)) {
                int n_pruned_groups = 0;
                bool root_pruned = false;
                for (int d = D - 1; d > 0; --d) {
                    const int off = P.layer_off[d], poff = P.layer_off[d - 1];
                    const int groups = 1 << (3 * (d - 1));
                    bool did = false;
                    for (int g = lane; g < groups; g += 32) {
                        const unsigned char s0 = sst[off + 8 * g] & 7;
                        if (s0 == LA3DM_FREE || s0 == LA3DM_OCCUPIED) {
                            bool same = true;
#pragma unroll
                            for (int i = 1; i < 8; ++i) same = same && ((sst[off + 8 * g + i] & 7) == s0);
                            if (same) {
                                gab[poff + g] = __ldcg(&gab[off + 8 * g]);    // parent := child 0 (classified is not copied)
                                sst[poff + g] = (sst[poff + g] & 0x80) | s0;
#pragma unroll
                                for (int i = 0; i < 8; ++i) sst[off + 8 * g + i] = (sst[off + 8 * g + i] & 0x80) | kStPRUNED;
                                did = true;
                            }
                        }
                    }
                    n_pruned_groups += __popc(__ballot_sync(full, did));      // (groups <= 32 per layer for block_depth <= 3)
                    if (d == 1) root_pruned = __any_sync(full, did);
                    __syncwarp();
                }
                // every collapsed group turns 8 leaves into 1
                if (n_pruned_groups && lane == 0) sst[nodes] = (unsigned char) (Lf - 7 * n_pruned_groups);
                (void) root_pruned;
                __syncwarp();
            }
            if (lane < nst_words) gst[lane] = S.st[lane];
        }
    }

    // stats: one atomic per warp
    unsigned long long v64 = visits, u64 = updates;
    for (int d = 16; d > 0; d >>= 1) {
        v64 += __shfl_xor_sync(full, v64, d);
        u64 += __shfl_xor_sync(full, u64, d);
        pairs += __shfl_xor_sync(full, pairs, d);
    }
    if (lane == 0 && v64) {
        atomicAdd(&cnt->visits, v64);
        atomicAdd(&cnt->updates, u64);
        atomicAdd(&cnt->pairs, pairs);
    }
#ifdef LA3DM_FLAT_STATS
    if (lane == 0) {
        atomicAdd(&cnt->n_long_runs[0], dbg_it); atomicAdd(&cnt->n_long_runs[1], dbg_chunks);
        atomicAdd(&cnt->n_mid_runs[0], dbg_drains); atomicAdd(&cnt->n_mid_runs[1], dbg_surv);
        atomicAdd(&cnt->reserved_, dbg_in); atomicAdd(&cnt->pad2_, dbg_blocks); atomicAdd(&cnt->vg_cells_needed, dbg_np);
    }
#endif
}


}  // namespace

void Map::enqueue_predict() {
    if (hp.method != LA3DM_BGK) throw StatusError{LA3DM_ERR_UNSUPPORTED, "predict: method not implemented yet"};
    const int ctas = num_sms * 4;
    static const bool force_v1 = getenv("LA3DM_PREDICT_V1") != nullptr;     // debugging: the one-warp-per-block kernel
    record_event(ev_p0);
    static const bool force_oct = getenv("LA3DM_PREDICT_OCT") != nullptr;   // A/B: round 1's eight-lanes-per-block kernel
    if (hp.depth <= 3 && !force_v1 && !force_oct) {
        auto kern = hp.depth == 3 ? k_predict_bgk_flat<true> : k_predict_bgk_flat<false>;
        int occ = 0;
        LA3DM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFlatWarps * 32, 0));
        if (occ < 1) occ = 1;
        kern<<<num_sms * occ, kFlatWarps * 32, 0, stream>>>(
            plan.as<NeighbourPlan>(), pts_sorted.as<float4>(), keys.as<long long>(), pool.as<unsigned char>(), d_lut,
            d_params, d_args, d_cnt, getenv("LA3DM_OCT_NO_HEAVY") ? nullptr : heavy_list.as<unsigned int>());
        LA3DM_CUDA(cudaGetLastError());
    }
    else if (hp.depth == 3 && !force_v1) {
        static const int oct_ctas = getenv("LA3DM_OCT_CTAS") ? atoi(getenv("LA3DM_OCT_CTAS")) : 2;
        const int n = oct_ctas == 1 ? 1 : 2;
        auto kern = n == 1 ? k_predict_bgk_oct<1> : k_predict_bgk_oct<2>;
        static bool attr_set[3] = {false, false, false};
        if (!attr_set[n]) {
            LA3DM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kOctSmemBytes));
            attr_set[n] = true;
        }
        kern<<<num_sms * n, kWarpsPerCta * 32, kOctSmemBytes, stream>>>(
            plan.as<NeighbourPlan>(), pts_sorted.as<float4>(), keys.as<long long>(), pool.as<unsigned char>(), d_lut,
            d_params, d_args, d_cnt, getenv("LA3DM_OCT_NO_HEAVY") ? nullptr : heavy_list.as<unsigned int>());
    }
    else if (hp.depth <= 3)
        k_predict_bgk<<<ctas, kWarpsPerCta * 32, 0, stream>>>(plan.as<NeighbourPlan>(), pts_sorted.as<float4>(),
                                                              keys.as<long long>(), pool.as<unsigned char>(), d_lut,
                                                              d_params, d_args, d_cnt);
    else if (hp.depth == 4)
        k_predict_bgk_deep<<<ctas, kWarpsPerCta * 32, 0, stream>>>(plan.as<NeighbourPlan>(), pts_sorted.as<float4>(),
                                                                   keys.as<long long>(), pool.as<unsigned char>(),
                                                                   d_lut, d_params, d_args, d_cnt);
    else throw StatusError{LA3DM_ERR_UNSUPPORTED, "block_depth > 4 not supported by the BGK kernel yet"};
    record_event(ev_p1);
    ++launches;
}


}  // namespace la3dm_b200
