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
    sincosf(t, &s, &c);
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

__global__ void k_scan_end(ScanCounters *c, const ScanArgs *__restrict__ A) {
    c->n_blocks = A->n_blocks + (c->overflow ? 0u : c->n_new_blocks);
}

}  // namespace

void Map::enqueue_predict() {
    if (hp.method != LA3DM_BGK) throw StatusError{LA3DM_ERR_UNSUPPORTED, "predict: method not implemented yet"};
    const int ctas = num_sms * 4;
    record_event(ev_p0);
    if (hp.depth <= 3)
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

void Map::enqueue_scan_end() {
    k_scan_end<<<1, 1, 0, stream>>>(d_cnt, d_args);
    ++launches;
}

}  // namespace la3dm_b200
