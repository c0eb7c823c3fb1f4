// la3dm_b200 -- the hot loop: fused  predict -> Occupancy::update -> OcTree::prune  for BGKOctoMap.
//
// Replaces the PREDICT and PRUNE loops of BGKOctoMap::insert_pointcloud (src/bgkoctomap/bgkoctomap.cpp:293-353):
//   BGKInference::predict / covSparse  (include/bgkoctomap/bgkinference.h:73-79, 113-126),
//   Occupancy::update / get_var / get_prob (src/bgkoctomap/bgkoctree_node.cpp:27-44, bgkoctree_node.h:60),
//   OcTree::is_leaf / prune (src/bgkoctomap/bgkoctree.cpp:72-82, 101-148), Block::get_loc (bgkblock.h:64-66).
//
// Mapping: one warp per test block (persistent grid, warps stride over the test-block list).  A lane owns the finest
// octree slots lane, lane+32, ...; a slot whose ancestors were pruned resolves to the coarser leaf, handled by the
// lane that owns the leaf's first descendant.  For each of the 7 neighbour blocks in ExtendedBlock order the warp
// stages the neighbour's training points (float4: x/ell, y/ell, z/ell, label) through shared memory in 32-point tiles
// and every lane accumulates (ybar, kbar) for its leaves sequentially in training-array order -- the same order of
// fp32 additions as the CPU oracle -- then applies one Occupancy::update per neighbour with kbar > 0.
// Results go back as coalesced float2 (alpha, beta) + state bytes; pruning runs in the same warp afterwards.
//
// Bound: FP32/SFU pipe (~24 flop per pair vs 17 B per voxel visit).
#include "engine.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kTile = 32;

struct LeafRef {
    int node;     // index into the block's node array (layer_off[d] + index); -1: slot not owned by this lane
};

// covSparse element (bgkinference.h:115-116), d already scaled by 1/ell; caller guarantees d < 1
__device__ __forceinline__ float sparse_kernel(float d, float sf2) {
    const float t = d * 2.0f * 3.1415926f;
    float s, c;
    sincosf(t, &s, &c);
    float k = (((2.0f + c) * (1.0f - d) / 3.0f) + s / (2.0f * 3.1415926f)) * sf2;
    return k < 0.0f ? 0.0f : k;      // bgkinference.h:120-125
}

// Occupancy::update (bgkoctree_node.cpp:31-44); returns the new state
__device__ __forceinline__ unsigned char bgk_update(float &a, float &b, float ybar, float kbar, const DevParams &P) {
    a += ybar;
    b += kbar - ybar;
    const float var = (a * b) / ((a + b) * (a + b) * (a + b + 1.0f));
    if (var > P.var_thresh) return LA3DM_UNKNOWN;
    const float p = a / (a + b);
    return p > P.occupied_thresh ? LA3DM_OCCUPIED : (p < P.free_thresh ? LA3DM_FREE : LA3DM_UNKNOWN);
}

template <int kSlotsPerLane>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_predict_bgk(const NeighbourPlan *__restrict__ plan, const unsigned int *__restrict__ d_t,
              const float4 *__restrict__ pts, const long long *__restrict__ keys, float2 *__restrict__ ab,
              unsigned char *__restrict__ st, const float3 *__restrict__ lut, const DevParams *__restrict__ Pg,
              int nodes_pad, int shard_rank, int shard_world, ScanCounters *cnt) {
    __shared__ float4 tile[kWarpsPerCta][kTile];
    __shared__ DevParams Ps;
    if (threadIdx.x < sizeof(DevParams) / 4) reinterpret_cast<int *>(&Ps)[threadIdx.x] = reinterpret_cast<const int *>(Pg)[threadIdx.x];
    __syncthreads();
    const DevParams &P = Ps;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int T = *d_t;
    const unsigned int warps_total = gridDim.x * kWarpsPerCta;
    const int D = P.depth, finest_off = P.layer_off[D - 1], finest = P.finest;
    const float ell = P.ell, sf2 = P.sf2;

    unsigned long long visits = 0, updates = 0, pairs = 0;

    for (unsigned int t = blockIdx.x * kWarpsPerCta + warp; t < T; t += warps_total) {
        if (shard_world > 1 && (int) (t % (unsigned int) shard_world) != shard_rank) continue;
        const NeighbourPlan pl = plan[t];
        const size_t slot = pl.slot;
        float2 *bab = ab + slot * (size_t) P.nodes;
        unsigned char *bst = st + slot * (size_t) nodes_pad;

        if (pl.is_new) {   // fresh Block: every node = (prior_A, prior_B, UNKNOWN, !classified) (bgkoctree_node.h:34)
            for (int n = lane; n < P.nodes; n += 32) { bab[n] = make_float2(P.def_a, P.def_b); bst[n] = LA3DM_UNKNOWN; }
            __syncwarp();
        }
        // block centre from its key (hash_key_to_block, bgkblock.cpp:79-83)
        const long long key = keys[slot];
        const float cx = axis_center(key >> 40, P.block_size), cy = axis_center((key >> 20) & 0xFFFFF, P.block_size),
                    cz = axis_center(key & 0xFFFFF, P.block_size);

        // resolve this lane's leaves
        int node[kSlotsPerLane];
        float px[kSlotsPerLane], py[kSlotsPerLane], pz[kSlotsPerLane], a[kSlotsPerLane], b[kSlotsPerLane];
        unsigned char state[kSlotsPerLane], touched[kSlotsPerLane];
#pragma unroll
        for (int s = 0; s < kSlotsPerLane; ++s) {
            const int j = lane + 32 * s;
            node[s] = -1;
            touched[s] = 0;
            state[s] = LA3DM_UNKNOWN;
            a[s] = b[s] = px[s] = py[s] = pz[s] = 0.f;
            if (j < finest) {
                // walk up while PRUNED: leaf (d, i) is owned by the lane of its first finest descendant
                int d = D - 1, i = j, shift = 0;
                while (d > 0 && (bst[P.layer_off[d] + i] & 7) == kStPRUNED) { --d; i >>= 3; shift += 3; }
                const unsigned char sb = bst[P.layer_off[d] + i];
                const bool owner = ((i << shift) == j) && ((sb & 7) != kStPRUNED);
                if (owner) {
                    const int n = P.layer_off[d] + i;
                    node[s] = n;
                    state[s] = sb;
                    const float2 v = bab[n];
                    a[s] = v.x; b[s] = v.y;
                    const float3 off = lut[n];
                    // Block::get_loc: LUT offset + centre, then covSparse's  xs / ell
                    px[s] = (off.x + cx) / ell; py[s] = (off.y + cy) / ell; pz[s] = (off.z + cz) / ell;
                    ++visits;
                }
            }
        }

        // 7 neighbours in ExtendedBlock order, one Occupancy::update each (bgkoctomap.cpp:314-335)
        for (int nb = 0; nb < 7; ++nb) {
            const unsigned int cntp = pl.count[nb];
            if (cntp == 0) continue;
            const float4 *src = pts + pl.start[nb];
            float yb[kSlotsPerLane], kb[kSlotsPerLane];
#pragma unroll
            for (int s = 0; s < kSlotsPerLane; ++s) { yb[s] = 0.f; kb[s] = 0.f; }
            for (unsigned int base = 0; base < cntp; base += kTile) {
                const unsigned int m = min((unsigned int) kTile, cntp - base);
                __syncwarp();
                if ((unsigned int) lane < m) tile[warp][lane] = src[base + lane];
                __syncwarp();
                for (unsigned int q = 0; q < m; ++q) {
                    const float4 z = tile[warp][q];
#pragma unroll
                    for (int s = 0; s < kSlotsPerLane; ++s) {
                        if (node[s] < 0) continue;
                        const float dx = z.x - px[s], dy = z.y - py[s], dz = z.z - pz[s];
                        const float d = sqrtf(dx * dx + (dy * dy + dz * dz));   // Eigen rowwise().norm() of a 3-vector
                        if (d < 1.0f) {                                          // k <= 0 for d >= 1 (clamped upstream)
                            const float k = sparse_kernel(d, sf2);
                            yb[s] += k * z.w;
                            kb[s] += k;
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < kSlotsPerLane; ++s) {
                if (node[s] >= 0) {
                    pairs += cntp;
                    if (kb[s] > 0.0f) {                                          // bgkoctomap.cpp:332
                        state[s] = bgk_update(a[s], b[s], yb[s], kb[s], P) | 0x80;   // classified = true
                        touched[s] = 1;
                    }
                }
            }
        }

        // write back
#pragma unroll
        for (int s = 0; s < kSlotsPerLane; ++s) {
            if (node[s] >= 0 && touched[s]) {
                bab[node[s]] = make_float2(a[s], b[s]);
                bst[node[s]] = state[s];
                ++updates;
            }
        }
        __syncwarp();

        // OcTree::prune (bgkoctree.cpp:101-148): deepest layer first; 8 equal FREE/OCCUPIED siblings collapse into the
        // parent (copy of child 0's m_A, m_B, state -- `classified` is not copied, bgkoctree_node.h:40-45)
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

}  // namespace

void Map::predict() {
    if (last_T == 0) return;
    if (hp.method != LA3DM_BGK) throw StatusError{LA3DM_ERR_UNSUPPORTED, "predict: method not implemented yet"};
    const int ctas = num_sms * 4;
    const int slots = (hp.finest + 31) / 32;
    LA3DM_CUDA(cudaEventRecord(ev_p0, stream));
#define LAUNCH(S)                                                                                              \
    k_predict_bgk<S><<<ctas, kWarpsPerCta * 32, 0, stream>>>(                                                  \
        plan.as<NeighbourPlan>(), &d_cnt->n_test_blocks, pts_sorted.as<float4>(), keys.as<long long>(),        \
        ab.as<float2>(), st.as<unsigned char>(), d_lut, d_params, nodes_pad, shard_rank, shard_world, d_cnt)
    if (slots <= 1) LAUNCH(1);
    else if (slots <= 2) LAUNCH(2);
    else if (slots <= 16) LAUNCH(16);
    else throw StatusError{LA3DM_ERR_UNSUPPORTED, "block_depth > 4 not supported by the BGK kernel yet"};
#undef LAUNCH
    LA3DM_CUDA(cudaEventRecord(ev_p1, stream));
    ++launches;
}

}  // namespace la3dm_b200
