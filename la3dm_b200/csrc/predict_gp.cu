// la3dm_b200 -- GPOctoMap: per-data-block GP training and the fused predict -> Occupancy::update -> prune.
//
// Replaces the TRAIN and PREDICT loops of GPOctoMap::insert_pointcloud (src/gpoctomap/gpoctomap.cpp:241-333):
//   GPRegressor::train   (include/gpoctomap/gpregressor.h:42-51):  K = sf2 (1 + r) exp(-r), r = |sqrt(3)/ell (xi - xj)|,
//                         K += noise I, LL^T = K, alpha = K^-1 y
//   GPRegressor::predict (:80-92):  Ks = k(x, xs); m = Ks^T alpha; v = L^-1 Ks; var = sf2 - diag(v^T v)
//   Occupancy::update    (src/gpoctomap/gpoctree_node.cpp:36-49), applied for EVERY leaf of a test block
//                         (gpoctomap.cpp:317: no kbar-style guard), prune as in the BGK maps.
//
// cond(K + noise I) ~ 1e4 in fp32: a different order of the fp32 operations moves the occupancy probability by more
// than the 1e-4 parity bound, so both kernels keep the order of the scalar formulation (every dot product is
// accumulated in ascending index order, products and differences rounded separately); the parallelism is over rows /
// leaves and over blocks, not inside a dot product.  exp() is evaluated in double and rounded once to fp32.
#include <cub/cub.cuh>

#include "block_common.cuh"
#include "gp_common.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kGpWarps = 4;           // warps per CTA in both kernels
constexpr int kGpSmemN = 64;          // k_gp_train factorises blocks of up to this many points in shared memory
constexpr int kGpWarpN = 32;          // ... and leaves the blocks above this size (up to kGpBigN) to k_gp_train_big

// floats of factor storage per data block: packed lower triangle of L, then alpha
__global__ void k_gp_sizes(const unsigned int *__restrict__ db_start, const ScanCounters *__restrict__ c,
                           unsigned int cap, unsigned long long *sizes) {
    const unsigned int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= cap) return;
    unsigned long long s = 0;
    if (!c->overflow && d < c->n_data_blocks) {
        const unsigned long long n = db_start[d + 1] - db_start[d];
        s = n * (n + 1) / 2 + n;
    }
    sizes[d] = s;
}

__global__ void k_gp_check(const unsigned long long *__restrict__ off, const unsigned long long *__restrict__ sizes,
                           ScanCounters *c, unsigned int cap, unsigned long long store_cap) {
    if (c->overflow) return;
    const unsigned int D = min(c->n_data_blocks, cap);
    const unsigned long long total = D ? off[D - 1] + sizes[D - 1] : 0ull;
    c->gp_store_needed = total;
    if (total > store_cap) atomicOr(&c->overflow, OVF_GPSTORE);
}

// Backward substitution alpha = L^-T z in place, by one warp: s = alpha_i - sum_{k > i, ascending} L_ki alpha_k, / L_ii.
// The order of the subtractions is the scalar one; what the lanes share is everything else: the products L_ki alpha_k of 32
// terms are formed in parallel (a lane each, rounded once like the scalar product) and handed round by shuffles, so the
// dependent chain is one subtraction per term instead of an address computation, two loads, a product and a subtraction.
__device__ __forceinline__ void gp_backward_warp(const float *L, float *alpha, unsigned int n, int lane) {
    for (unsigned int ii = n; ii-- > 0;) {
        float s = alpha[ii];
        for (unsigned int k0 = ii + 1; k0 < n; k0 += 32) {
            const unsigned int k = k0 + (unsigned int) lane;
            float p = 0.f;
            if (k < n) p = L[(size_t) k * (k + 1) / 2 + ii] * alpha[k];
            const unsigned int cnt = min(32u, n - k0);
            if (cnt == 32u) {
#pragma unroll
                for (int q = 0; q < 32; ++q) s -= __shfl_sync(0xffffffffu, p, q);
            } else {
                for (unsigned int q = 0; q < cnt; ++q) s -= __shfl_sync(0xffffffffu, p, (int) q);
            }
        }
        s = s / L[(size_t) ii * (ii + 1) / 2 + ii];
        __syncwarp();
        if (lane == 0) alpha[ii] = s;
        __syncwarp();
    }
}

// GPRegressor::train, one warp per data block
__global__ void __launch_bounds__(kGpWarps * 32)
k_gp_train(const float4 *__restrict__ pts, const unsigned int *__restrict__ db_start,
           const unsigned long long *__restrict__ off, float *store, const DevParams *__restrict__ Pg,
           const ScanCounters *__restrict__ c, unsigned int big_max) {
    __shared__ float sL[kGpWarps][kGpSmemN * (kGpSmemN + 1) / 2 + kGpSmemN];
    if (c->overflow) return;
    const int lane = threadIdx.x & 31;
    const unsigned int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_w = (gridDim.x * blockDim.x) >> 5;
    const float sf2 = Pg->sf2, noise = Pg->noise;
    const unsigned int D = c->n_data_blocks;
    for (unsigned int d = gw; d < D; d += n_w) {
        const unsigned int first = db_start[d], n = db_start[d + 1] - first;
        if (n > (unsigned int) kGpWarpN && n <= big_max) continue;      // k_gp_train_big's
        const float4 *x = pts + first;
        float *Lg = store + off[d];                      // L[i][j] at i (i + 1) / 2 + j, then alpha
        // a block of up to kGpSmemN points is factorised in shared memory (the column sweeps are chains of dependent
        // loads: ~25 cycles each there, several hundred from L2) and copied out at the end
        const bool staged = n <= (unsigned int) kGpSmemN;
        float *L = staged ? sL[threadIdx.x >> 5] : Lg;
        float *alpha = L + (size_t) n * (n + 1) / 2;
        // K + noise I, lower triangle (:44-46)
        for (unsigned int i = 0; i < n; ++i) {
            const float4 xi = x[i];
            float *row = L + (size_t) i * (i + 1) / 2;
            for (unsigned int j = lane; j <= i; j += 32) {
                const float4 xj = x[j];
                float k = matern3(xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, sf2);
                if (i == j) k = k + noise * 1.0f;
                row[j] = k;
            }
        }
        __syncwarp();
        // Cholesky (LLT, :47): column by column, a lane per row; each element's sum runs in ascending k
        for (unsigned int j = 0; j < n; ++j) {
            const float *rj = L + (size_t) j * (j + 1) / 2;
            float ljj = 0.f;
            for (unsigned int i0 = j; i0 < n; i0 += 32) {
                const unsigned int i = i0 + lane;
                float s = 0.f;
                float *ri = nullptr;
                if (i < n) {
                    ri = L + (size_t) i * (i + 1) / 2;
                    s = ri[j];
                    unsigned int k = 0;
                    for (; k + 4 <= j; k += 4) {        // loads of four terms together, subtractions in ascending k
                        const float a0 = ri[k], a1 = ri[k + 1], a2 = ri[k + 2], a3 = ri[k + 3];
                        const float b0 = rj[k], b1 = rj[k + 1], b2 = rj[k + 2], b3 = rj[k + 3];
                        s -= a0 * b0; s -= a1 * b1; s -= a2 * b2; s -= a3 * b3;
                    }
                    for (; k < j; ++k) s -= ri[k] * rj[k];
                }
                if (i0 == j) {       // lane 0 holds the diagonal element
                    const float dg = sqrtf(s);
                    ljj = __shfl_sync(0xffffffffu, dg, 0);
                    if (lane == 0) s = dg;
                    else if (i < n) s = s / ljj;
                } else if (i < n) s = s / ljj;
                __syncwarp();        // every read of row j's old element is done before the diagonal is overwritten
                if (i < n) ri[j] = s;
            }
            __syncwarp();
        }
        // alpha = L^-T (L^-1 y) (:48-49).  Forward: column oriented, every s_i is reduced in ascending k.
        for (unsigned int i = lane; i < n; i += 32) alpha[i] = x[i].w;
        __syncwarp();
        for (unsigned int k = 0; k < n; ++k) {
            float ak = 0.f;
            if (lane == 0) { ak = alpha[k] / L[(size_t) k * (k + 1) / 2 + k]; alpha[k] = ak; }
            ak = __shfl_sync(0xffffffffu, ak, 0);
            for (unsigned int i = k + 1 + lane; i < n; i += 32) alpha[i] -= L[(size_t) i * (i + 1) / 2 + k] * ak;
            __syncwarp();
        }
        // Backward: s = alpha_i - sum_{k > i, ascending} L_ki alpha_k needs every later alpha first
        gp_backward_warp(L, alpha, n, lane);
        __syncwarp();
        if (staged) {
            const unsigned int words = n * (n + 1) / 2 + n;
            for (unsigned int w = lane; w < words; w += 32) Lg[w] = L[w];
            __syncwarp();
        }
    }
}


// GPRegressor::train for the data blocks of more than kGpWarpN points (up to kGpBigN): a CTA per block, the factor in
// shared memory, a thread per row.  Every element's sum runs in ascending k exactly like k_gp_train (same results, bit
// for bit); what changes is that the rows of a column are computed by 256 threads instead of 32 lanes, out of shared
// memory instead of L2 -- left to one warp these few blocks outlast the rest of the scan's training.
constexpr int kGpBigN = 224;
constexpr int kGpBigThreads = 256;
constexpr size_t kGpBigSmem = ((size_t) kGpBigN * (kGpBigN + 1) / 2 + kGpBigN) * sizeof(float);

__global__ void __launch_bounds__(kGpBigThreads)
k_gp_train_big(const float4 *__restrict__ pts, const unsigned int *__restrict__ db_start,
               const unsigned long long *__restrict__ off, float *store, const DevParams *__restrict__ Pg,
               const ScanCounters *__restrict__ c) {
    extern __shared__ __align__(16) float gp_big_smem[];
    __shared__ float s_ljj;
    if (c->overflow) return;
    const float sf2 = Pg->sf2, noise = Pg->noise;
    const unsigned int D = c->n_data_blocks;
    const unsigned int tid = threadIdx.x;
    for (unsigned int d = blockIdx.x; d < D; d += gridDim.x) {
        const unsigned int first = db_start[d], n = db_start[d + 1] - first;
        if (n <= (unsigned int) kGpWarpN || n > (unsigned int) kGpBigN) continue;
        const float4 *x = pts + first;
        float *Lg = store + off[d];
        float *L = gp_big_smem;
        float *alpha = L + (size_t) n * (n + 1) / 2;
        __syncthreads();
        // K + noise I, lower triangle (:44-46)
        const unsigned int tri = n * (n + 1) / 2;
        for (unsigned int e = tid; e < tri; e += kGpBigThreads) {
            // (i, j) of the packed index e: i = floor((sqrt(8 e + 1) - 1) / 2), corrected for rounding
            unsigned int i = (unsigned int) ((sqrtf(8.0f * (float) e + 1.0f) - 1.0f) * 0.5f);
            while (i * (i + 1) / 2 > e) --i;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            const unsigned int j = e - i * (i + 1) / 2;
            const float4 xi = x[i], xj = x[j];
            float k = matern3(xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, sf2);
            if (i == j) k = k + noise * 1.0f;
            L[e] = k;
        }
        __syncthreads();
        // Cholesky (LLT, :47): column by column, a thread per row; each element's sum runs in ascending k
        for (unsigned int j = 0; j < n; ++j) {
            const float *rj = L + (size_t) j * (j + 1) / 2;
            const unsigned int i = j + tid;
            float s = 0.f;
            float *ri = nullptr;
            if (i < n) {
                ri = L + (size_t) i * (i + 1) / 2;
                s = ri[j];
                unsigned int k = 0;
                for (; k + 4 <= j; k += 4) {
                    const float a0 = ri[k], a1 = ri[k + 1], a2 = ri[k + 2], a3 = ri[k + 3];
                    const float b0 = rj[k], b1 = rj[k + 1], b2 = rj[k + 2], b3 = rj[k + 3];
                    s -= a0 * b0; s -= a1 * b1; s -= a2 * b2; s -= a3 * b3;
                }
                for (; k < j; ++k) s -= ri[k] * rj[k];
            }
            if (tid == 0) { s = sqrtf(s); s_ljj = s; }
            __syncthreads();                 // the diagonal is known; every read of row j's old elements is done
            if (i < n) ri[j] = tid == 0 ? s : s / s_ljj;
            __syncthreads();
        }
        // alpha = L^-T (L^-1 y) (:48-49).  Forward: column oriented, every s_i is reduced in ascending k.
        for (unsigned int i = tid; i < n; i += kGpBigThreads) alpha[i] = x[i].w;
        __syncthreads();
        for (unsigned int k = 0; k < n; ++k) {
            if (tid == 0) alpha[k] = alpha[k] / L[(size_t) k * (k + 1) / 2 + k];
            __syncthreads();
            const float ak = alpha[k];
            const unsigned int i = k + 1 + tid;
            if (i < n) alpha[i] -= L[(size_t) i * (i + 1) / 2 + k] * ak;
            __syncthreads();
        }
        // Backward: s = alpha_i - sum_{k > i, ascending} L_ki alpha_k needs every later alpha first: one warp
        if (tid < 32) gp_backward_warp(L, alpha, n, (int) tid);
        __syncthreads();
        for (unsigned int w = tid; w < tri + n; w += kGpBigThreads) Lg[w] = L[w];
    }
}

// Occupancy::update (gpoctree_node.cpp:36-49): a = m_ivar, b = ivar
__device__ __forceinline__ unsigned char gp_update(float &a, float &b, float m, float var, const DevParams &P,
                                                   unsigned char old_state) {
    b = (float) ((double) b + (1.0 / (double) var - (double) P.sf2));
    a += m / var;
    if (b < P.min_known_ivar) return LA3DM_UNKNOWN;
    b = b > P.max_ivar ? P.max_ivar : b;
    const float p = 1.0f / (1.0f + (float) exp((double) (-P.l * a / P.max_ivar)));
    (void) old_state;
    return p > P.occupied_thresh ? LA3DM_OCCUPIED : (p < P.free_thresh ? LA3DM_FREE : LA3DM_UNKNOWN);
}

// test blocks per (k_gp_mv, k_gp_apply) pair: bounds the (mean, variance) buffer at 64 leaves x 32768 blocks x 7 x 8 B
inline unsigned int gp_chunk(const DevParams &P) {
    const unsigned int groups = (unsigned int) (P.finest + 31) / 32;
    return std::max(1u, 65536u / groups);
}

// GPRegressor::predict for one (test block, neighbour, group of 32 leaves): a warp per unit, a lane per leaf.
//   ks (n kernel values), m = ks . alpha, forward substitution with L, var = sf2 - v . v   (gpregressor.h:80-92)
// Nothing here depends on the order of the neighbours, so the 7 x groups units of a test block run in parallel; the
// sequential part -- Occupancy::update neighbour after neighbour -- is k_gp_apply.
// `scratch` holds ks / v per leaf: [warp][2 arrays][n_max][32 leaves]; mv: [block - t0][7][32 groups] (mean, variance).
__global__ void __launch_bounds__(kGpWarps * 32)
k_gp_mv(const NeighbourPlan *__restrict__ plan, const unsigned int *__restrict__ plan_db,
        const float4 *__restrict__ pts, const unsigned long long *__restrict__ off, const float *__restrict__ store,
        const long long *__restrict__ keys, const unsigned char *__restrict__ pool, const float3 *__restrict__ lut,
        const DevParams *__restrict__ Pg, const ScanArgs *__restrict__ A, const ScanCounters *__restrict__ cnt,
        float *scratch, unsigned int n_max, unsigned int t0, unsigned int chunk, float2 *mv, unsigned int tc_max_n) {
    __shared__ DevParams Ps;
    load_params(Ps, Pg);
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int T = cnt->n_test_blocks;
    if (t0 >= T) return;
    const unsigned int t1 = min(T, t0 + chunk);
    const unsigned int gw = blockIdx.x * kGpWarps + warp, n_w = gridDim.x * kGpWarps;
    float *ks = scratch + (size_t) gw * 2 * n_max * 32;       // [n_max][32]
    float *vv = ks + (size_t) n_max * 32;
    const float sf2 = P.sf2, bs = P.block_size;
    const float scale = (float) (1.73205 / (double) P.ell);    // gpregressor.h:115
    const unsigned int shard_world = (unsigned int) A->shard_world, shard_rank = (unsigned int) A->shard_rank;
    const unsigned int groups = (unsigned int) (P.finest + 31) / 32, per_block = 7u * groups;
    const unsigned int units = (t1 - t0) * per_block;
    const int pruned = P.pruned_state;
    for (unsigned int u = gw; u < units; u += n_w) {
        const unsigned int t = t0 + u / per_block, r = u % per_block;
        const int nb = (int) (r / groups), s = (int) (r % groups);
        if (t % shard_world != shard_rank) continue;
        const NeighbourPlan *pl = plan + t;
        const unsigned int n = pl->count[nb];
        if (n == 0 || n <= tc_max_n) continue;           // (small regressors: the tensor-core kernel, predict_gp_tc.cu)
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
        if (node >= 0) {
            const long long key = keys[slot];
            const float cx = axis_center(key >> 40, bs), cy = axis_center((key >> 20) & 0xFFFFF, bs),
                        cz = axis_center(key & 0xFFFFF, bs);
            const float3 o = lut[node];
            // Block::get_loc, then predict()'s  scale * xs  (:84-85 with the stand-in's operand order)
            const float qx = scale * (o.x + cx), qy = scale * (o.y + cy), qz = scale * (o.z + cz);
            const float4 *x = pts + pl->start[nb];
            const float *L = store + off[plan_db[(size_t) t * 8 + nb] - 1];
            const float *alpha = L + (size_t) n * (n + 1) / 2;
            float mu = 0.f;
            for (unsigned int i = 0; i < n; ++i) {
                const float4 xi = x[i];
                const float k = matern3(xi.x, xi.y, xi.z, qx, qy, qz, sf2);
                ks[(size_t) i * 32 + lane] = k;
                mu += k * alpha[i];
            }
            float v2 = 0.f;
            for (unsigned int i = 0; i < n; ++i) {
                const float *ri = L + (size_t) i * (i + 1) / 2;
                float sacc = ks[(size_t) i * 32 + lane];
                for (unsigned int k = 0; k < i; ++k) sacc -= ri[k] * vv[(size_t) k * 32 + lane];
                const float vi = sacc / ri[i];
                vv[(size_t) i * 32 + lane] = vi;
                v2 += vi * vi;
            }
            mv[((size_t) (t - t0) * 7 + nb) * (groups * 32) + j] = make_float2(mu, sf2 - v2);
        }
    }
}

// one warp per test block, the record staged in (dynamic) shared memory; leaf after leaf (32 at a time, a lane each):
// Occupancy::update with the (mean, variance) of every neighbour that has a trained regressor, in ExtendedBlock order
// (gpoctomap.cpp:305-319); then prune and write back
__global__ void __launch_bounds__(kGpWarps * 32)
k_gp_apply(const NeighbourPlan *__restrict__ plan, unsigned char *__restrict__ pool, const DevParams *__restrict__ Pg,
           const ScanArgs *__restrict__ A, ScanCounters *cnt, unsigned int t0, unsigned int chunk,
           const float2 *__restrict__ mv, int staged) {
    extern __shared__ __align__(16) unsigned char gp_smem_raw[];
    __shared__ DevParams Ps;
    load_params(Ps, Pg);
    if (cnt->overflow) return;
    const DevParams &P = Ps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the record is staged in shared memory when it fits (block_depth <= 4: 5.3 KB), updated in place in global memory
    // otherwise (42 KB at depth 5)
    uint4 *smem_rec = reinterpret_cast<uint4 *>(gp_smem_raw + (size_t) warp * (staged ? P.rec_bytes : 0));
    const unsigned int T = cnt->n_test_blocks;
    if (t0 >= T) return;
    const unsigned int t1 = min(T, t0 + chunk);
    const unsigned int gw = blockIdx.x * kGpWarps + warp, n_w = gridDim.x * kGpWarps;
    const unsigned int shard_world = (unsigned int) A->shard_world, shard_rank = (unsigned int) A->shard_rank;
    const int groups = (P.finest + 31) / 32;
    const int pruned = P.pruned_state;
    unsigned long long visits = 0, updates = 0, pairs = 0;

    for (unsigned int t = t0 + gw; t < t1; t += n_w) {
        if (t % shard_world != shard_rank) continue;
        const NeighbourPlan pl = plan[t];
        uint4 *grec = reinterpret_cast<uint4 *>(pool + (size_t) pl.slot * (size_t) P.rec_bytes);
        uint4 *srec = staged ? smem_rec : grec;
        float2 *rab = reinterpret_cast<float2 *>(srec);
        unsigned char *rst = reinterpret_cast<unsigned char *>(srec) + P.st_off;
        __syncwarp();
        if (staged || pl.is_new) stage_record(srec, grec, pl.is_new != 0, P, lane);
        __syncwarp();
        unsigned int n_total = 0;
        for (int nb = 0; nb < 7; ++nb) n_total += pl.count[nb];
        bool any = false;
        for (int s = 0; s < groups; ++s) {
            // leaf of finest slot j (is_leaf, gpoctree.cpp: the leaf (d, i) is handled at its first finest descendant)
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
            unsigned char state = rst[node];
            for (int nb = 0; nb < 7; ++nb) {
                if (pl.count[nb] == 0) continue;
                const float2 m = mv[((size_t) (t - t0) * 7 + nb) * (size_t) (groups * 32) + j];
                state = gp_update(v.x, v.y, m.x, m.y, P, state) | 0x80;     // gpoctomap.cpp:317
            }
            // (a coarse leaf's node index is below every finest node's: no lane reads what another lane writes here)
            rab[node] = v;
            rst[node] = state;
            ++updates;
            any = true;
        }
        const bool dirty = __any_sync(0xffffffffu, any) || pl.is_new;
        __syncwarp();
        if (dirty) {
            prune_record(rab, rst, P, lane);
            if (staged) for (int w = lane; w < (P.rec_bytes >> 4); w += 32) grec[w] = srec[w];
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

// largest data block of the scan (bounds the per-leaf scratch)
__global__ void k_gp_nmax(const unsigned int *__restrict__ db_start, ScanCounters *c, unsigned int cap,
                          unsigned int n_max_cap) {
    if (c->overflow) return;
    const unsigned int d = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int n = 0;
    if (d < cap && d < c->n_data_blocks) n = db_start[d + 1] - db_start[d];
    for (int o = 16; o > 0; o >>= 1) n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
    if ((threadIdx.x & 31) == 0 && n) {
        atomicMax(&c->gp_n_max, n);
        if (n > n_max_cap) atomicOr(&c->overflow, OVF_GPN);
    }
}

}  // namespace

int gp_tc_max_n();   // predict_gp_tc.cu

// storage offsets of the regressors and the capacity checks (runs before the map is touched)
void Map::enqueue_gp_sizes() {
    // k_gp_apply stages a block record per warp in shared memory: 5.3 KB at depth 4, 42 KB at depth 5
    const unsigned int cap = caps.members;     // data blocks <= memberships
    const int grid = ceil_div(cap, 256);
    unsigned long long *sizes = gp_sizes.as<unsigned long long>(), *off = gp_off.as<unsigned long long>();
    k_gp_sizes<<<grid, 256, 0, stream>>>(db_start.as<unsigned int>(), d_cnt, cap, sizes);
    k_gp_nmax<<<grid, 256, 0, stream>>>(db_start.as<unsigned int>(), d_cnt, cap, caps.gp_n_max);
    size_t tmp = cub_tmp_bytes;
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp, sizes, off, (int) cap, stream));
    k_gp_check<<<1, 1, 0, stream>>>(off, sizes, d_cnt, cap, (unsigned long long) caps.gp_store);
    launches += 5;
}

void Map::enqueue_gp() {
    const unsigned long long *off = gp_off.as<unsigned long long>();
    const int ctas = num_sms * 4;
    record_event(ev_p0);
    // the few big data blocks first (a CTA each, they are the long poles), then everything else (a warp each)
    {
        static bool attr_done[64] = {};
        if (device < 64 && !attr_done[device]) {
            LA3DM_CUDA(cudaFuncSetAttribute(k_gp_train_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kGpBigSmem));
            attr_done[device] = true;
        }
    }
    k_gp_train_big<<<num_sms, kGpBigThreads, kGpBigSmem, stream>>>(pts_sorted.as<float4>(), db_start.as<unsigned int>(), off,
                                                                  gp_store.as<float>(), d_params, d_cnt);
    k_gp_train<<<ctas, kGpWarps * 32, 0, stream>>>(pts_sorted.as<float4>(), db_start.as<unsigned int>(), off,
                                                   gp_store.as<float>(), d_params, d_cnt, (unsigned int) kGpBigN);
    ++launches;
    // predict: (mean, variance) of every (test block, neighbour, leaf) in parallel, then the sequential fusion per block
    const unsigned int chunk = gp_chunk(hp);
    const size_t rec_smem = (size_t) hp.rec_bytes * kGpWarps;
    const int staged = rec_smem <= 40 * 1024 ? 1 : 0;          // block_depth <= 4 (5.3 KB per record)
    const size_t apply_smem = staged ? rec_smem : 0;
    // regressors of up to gp_tc_max_n() points go through tcgen05.mma (block_depth <= 3: 64 finest leaves per block);
    // LA3DM_GP_SIMT=1 keeps everything on the scalar kernel (A/B, and the fallback if TMEM cannot be had)
    static const bool simt_only = getenv("LA3DM_GP_SIMT") != nullptr;
    const bool use_tc = !simt_only && hp.depth <= 3;
    const unsigned int tc_max_n = use_tc ? (unsigned int) gp_tc_max_n() : 0u;
    for (unsigned int t0 = 0; t0 < caps.tests; t0 += chunk) {
        if (use_tc) enqueue_gp_mv_tc(t0, chunk);
        k_gp_mv<<<gp_ctas, kGpWarps * 32, 0, stream>>>(plan.as<NeighbourPlan>(), plan_db.as<unsigned int>(),
                                                       pts_sorted.as<float4>(), off, gp_store.as<float>(),
                                                       keys.as<long long>(), pool.as<unsigned char>(), d_lut, d_params,
                                                       d_args, d_cnt, gp_scratch.as<float>(), caps.gp_n_max, t0, chunk,
                                                       gp_mv.as<float2>(), tc_max_n);
        k_gp_apply<<<gp_ctas, kGpWarps * 32, apply_smem, stream>>>(plan.as<NeighbourPlan>(), pool.as<unsigned char>(),
                                                                   d_params, d_args, d_cnt, t0, chunk,
                                                                   gp_mv.as<float2>(), staged);
        launches += 2;
    }
    record_event(ev_p1);
    launches += 1;
}

size_t scan_temp_bytes(unsigned int items) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (unsigned long long *) nullptr, (unsigned long long *) nullptr,
                                  (int) items, nullptr);
    return tmp;
}

}  // namespace la3dm_b200
