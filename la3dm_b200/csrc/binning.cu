// la3dm_b200 -- per-scan spatial binning.  Replaces the throw-away R-tree of the reference
// (include/common/rtree.h; Insert at src/bgkoctomap/bgkoctomap.cpp:240-243, box queries :497-552) and the TRAIN loop's
// bookkeeping (:250-284) with: closed-box block membership per training entry -> radix sort by block -> contiguous
// per-block ranges; test blocks = union of the 7-neighbourhoods of data blocks inside the float-stepped block grid
// (:486-495); lookup / creation of the blocks in the persistent device map; per-test-block neighbour plan.
#include <cub/cub.cuh>

#include "engine.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kThreads = 256;
constexpr unsigned int kInvalid = 0xFFFFFFFFu;

// ---- block grid of the scan (get_blocks_in_bbox) ------------------------------------------------------------------
__global__ void k_grid(const unsigned int *__restrict__ mm, float bs, GridDesc *g) {
    const int a = threadIdx.x;
    if (a < 3) {
        const float mn = float_unflip(mm[a]), mx = float_unflip(mm[3 + a]);
        long long first = 0, prev = 0;
        int steps = 0, irregular = 0, overflow = 0;
        const float hi = mx + 2 * bs;
        for (float x = mn - bs; x <= hi; x += bs) {
            const long long idx = axis_index(x, bs);
            if (steps == 0) first = idx;
            else if (idx != prev + 1) irregular = 1;
            const long long rel = idx - first;
            if (rel < 0 || rel >= kMaxAxis || steps >= kMaxAxis) { overflow = 1; break; }
            g->present[a][rel] = 1;
            prev = idx;
            ++steps;
        }
        g->base[a] = first;
        g->n[a] = steps == 0 ? 0 : (int) (prev - first + 1);
        if (irregular) atomicOr(&g->irregular, 1);
        if (overflow) atomicOr(&g->overflow, 1);
    }
    __syncthreads();
    if (a == 0) {
        const unsigned long long cells = (unsigned long long) g->n[0] * (unsigned long long) g->n[1] *
                                         (unsigned long long) g->n[2];
        if (cells >= 0xFFFFFFF0ull) g->overflow = 1;
    }
}

__global__ void k_grid_clear(GridDesc *g) {
    // present[] of the previous scan is cleared span-by-span by the caller's memset; reset the flags here
    g->irregular = 0;
    g->overflow = 0;
}

// candidate block indices of one coordinate: the closed box [c - bs/2, c + bs/2] in fp32 around the centres of the
// blocks i0-1, i0, i0+1 (src/bgkoctomap/bgkoctomap.cpp:497-503; closed intervals: rtree.h:1519-1532)
__device__ inline int axis_candidates(float q, float bs, float half, const GridDesc *g, int a, int out[3]) {
    const long long i0 = axis_index(q, bs);
    int n = 0;
#pragma unroll
    for (int k = -1; k <= 1; ++k) {
        const long long ii = i0 + k;
        const float c = axis_center(ii, bs);
        const float lo = c - half, hi = c + half;
        if (lo > q || q > hi) continue;
        const long long rel = ii - g->base[a];
        if (rel < 0 || rel >= g->n[a] || !g->present[a][rel]) continue;   // block not enumerated this scan
        out[n++] = (int) rel;
    }
    return n;
}

__global__ void k_member_count(const float4 *__restrict__ xy, const unsigned int *__restrict__ d_n,
                               const DevParams *__restrict__ P, const GridDesc *__restrict__ g, unsigned int *cnt,
                               unsigned int n_upper) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_upper) return;
    if (i >= *d_n) { cnt[i] = 0; return; }
    const float4 p = xy[i];
    int c[3];
    const int nx = axis_candidates(p.x, P->block_size, P->half_size, g, 0, c);
    const int ny = axis_candidates(p.y, P->block_size, P->half_size, g, 1, c);
    const int nz = axis_candidates(p.z, P->block_size, P->half_size, g, 2, c);
    cnt[i] = (unsigned int) (nx * ny * nz);
}

__global__ void k_member_total(const unsigned int *__restrict__ cnt, const unsigned int *__restrict__ off,
                               unsigned int n_upper, ScanCounters *c) {
    c->n_members = n_upper ? off[n_upper - 1] + cnt[n_upper - 1] : 0;
}

__global__ void k_member_fill(const float4 *__restrict__ xy, const unsigned int *__restrict__ d_n,
                              const DevParams *__restrict__ P, const GridDesc *__restrict__ g,
                              const unsigned int *__restrict__ off, unsigned int *keys, unsigned int *vals) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *d_n) return;
    const float4 p = xy[i];
    int cx[3], cy[3], cz[3];
    const int nx = axis_candidates(p.x, P->block_size, P->half_size, g, 0, cx);
    const int ny = axis_candidates(p.y, P->block_size, P->half_size, g, 1, cy);
    const int nz = axis_candidates(p.z, P->block_size, P->half_size, g, 2, cz);
    unsigned int o = off[i];
    for (int a = 0; a < nx; ++a)
        for (int b = 0; b < ny; ++b)
            for (int c = 0; c < nz; ++c) {
                keys[o] = ((unsigned int) cx[a] * (unsigned int) g->n[1] + (unsigned int) cy[b]) *
                              (unsigned int) g->n[2] + (unsigned int) cz[c];
                vals[o] = i;
                ++o;
            }
}

__global__ void k_heads(const unsigned int *__restrict__ keys, unsigned int n, unsigned int *flags) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (keys[i] != kInvalid && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
}

// compaction of run heads: id[rank] = key, start[rank] = i; sentinel start[total] = n_valid
__global__ void k_runs(const unsigned int *__restrict__ keys, const unsigned int *__restrict__ flags,
                       const unsigned int *__restrict__ ranks, unsigned int n, unsigned int *id, unsigned int *start,
                       unsigned int *d_total) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) { id[ranks[i]] = keys[i]; if (start) start[ranks[i]] = i; }
    // first invalid position (or n) terminates the last run
    const bool valid = keys[i] != kInvalid;
    const bool next_valid = (i + 1 < n) && keys[i + 1] != kInvalid;
    if (valid && !next_valid) {
        const unsigned int total = ranks[i] + flags[i];
        if (start) start[total] = i + 1;
        *d_total = total;
    }
    if (i == 0 && !valid) { if (start) start[0] = 0; *d_total = 0; }
}

// block-sorted training array, coordinates pre-divided by ell: covSparse's  x / ell  (bgkinference.h:114) hoisted
__global__ void k_gather_scaled(const float4 *__restrict__ xy, const unsigned int *__restrict__ vals, unsigned int n,
                                float ell, float4 *out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = xy[vals[i]];
    out[i] = make_float4(p.x / ell, p.y / ell, p.z / ell, p.w);
}

// 7 candidates per data block: the blocks whose ExtendedBlock contains it = itself and its 6 face neighbours
__global__ void k_candidates(const unsigned int *__restrict__ db_id, unsigned int d, const GridDesc *__restrict__ g,
                             unsigned int *cand) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const unsigned int id = db_id[i];
    const int nz = g->n[2], ny = g->n[1], nx = g->n[0];
    const int z = (int) (id % (unsigned int) nz), y = (int) ((id / (unsigned int) nz) % (unsigned int) ny),
              x = (int) (id / ((unsigned int) nz * (unsigned int) ny));
    const int dx[7] = {0, 1, -1, 0, 0, 0, 0}, dy[7] = {0, 0, 0, 1, -1, 0, 0}, dz[7] = {0, 0, 0, 0, 0, 1, -1};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int xx = x + dx[k], yy = y + dy[k], zz = z + dz[k];
        unsigned int out = kInvalid;
        if (xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz && g->present[0][xx] &&
            g->present[1][yy] && g->present[2][zz])
            out = ((unsigned int) xx * (unsigned int) ny + (unsigned int) yy) * (unsigned int) nz + (unsigned int) zz;
        cand[(size_t) i * 7 + k] = out;
    }
}

__device__ inline long long dense_to_key(unsigned int id, const GridDesc *g) {
    const unsigned int nz = (unsigned int) g->n[2], ny = (unsigned int) g->n[1];
    const long long z = id % nz, y = (id / nz) % ny, x = id / (nz * ny);
    return make_key(g->base[0] + x, g->base[1] + y, g->base[2] + z);
}

// ---- persistent block map -----------------------------------------------------------------------------------------
__device__ inline int hash_find(const long long *__restrict__ hkeys, const int *__restrict__ hvals, size_t mask,
                                long long key) {
    size_t h = (size_t) mix64((unsigned long long) key) & mask;
    while (true) {
        const long long k = hkeys[h];
        if (k == key) return hvals[h];
        if (k == -1) return -1;
        h = (h + 1) & mask;
    }
}

__device__ inline void hash_insert(long long *hkeys, int *hvals, size_t mask, long long key, int val) {
    size_t h = (size_t) mix64((unsigned long long) key) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS((unsigned long long *) &hkeys[h], (unsigned long long) -1LL,
                                                  (unsigned long long) key);
        if (prev == (unsigned long long) -1LL || prev == (unsigned long long) key) { hvals[h] = val; return; }
        h = (h + 1) & mask;
    }
}

__global__ void k_test_lookup(const unsigned int *__restrict__ test_id, const unsigned int *__restrict__ d_t,
                              unsigned int t_upper, const GridDesc *__restrict__ g,
                              const long long *__restrict__ hkeys, const int *__restrict__ hvals, size_t mask,
                              NeighbourPlan *plan, unsigned int *miss) {
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_upper) return;
    if (t >= *d_t) { miss[t] = 0; return; }
    const long long key = dense_to_key(test_id[t], g);
    const int slot = hash_find(hkeys, hvals, mask, key);
    plan[t].slot = (unsigned int) slot;
    plan[t].is_new = slot < 0 ? 1u : 0u;
    miss[t] = slot < 0 ? 1u : 0u;
}

__global__ void k_test_insert(const unsigned int *__restrict__ test_id, const unsigned int *__restrict__ d_t,
                              const GridDesc *__restrict__ g, const unsigned int *__restrict__ miss,
                              const unsigned int *__restrict__ rank, long long *hkeys, int *hvals, size_t mask,
                              long long *keys, const unsigned int *__restrict__ d_nblocks, NeighbourPlan *plan,
                              ScanCounters *c) {
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int T = *d_t;
    if (t >= T) return;
    if (t == T - 1) c->n_new_blocks = rank[t] + miss[t];
    if (!miss[t]) return;
    const unsigned int slot = *d_nblocks + rank[t];
    const long long key = dense_to_key(test_id[t], g);
    keys[slot] = key;
    hash_insert(hkeys, hvals, mask, key, (int) slot);
    plan[t].slot = slot;
}

__global__ void k_commit_blocks(unsigned int *d_nblocks, ScanCounters *c) {
    if (c->n_test_blocks == 0) c->n_new_blocks = 0;
    *d_nblocks += c->n_new_blocks;
}

// neighbour ranges: binary search of the neighbour's dense id in the sorted data-block list
__global__ void k_plan(const unsigned int *__restrict__ test_id, const unsigned int *__restrict__ d_t,
                       const GridDesc *__restrict__ g, const unsigned int *__restrict__ db_id,
                       const unsigned int *__restrict__ db_start, const unsigned int *__restrict__ d_d,
                       NeighbourPlan *plan) {
    const unsigned int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int t = gid / 8, k = gid % 8;
    if (t >= *d_t || k >= 7) return;
    const unsigned int id = test_id[t];
    const int nz = g->n[2], ny = g->n[1], nx = g->n[0];
    const int z = (int) (id % (unsigned int) nz), y = (int) ((id / (unsigned int) nz) % (unsigned int) ny),
              x = (int) (id / ((unsigned int) nz * (unsigned int) ny));
    // ExtendedBlock order: self, +x, -x, +y, -y, +z, -z  (src/bgkoctomap/bgkblock.cpp:85-101)
    const int dx[7] = {0, 1, -1, 0, 0, 0, 0}, dy[7] = {0, 0, 0, 1, -1, 0, 0}, dz[7] = {0, 0, 0, 0, 0, 1, -1};
    const int xx = x + dx[k], yy = y + dy[k], zz = z + dz[k];
    unsigned int start = 0, count = 0;
    if (xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz) {
        const unsigned int nid = ((unsigned int) xx * (unsigned int) ny + (unsigned int) yy) * (unsigned int) nz +
                                 (unsigned int) zz;
        unsigned int lo = 0, hi = *d_d;
        while (lo < hi) {
            const unsigned int mid = (lo + hi) >> 1;
            if (db_id[mid] < nid) lo = mid + 1; else hi = mid;
        }
        if (lo < *d_d && db_id[lo] == nid) { start = db_start[lo]; count = db_start[lo + 1] - start; }
    }
    plan[t].start[k] = start;
    plan[t].count[k] = count;
}

__global__ void k_hash_clear(long long *hkeys, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hkeys[i] = -1;
}

__global__ void k_hash_rebuild(const long long *__restrict__ keys, unsigned int n, long long *hkeys, int *hvals,
                               size_t mask) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hash_insert(hkeys, hvals, mask, keys[i], (int) i);
}

}  // namespace

// capacity for `blocks` blocks in the pool and a hash table at <= 50 % load
void Map::ensure_pool(size_t blocks) {
    if (blocks > pool_cap) {
        const size_t want = blocks + blocks / 2 + 1024;
        keys.grow_keep(want * sizeof(long long), stream);
        ab.grow_keep(want * (size_t) hp.nodes * sizeof(float2), stream);
        st.grow_keep(want * (size_t) nodes_pad, stream);
        pool_cap = want;
    }
    size_t need = 1024;
    while (need < 2 * pool_cap) need <<= 1;
    if (need > hash_cap) {
        hkeys.reserve(need * sizeof(long long), stream);
        hvals.reserve(need * sizeof(int), stream);
        hash_cap = need;
        k_hash_clear<<<ceil_div((long long) need, kThreads), kThreads, 0, stream>>>(hkeys.as<long long>(), need);
        if (n_blocks > 0)
            k_hash_rebuild<<<ceil_div(n_blocks, kThreads), kThreads, 0, stream>>>(
                keys.as<long long>(), (unsigned int) n_blocks, hkeys.as<long long>(), hvals.as<int>(), need - 1);
        launches += 2;
    }
}

// Input: xy[0..n_train) on the device, h_cnt->n_hits / n_raw_frees on the host (upper bound of n_train).
// Output: pts_sorted, db_id/db_start, test_id, plan[] (slot, is_new, 7 ranges); counters on the device.
void Map::bin_and_plan() {
    const unsigned int n_upper = h_cnt->n_hits + h_cnt->n_raw_frees;
    const float4 *d_xy = xy.as<float4>();

    // bbox of the training set (src/bgkoctomap/bgkoctomap.cpp:464-484) and the float-stepped block grid
    unsigned int *mm = d_mm + 12;
    minmax_points(reinterpret_cast<const float *>(d_xy), 4, 0, &d_cnt->n_train, mm, n_upper);
    LA3DM_CUDA(cudaMemsetAsync(d_grid, 0, sizeof(GridDesc), stream));
    k_grid<<<1, 32, 0, stream>>>(mm, hp.block_size, d_grid);

    // memberships
    mem_cnt.reserve((size_t) n_upper * 4, stream);
    mem_off.reserve((size_t) n_upper * 4, stream);
    const int g_n = ceil_div(n_upper, kThreads);
    k_member_count<<<g_n, kThreads, 0, stream>>>(d_xy, &d_cnt->n_train, d_params, d_grid, mem_cnt.as<unsigned int>(),
                                                 n_upper);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, mem_cnt.as<unsigned int>(), mem_off.as<unsigned int>(), (int) n_upper,
                                  stream);
    cub_tmp.reserve(tmp, stream);
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp, mem_cnt.as<unsigned int>(), mem_off.as<unsigned int>(),
                                             (int) n_upper, stream));
    k_member_total<<<1, 1, 0, stream>>>(mem_cnt.as<unsigned int>(), mem_off.as<unsigned int>(), n_upper, d_cnt);
    launches += 5;

    // sync #2: number of memberships (sizes the sort)
    d2h_bytes += sizeof(ScanCounters) + offsetof(GridDesc, present);
    LA3DM_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(ScanCounters), cudaMemcpyDeviceToHost, stream));
    GridDesc hg_head;   // only the small header is needed on the host
    LA3DM_CUDA(cudaMemcpyAsync(&hg_head, d_grid, offsetof(GridDesc, present), cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    if (hg_head.overflow) throw StatusError{LA3DM_ERR_EXTENT, "scan bounding box spans too many blocks"};
    stats.grid_irregular = hg_head.irregular;
    const unsigned int nm = h_cnt->n_members;
    if (nm == 0) {
        LA3DM_CUDA(cudaMemsetAsync(&d_cnt->n_data_blocks, 0, 4 * sizeof(unsigned int), stream));
        h_cnt->n_data_blocks = h_cnt->n_test_blocks = 0;
        last_T = 0;
        return;
    }

    for (int i = 0; i < 2; ++i) { sort_keys[i].reserve((size_t) nm * 4, stream); sort_vals[i].reserve((size_t) nm * 4, stream); }
    flags.reserve((size_t) std::max(nm, 8u) * 4 * 7, stream);
    ranks.reserve((size_t) std::max(nm, 8u) * 4 * 7, stream);
    k_member_fill<<<g_n, kThreads, 0, stream>>>(d_xy, &d_cnt->n_train, d_params, d_grid, mem_off.as<unsigned int>(),
                                                sort_keys[0].as<unsigned int>(), sort_vals[0].as<unsigned int>());
    cub::DoubleBuffer<unsigned int> dk(sort_keys[0].as<unsigned int>(), sort_keys[1].as<unsigned int>());
    cub::DoubleBuffer<unsigned int> dv(sort_vals[0].as<unsigned int>(), sort_vals[1].as<unsigned int>());
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, dk, dv, (int) nm, 0, 32, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, flags.as<unsigned int>(), ranks.as<unsigned int>(), (int) nm, stream);
    cub_tmp.reserve(std::max(t1, t2), stream);
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, t1, dk, dv, (int) nm, 0, 32, stream));
    const int g_m = ceil_div(nm, kThreads);
    k_heads<<<g_m, kThreads, 0, stream>>>(dk.Current(), nm, flags.as<unsigned int>());
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, t2, flags.as<unsigned int>(), ranks.as<unsigned int>(),
                                             (int) nm, stream));
    db_id.reserve((size_t) (nm + 1) * 4, stream);
    db_start.reserve((size_t) (nm + 2) * 4, stream);
    k_runs<<<g_m, kThreads, 0, stream>>>(dk.Current(), flags.as<unsigned int>(), ranks.as<unsigned int>(), nm,
                                         db_id.as<unsigned int>(), db_start.as<unsigned int>(),
                                         &d_cnt->n_data_blocks);
    pts_sorted.reserve((size_t) nm * sizeof(float4), stream);
    k_gather_scaled<<<g_m, kThreads, 0, stream>>>(d_xy, dv.Current(), nm, hp.ell, pts_sorted.as<float4>());
    launches += 5 + 6;

    // sync #3: number of data blocks
    d2h_bytes += 4;
    LA3DM_CUDA(cudaMemcpyAsync(&h_cnt->n_data_blocks, &d_cnt->n_data_blocks, sizeof(unsigned int),
                               cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    const unsigned int D = h_cnt->n_data_blocks;
    const unsigned int nc = D * 7;

    // candidates -> sort -> unique = test blocks
    for (int i = 0; i < 2; ++i) cand[i].reserve((size_t) nc * 4, stream);
    test_id.reserve((size_t) (nc + 1) * 4, stream);
    k_candidates<<<ceil_div(D, kThreads), kThreads, 0, stream>>>(db_id.as<unsigned int>(), D, d_grid,
                                                                 cand[0].as<unsigned int>());
    cub::DoubleBuffer<unsigned int> dc(cand[0].as<unsigned int>(), cand[1].as<unsigned int>());
    size_t t3 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, t3, dc, (int) nc, 0, 32, stream);
    cub_tmp.reserve(std::max(t3, t2), stream);
    LA3DM_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp.p, t3, dc, (int) nc, 0, 32, stream));
    const int g_c = ceil_div(nc, kThreads);
    k_heads<<<g_c, kThreads, 0, stream>>>(dc.Current(), nc, flags.as<unsigned int>());
    size_t t4 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t4, flags.as<unsigned int>(), ranks.as<unsigned int>(), (int) nc, stream);
    cub_tmp.reserve(t4, stream);
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, t4, flags.as<unsigned int>(), ranks.as<unsigned int>(),
                                             (int) nc, stream));
    k_runs<<<g_c, kThreads, 0, stream>>>(dc.Current(), flags.as<unsigned int>(), ranks.as<unsigned int>(), nc,
                                         test_id.as<unsigned int>(), nullptr, &d_cnt->n_test_blocks);
    launches += 3 + 6;

    // block lookup / creation.  T <= 7 D; the pool is sized for the bound so no further sync is needed.
    ensure_pool((size_t) n_blocks + nc);
    plan.reserve((size_t) nc * sizeof(NeighbourPlan), stream);
    miss.reserve((size_t) nc * 8, stream);
    unsigned int *d_miss = miss.as<unsigned int>(), *d_rank = d_miss + nc;
    k_test_lookup<<<g_c, kThreads, 0, stream>>>(test_id.as<unsigned int>(), &d_cnt->n_test_blocks, nc, d_grid,
                                                hkeys.as<long long>(), hvals.as<int>(), hash_cap - 1,
                                                plan.as<NeighbourPlan>(), d_miss);
    size_t t5 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t5, d_miss, d_rank, (int) nc, stream);
    cub_tmp.reserve(t5, stream);
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, t5, d_miss, d_rank, (int) nc, stream));
    k_test_insert<<<g_c, kThreads, 0, stream>>>(test_id.as<unsigned int>(), &d_cnt->n_test_blocks, d_grid, d_miss,
                                                d_rank, hkeys.as<long long>(), hvals.as<int>(), hash_cap - 1,
                                                keys.as<long long>(), d_nblocks, plan.as<NeighbourPlan>(), d_cnt);
    k_commit_blocks<<<1, 1, 0, stream>>>(d_nblocks, d_cnt);
    k_plan<<<ceil_div((long long) nc * 8, kThreads), kThreads, 0, stream>>>(
        test_id.as<unsigned int>(), &d_cnt->n_test_blocks, d_grid, db_id.as<unsigned int>(),
        db_start.as<unsigned int>(), &d_cnt->n_data_blocks, plan.as<NeighbourPlan>());
    launches += 6;
    last_T = nc;   // upper bound until the counters are read back
}

}  // namespace la3dm_b200
