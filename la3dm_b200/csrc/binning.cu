// la3dm_b200 -- per-scan spatial binning.  Replaces the throw-away R-tree of the reference
// (include/common/rtree.h; Insert at src/bgkoctomap/bgkoctomap.cpp:240-243, box queries :497-552) and the TRAIN loop's
// bookkeeping (:250-284) with: closed-box block membership per training entry -> radix sort by block -> contiguous
// per-block ranges; test blocks = union of the 7-neighbourhoods of data blocks inside the float-stepped block grid
// (:486-495), found through a bit per cell of the scan's dense block grid; lookup / creation of the blocks in the
// persistent device map; per-test-block neighbour plan.
//
// Like the front-end, nothing here synchronises with the host (see frontend.cu).
#include <cub/cub.cuh>

#include "engine.cuh"
#include "hash.cuh"
#include "runs.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kThreads = 256;
constexpr unsigned int kPad = 0xFFFFFFFFu;

// (the bounding box of the training set, src/bgkoctomap/bgkoctomap.cpp:464-484, is accumulated by the front-end kernels
// that write xy: k_hit_fill, k_vg_centroid<1>, k_vg_long<1>)

// ---- block grid of the scan (get_blocks_in_bbox, :486-495); *g was zeroed by a memset ------------------------------
// Three groups of 128 threads, one per axis.  The stepping  x += block_size  is a chain of dependent fp32 additions and
// stays sequential (one thread per axis writes the positions to shared memory, kGridChunk at a time); the block index of
// every position (a double division) and the bookkeeping are done by the whole group.
constexpr int kGridChunk = 2048;
constexpr int kGridThreads = 384;

__global__ void __launch_bounds__(kGridThreads)
k_grid(const unsigned int *__restrict__ mm, const DevParams *__restrict__ P, GridDesc *g, ScanCounters *c,
       unsigned int cells_cap) {
    __shared__ float xs[3][kGridChunk];
    __shared__ float s_x[3];
    __shared__ long long s_last[3];
    __shared__ int s_cnt[3], s_steps[3], s_done[3], s_bad, s_irr;
    const bool live = !c->overflow && c->n_train > 0;
    const int a = threadIdx.x >> 7, t = threadIdx.x & 127;
    const float bs = P->block_size;
    const float mn = float_unflip(mm[a]), mx = float_unflip(mm[3 + a]);
    const float hi = mx + 2 * bs;
    const long long first = axis_index(mn - bs, bs);
    if (threadIdx.x == 0) { s_bad = 0; s_irr = 0; }
    if (t == 0) { s_x[a] = mn - bs; s_steps[a] = 0; s_done[a] = live ? 0 : 1; s_last[a] = first - 1; }
    __syncthreads();
    while (!(s_done[0] && s_done[1] && s_done[2]) && !s_bad) {
        if (t == 0) {
            int n = 0;
            if (!s_done[a]) {
                float x = s_x[a];
                while (n < kGridChunk && x <= hi) { xs[a][n++] = x; x += bs; }
                s_x[a] = x;
            }
            s_cnt[a] = n;
        }
        __syncthreads();
        const int n = s_cnt[a], base = s_steps[a];
        const long long last_before = s_last[a];
        __syncthreads();
        for (int j = t; j < n; j += 128) {
            const long long idx = axis_index(xs[a][j], bs);
            const long long prev = j ? axis_index(xs[a][j - 1], bs) : last_before;
            if (base + j > 0 && idx != prev + 1) s_irr = 1;
            const long long rel = idx - first;
            if (rel < 0 || rel >= kMaxAxis || base + j >= kMaxAxis) s_bad = 1;
            else g->present[a][rel] = 1;
            if (j == n - 1) s_last[a] = idx;
        }
        if (t == 0) {
            s_steps[a] = base + n;
            if (n < kGridChunk) s_done[a] = 1;          // the stepping passed max + 2 block_size
        }
        __syncthreads();
    }
    if (live && t == 0) {
        g->base[a] = first;
        g->n[a] = s_steps[a] == 0 ? 0 : (int) (s_last[a] - first + 1);
    }
    __syncthreads();
    if (live && threadIdx.x == 0) {
        if (s_irr) g->irregular = 1;
        const unsigned long long cells = (unsigned long long) g->n[0] * (unsigned long long) g->n[1] *
                                         (unsigned long long) g->n[2];
        if (s_bad || cells >= 0x7FFFFFF0ull) atomicOr(&c->overflow, OVF_EXTENT);
        else {
            g->n_cells = (unsigned int) cells;
            c->n_cells = (unsigned int) cells;
            c->grid_irregular = (unsigned int) g->irregular;
            if (cells > (unsigned long long) cells_cap) atomicOr(&c->overflow, OVF_CELLS);
        }
    }
}

// candidate block indices of one coordinate: the closed box [c - bs/2, c + bs/2] in fp32 around the centres of the
// blocks i0-1, i0, i0+1 (src/bgkoctomap/bgkoctomap.cpp:497-503; closed intervals: rtree.h:1519-1532).
// Returns a 3-bit mask (bit k: block i0 + k - 1 holds q and is enumerated this scan); rel0 = i0 relative to the grid.
__device__ inline unsigned int axis_candidates(float q, float bs, float half, const GridDesc *g, int a, int &rel0) {
    const long long i0 = axis_index(q, bs);
    rel0 = (int) (i0 - g->base[a]);
    unsigned int mask = 0;
#pragma unroll
    for (int k = -1; k <= 1; ++k) {
        const long long ii = i0 + k;
        const float c = axis_center(ii, bs);
        const float lo = c - half, hi = c + half;
        if (lo > q || q > hi) continue;
        const long long rel = ii - g->base[a];
        if (rel < 0 || rel >= g->n[a] || !g->present[a][rel]) continue;   // block not enumerated this scan
        mask |= 1u << (k + 1);
    }
    return mask;
}

// tile_sums[tile] = memberships of the tile's training entries
__global__ void k_member_count(const float4 *__restrict__ xy, const ScanCounters *__restrict__ c,
                               const DevParams *__restrict__ P, const GridDesc *__restrict__ g,
                               unsigned int *tile_sums) {
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned int n = c->overflow ? 0u : c->n_train;
    const float bs = P->block_size, half = P->half_size;
    unsigned int cnt = 0;
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int i = blockIdx.x * kTile + k * kTileThreads + threadIdx.x;
        if (i < n) {
            const float4 p = xy[i];
            int r;
            const int nx = __popc(axis_candidates(p.x, bs, half, g, 0, r));
            const int ny = __popc(axis_candidates(p.y, bs, half, g, 1, r));
            const int nz = __popc(axis_candidates(p.z, bs, half, g, 2, r));
            cnt += (unsigned int) (nx * ny * nz);
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_cnt;
}

// (dense cell id, entry) pairs in entry order; slots past the membership count get the pad key
__global__ void k_member_fill(const float4 *__restrict__ xy, ScanCounters *c, const DevParams *__restrict__ P,
                              const GridDesc *__restrict__ g, const unsigned int *__restrict__ tile_sums,
                              unsigned int n_tiles, unsigned int *keys, unsigned int *vals, unsigned int cap) {
    __shared__ unsigned int smem[66];
    const unsigned int n = c->overflow ? 0u : c->n_train;
    unsigned int prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->n_members = total;
        if (total > cap) atomicOr(&c->overflow, OVF_MEMBERS);
    }
    if (total > cap) return;
    // pad the tail for the fixed-size sort
    for (unsigned int i = total + blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        keys[i] = kPad;
        vals[i] = 0;
    }
    const float bs = P->block_size, half = P->half_size;
    const unsigned int base = blockIdx.x * kTile + threadIdx.x * kTileItems;
    int rx[kTileItems], ry[kTileItems], rz[kTileItems];
    unsigned int masks[kTileItems];
    unsigned int mine = 0;
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int i = base + k;
        masks[k] = 0;
        rx[k] = ry[k] = rz[k] = 0;
        if (i < n) {
            const float4 p = xy[i];
            const unsigned int mx = axis_candidates(p.x, bs, half, g, 0, rx[k]);
            const unsigned int my = axis_candidates(p.y, bs, half, g, 1, ry[k]);
            const unsigned int mz = axis_candidates(p.z, bs, half, g, 2, rz[k]);
            masks[k] = mx | (my << 3) | (mz << 6);
            mine += (unsigned int) (__popc(mx) * __popc(my) * __popc(mz));
        }
    }
    unsigned int cta_total;
    unsigned int o = prefix + block_exclusive_scan(mine, smem, cta_total);
    const unsigned int n1 = (unsigned int) g->n[1], n2 = (unsigned int) g->n[2];
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int m = masks[k];
        if (!m) continue;
        for (int a = 0; a < 3; ++a) {
            if (!(m & (1u << a))) continue;
            for (int b = 0; b < 3; ++b) {
                if (!(m & (8u << b))) continue;
                for (int d = 0; d < 3; ++d) {
                    if (!(m & (64u << d))) continue;
                    keys[o] = ((unsigned int) (rx[k] + a - 1) * n1 + (unsigned int) (ry[k] + b - 1)) * n2 +
                              (unsigned int) (rz[k] + d - 1);
                    vals[o] = base + k;
                    ++o;
                }
            }
        }
    }
}

// Per sorted membership: the block-sorted training array (coordinates pre-scaled for the method's kernel); per run
// head: the data-block tables, cell -> data block, and one bit for each of the 7 blocks whose ExtendedBlock contains
// this data block (itself and its 6 face neighbours) = the test-block candidates.
__global__ void k_db_place(const unsigned int *__restrict__ keys, const unsigned int *__restrict__ vals,
                           ScanCounters *c, unsigned int cap, const unsigned int *__restrict__ tile_sums,
                           unsigned int n_tiles, const float4 *__restrict__ xy, const DevParams *__restrict__ P,
                           const GridDesc *__restrict__ g, float4 *pts, unsigned int *db_id, unsigned int *db_start,
                           unsigned int *cell_db, unsigned int *test_bits) {
    __shared__ unsigned int smem[66];
    const unsigned int n = c->overflow ? 0u : min(c->n_members, cap);
    unsigned int prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    if (blockIdx.x == 0 && threadIdx.x == 0) { c->n_data_blocks = total; db_start[total] = n; }
    // gather, coalesced over the tile
    const int method = P->method;
    const float ell = P->ell;
    const float gp_scale = (float) (1.73205 / (double) ell);   // gpregressor.h:115 (float(1.73205 / ell))
    for (int k = 0; k < kTileItems; ++k) {
        const unsigned int i = blockIdx.x * kTile + k * kTileThreads + threadIdx.x;
        if (i < n) {
            const float4 p = xy[vals[i]];
            // covSparse's  x / ell  (bgkinference.h:114) | GP's  scale * x  (gpregressor.h:115) hoisted
            pts[i] = method == LA3DM_GP ? make_float4(gp_scale * p.x, gp_scale * p.y, gp_scale * p.z, p.w)
                                        : make_float4(p.x / ell, p.y / ell, p.z / ell, p.w);
        }
    }
    const unsigned int base = blockIdx.x * kTile + threadIdx.x * kTileItems;
    unsigned int flags = 0, cnt = 0;
    unsigned int hk[kTileItems];
    if (base < n) {
        unsigned int prev = base ? keys[base - 1] : 0u;
#pragma unroll
        for (int k = 0; k < kTileItems; ++k) {
            const unsigned int i = base + k;
            hk[k] = 0;
            if (i < n) {
                const unsigned int key = keys[i];
                if (i == 0 || key != prev) { flags |= 1u << k; ++cnt; hk[k] = key; }
                prev = key;
            }
        }
    }
    unsigned int cta_total;
    unsigned int pos = prefix + block_exclusive_scan(cnt, smem, cta_total);
    const int nz = g->n[2], ny = g->n[1], nx = g->n[0];
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
        if (!(flags & (1u << k))) continue;
        const unsigned int id = hk[k];
        db_id[pos] = id;
        db_start[pos] = base + k;
        cell_db[id] = pos + 1;
        ++pos;
        const int z = (int) (id % (unsigned int) nz), y = (int) ((id / (unsigned int) nz) % (unsigned int) ny),
                  x = (int) (id / ((unsigned int) nz * (unsigned int) ny));
        const int dx[7] = {0, 1, -1, 0, 0, 0, 0}, dy[7] = {0, 0, 0, 1, -1, 0, 0}, dz[7] = {0, 0, 0, 0, 0, 1, -1};
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const int xx = x + dx[q], yy = y + dy[q], zz = z + dz[q];
            if (xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz && g->present[0][xx] &&
                g->present[1][yy] && g->present[2][zz]) {
                const unsigned int nid = ((unsigned int) xx * (unsigned int) ny + (unsigned int) yy) *
                                             (unsigned int) nz + (unsigned int) zz;
                atomicOr(&test_bits[nid >> 5], 1u << (nid & 31));
            }
        }
    }
}

// ---- test blocks = set bits of the cell bitmap, in ascending cell order -------------------------------------------
__global__ void k_test_count(const unsigned int *__restrict__ bits, unsigned int n_words_cap,
                             const ScanCounters *__restrict__ c, unsigned int *tile_sums) {
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n_words = c->overflow ? 0u : min((c->n_cells + 31u) >> 5, n_words_cap);
    unsigned int cnt = w < n_words ? __popc(bits[w]) : 0u;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_cnt;
}

// in-place exclusive scan of the tile sums by one CTA (only used when there are many tiles); total at [n_tiles]
__global__ void k_tiles_scan(unsigned int *tile_sums, unsigned int n_tiles) {
    __shared__ unsigned int smem[33];
    unsigned int carry = 0;
    for (unsigned int b = 0; b < n_tiles; b += blockDim.x) {
        const unsigned int j = b + threadIdx.x;
        const unsigned int v = j < n_tiles ? tile_sums[j] : 0u;
        unsigned int total;
        const unsigned int ex = block_exclusive_scan(v, smem, total);
        if (j < n_tiles) tile_sums[j] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void k_test_place(const unsigned int *__restrict__ bits, unsigned int n_words_cap, ScanCounters *c,
                             const unsigned int *__restrict__ tile_sums, unsigned int n_tiles, int prescanned,
                             unsigned int *test_id, unsigned int tests_cap) {
    __shared__ unsigned int smem[66];
    unsigned int prefix, total;
    if (prescanned) { prefix = tile_sums[blockIdx.x]; total = tile_sums[n_tiles]; }
    else block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->n_test_blocks = total;
        if (total > tests_cap) atomicOr(&c->overflow, OVF_TESTS);
    }
    if (total > tests_cap) return;
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n_words = c->overflow ? 0u : min((c->n_cells + 31u) >> 5, n_words_cap);
    unsigned int word = w < n_words ? bits[w] : 0u;
    unsigned int cta_total;
    unsigned int pos = prefix + block_exclusive_scan((unsigned int) __popc(word), smem, cta_total);
    while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        test_id[pos++] = w * 32u + (unsigned int) b;
    }
}

// ---- persistent block map -----------------------------------------------------------------------------------------
// Block key of test block t (dense cell id -> absolute block indices -> BlockHashKey, bgkblock.cpp:73-77)
__device__ inline long long test_block_key(unsigned int id, const GridDesc *g, int &x, int &y, int &z) {
    const int nz = g->n[2], ny = g->n[1];
    z = (int) (id % (unsigned int) nz);
    y = (int) ((id / (unsigned int) nz) % (unsigned int) ny);
    x = (int) (id / ((unsigned int) nz * (unsigned int) ny));
    return make_key(g->base[0] + x, g->base[1] + y, g->base[2] + z);
}

// One thread per test block: look its key up in the map (slot, or 0xFFFFFFFF for a block that does not exist yet) and
// count the new blocks of the tile.  Read-only on the map.
__global__ void __launch_bounds__(kThreads)
k_plan_find(const unsigned int *__restrict__ test_id, const ScanCounters *__restrict__ c, const GridDesc *__restrict__ g,
            const long long *__restrict__ hkeys, const int *__restrict__ hvals, size_t mask, NeighbourPlan *plan,
            unsigned int *tile_sums) {
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_new = false;
    if (!c->overflow && t < c->n_test_blocks) {
        int x, y, z;
        const long long key = test_block_key(test_id[t], g, x, y, z);
        const int slot = hash_find(hkeys, hvals, mask, key);
        is_new = slot < 0;
        plan[t].slot = (unsigned int) slot;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, is_new);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, (unsigned int) __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_cnt;
}

// One thread per test block: create the blocks that do not exist yet, and look up the training ranges of the 7
// neighbours [self,+x,-x,+y,-y,+z,-z] (src/bgkoctomap/bgkblock.cpp:85-101) through the cell -> data block table.
// New blocks get their slots in TEST-BLOCK ORDER (n_blocks + rank among the new ones): the slot numbering -- hence the
// whole pool -- is a function of the scans alone, identical on every replica of a multi-GPU run, which is what lets a
// rank store its results straight into its peers' pools.  For BGKOctoMap the record of a new block is also written
// here (default node everywhere, bgkoctree_node.h:34), by the whole warp, so that every replica holds it whoever
// predicts the block.  This is the first kernel of the scan that writes to the persistent map; every capacity check
// has been made by now.
__global__ void __launch_bounds__(kThreads)
k_plan(const unsigned int *__restrict__ test_id, ScanCounters *c, const ScanArgs *__restrict__ A,
       const GridDesc *__restrict__ g, const unsigned int *__restrict__ cell_db,
       const unsigned int *__restrict__ db_start, long long *hkeys, int *hvals, size_t mask, long long *keys,
       NeighbourPlan *plan, unsigned int *plan_db, unsigned int *heavy_list, unsigned int *light_list,
       uint4 *mega_list, unsigned int *chunk_mega, unsigned char *dirty, unsigned char *touched, unsigned int *cell_test,
       const unsigned int *__restrict__ tile_sums,
       unsigned int n_tiles, int prescanned, unsigned char *pool, const DevParams *__restrict__ P, int init_records) {
    __shared__ unsigned int smem[66];
    if (c->overflow) return;
    unsigned int prefix, total;
    if (prescanned) { prefix = tile_sums[blockIdx.x]; total = tile_sums[n_tiles]; }
    else block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    if (blockIdx.x == 0 && threadIdx.x == 0) c->n_new_blocks = total;
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < c->n_test_blocks;
    NeighbourPlan pl;
    pl.slot = valid ? plan[t].slot : 0u;
    pl.is_new = (valid && pl.slot == 0xFFFFFFFFu) ? 1u : 0u;
    unsigned int cta_total;
    const unsigned int rank = block_exclusive_scan(pl.is_new, smem, cta_total);
    int x = 0, y = 0, z = 0;
    long long key = 0;
    if (valid) key = test_block_key(test_id[t], g, x, y, z);
    if (pl.is_new) {
        pl.slot = A->n_blocks + prefix + rank;                       // < pool_cap: the host keeps room for `tests`
        keys[pl.slot] = key;
        hash_insert(hkeys, hvals, mask, key, (int) pl.slot);
    }
    if (init_records) {
        // Default records of this warp's new blocks, 16 bytes per lane and step.  Multi-GPU: the rank that will predict
        // the block (t % world) writes the record into EVERY replica -- a peer may be ahead of us, and its results must
        // not be overwritten by a default record that we write later; stores of one GPU to one peer arrive in order, so
        // the owner's defaults land before the owner's results.
        const PeerTable *PT = A->peers;
        const int world = (PT && !PT->deferred) ? PT->world : 1, my_rank = PT ? PT->rank : 0;
        const bool mine = !PT || block_owner(key, t, PT->world, true) == PT->rank;
        if (PT && PT->deferred && mine && pl.is_new) dirty[pl.slot] = 1;     // a new block reaches the peers at the next sync
        unsigned int todo = __ballot_sync(0xffffffffu, pl.is_new != 0u && mine);
        const int lane = threadIdx.x & 31;
        const int nodes = P->nodes, st_off = P->st_off, words = P->rec_bytes >> 4;
        const float da = P->def_a, db = P->def_b;
        const unsigned int leaves = (unsigned int) (P->finest & 0xFF);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned int sl = __shfl_sync(0xffffffffu, pl.slot, src);
            const size_t rec_off = (size_t) sl * (size_t) P->rec_bytes;
            for (int w = lane; w < words; w += 32) {
                uint4 v;
                unsigned int *vw = reinterpret_cast<unsigned int *>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int byte0 = 16 * w + 4 * q;
                    unsigned int word;
                    if (byte0 + 4 <= st_off) word = __float_as_uint(((byte0 >> 2) & 1) ? db : da);
                    else {
                        word = 0;
#pragma unroll
                        for (int bb = 0; bb < 4; ++bb) {
                            const int n = byte0 + bb - st_off;
                            const unsigned int by = n < nodes ? (unsigned int) LA3DM_UNKNOWN : (n == nodes ? leaves : 0u);
                            word |= by << (8 * bb);
                        }
                    }
                    vw[q] = word;
                }
                reinterpret_cast<uint4 *>(pool + rec_off)[w] = v;
                for (int p = 0; p < world; ++p)
                    if (p != my_rank) reinterpret_cast<uint4 *>(PT->pool[p] + rec_off)[w] = v;
            }
        }
    }
    // Peers: a block that another rank owns only needs its slot here (every replica numbers the blocks alike); its
    // neighbour ranges, its plan entry and its place in a work list are the owner's business.
    const bool own = valid && (!A->peers || block_owner(key, t, A->shard_world, true) == A->shard_rank);
    unsigned int tot = 0;
    if (valid) touched[pl.slot] = 1;                      // read side: la3dm_export_touched
    if (valid && cell_test) cell_test[test_id[t]] = t + 1;
    if (own) {
        const int nz = g->n[2], ny = g->n[1], nx = g->n[0];
        const int dx[7] = {0, 1, -1, 0, 0, 0, 0}, dy[7] = {0, 0, 0, 1, -1, 0, 0}, dz[7] = {0, 0, 0, 0, 0, 1, -1};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int xx = x + dx[k], yy = y + dy[k], zz = z + dz[k];
            unsigned int start = 0, count = 0;
            if (xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz) {
                const unsigned int nid = ((unsigned int) xx * (unsigned int) ny + (unsigned int) yy) * (unsigned int) nz +
                                         (unsigned int) zz;
                const unsigned int d = cell_db[nid];
                if (d) { start = db_start[d - 1]; count = db_start[d] - start; }
                if (plan_db) plan_db[(size_t) t * 8 + k] = d;      // data block index + 1 (GP: locates the regressor)
            } else if (plan_db) plan_db[(size_t) t * 8 + k] = 0;
            pl.start[k] = start;
            pl.count[k] = count;
            tot += count;
        }
    }
    // this rank's test blocks: the heavy ones (predicted first) and the rest; one atomic per warp and list (every test
    // block of the scan passes here: millions of atomics on one counter serialise in L2)
    {
        const bool listed = own && heavy_list && block_owner(key, t, A->shard_world, A->peers != nullptr) == A->shard_rank;
        const bool mega = listed && tot > A->mega_tot;
        const bool heavy = listed && !mega && tot > A->heavy_tot;
        const bool light = listed && !mega && !heavy && A->shard_world > 1;   // (one rank: walked in cell order)
        const int lane = threadIdx.x & 31;
        const unsigned int lt = (1u << lane) - 1u;
        const unsigned int mh = __ballot_sync(0xffffffffu, heavy), ml = __ballot_sync(0xffffffffu, light);
        unsigned int bh = 0, bl = 0;
        if (lane == 0) {
            if (mh) bh = atomicAdd(&c->n_heavy, (unsigned int) __popc(mh));
            if (ml) bl = atomicAdd(&c->n_light, (unsigned int) __popc(ml));
        }
        bh = __shfl_sync(0xffffffffu, bh, 0);
        bl = __shfl_sync(0xffffffffu, bl, 0);
        if (heavy) heavy_list[bh + __popc(mh & lt)] = t;
        if (light) light_list[bl + __popc(ml & lt)] = t;
        if (mega) {                        // cut into chunks, each predicted as a unit of its own (predict_bgk.cu)
            const unsigned int nch = (tot + A->mega_chunk - 1u) / A->mega_chunk;
            const unsigned int first = atomicAdd(&c->n_mega_chunks, nch), m = atomicAdd(&c->n_mega, 1u);
            mega_list[m] = make_uint4(t, first, nch, 0u);
            for (unsigned int q = 0; q < nch; ++q) chunk_mega[first + q] = m;
        }
    }
    if (!valid) return;
    uint4 *dst = reinterpret_cast<uint4 *>(plan + t);
    const uint4 *src = reinterpret_cast<const uint4 *>(&pl);
    if (own) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; }
    dst[3] = src[3];                       // (count[5..6], slot, is_new)
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

inline int bits_for(unsigned int n) {
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long) n) ++b;
    return b;
}

}  // namespace

// capacity for `blocks` blocks in the pool and a hash table at <= 50 % load (host side, between scans)
void Map::ensure_pool(size_t blocks, bool exact) {
    if (blocks > pool_cap) {
        // grow-only (a move copies the pool and invalidates the captured graph); first allocation ~1 GB; doubling while
        // the pool is small, +25 % once it holds tens of GB (the copy needs old and new side by side)
        const size_t first = ((size_t) 1 << 30) / (size_t) hp.rec_bytes;
        const size_t big = ((size_t) 16 << 30) / (size_t) hp.rec_bytes;
        const size_t want = exact ? blocks : std::max((blocks > big ? blocks + blocks / 4 : 2 * blocks) + 1024, first);
        keys.grow_keep(want * sizeof(long long), stream);
        pool.grow_keep(want * (size_t) hp.rec_bytes, stream);
        {
            const size_t old = dirty.cap;
            dirty.grow_keep(want, stream);
            if (dirty.cap > old) LA3DM_CUDA(cudaMemsetAsync(dirty.as<unsigned char>() + old, 0, dirty.cap - old, stream));
        }
        {
            const size_t old = touched.cap;
            touched.grow_keep(want, stream);
            if (touched.cap > old) LA3DM_CUDA(cudaMemsetAsync(touched.as<unsigned char>() + old, 0, touched.cap - old, stream));
        }
        pool_cap = want;
        invalidate_graph();
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
        invalidate_graph();
    }
}

// key -> slot table from scratch over keys[0 .. n_blocks) (after an import)
void Map::rebuild_hash() {
    k_hash_clear<<<ceil_div((long long) hash_cap, kThreads), kThreads, 0, stream>>>(hkeys.as<long long>(), hash_cap);
    if (n_blocks > 0)
        k_hash_rebuild<<<ceil_div(n_blocks, kThreads), kThreads, 0, stream>>>(
            keys.as<long long>(), (unsigned int) n_blocks, hkeys.as<long long>(), hvals.as<int>(), hash_cap - 1);
}

// the float-stepped block grid of the scan from the bounding box of the training set (d_mm + 12)
void Map::enqueue_block_grid() {
    LA3DM_CUDA(cudaMemsetAsync(d_grid, 0, sizeof(GridDesc), stream));
    k_grid<<<1, kGridThreads, 0, stream>>>(d_mm + 12, d_params, d_grid, d_cnt, caps.cells);
    ++launches;
}

// Input: xy[0..n_train) and the counters on the device.
// Output: pts_sorted, db_id/db_start, test_id, plan[] (slot, is_new, 7 ranges); counters on the device.
void Map::enqueue_binning() {
    const float4 *d_xy = xy.as<float4>();
    unsigned int *tile_sums = tiles.as<unsigned int>();

    // the float-stepped block grid; dense per-cell tables start empty
    // (the bounding box mm was accumulated by the front-end kernels that wrote xy)
    enqueue_block_grid();
    LA3DM_CUDA(cudaMemsetAsync(cell_db.p, 0, (size_t) caps.cells * 4, stream));
    if (hp.method == LA3DM_GP) LA3DM_CUDA(cudaMemsetAsync(cell_test.p, 0, (size_t) caps.cells * 4, stream));
    const unsigned int n_words = (caps.cells + 31) / 32;
    LA3DM_CUDA(cudaMemsetAsync(test_bits.p, 0, (size_t) n_words * 4, stream));

    // memberships -> sort by cell
    const int t_tiles = ceil_div(caps.train, kTile);
    k_member_count<<<t_tiles, kTileThreads, 0, stream>>>(d_xy, d_cnt, d_params, d_grid, tile_sums);
    cub::DoubleBuffer<unsigned int> dk(sort_keys[0].as<unsigned int>(), sort_keys[1].as<unsigned int>());
    cub::DoubleBuffer<unsigned int> dv(sort_vals[0].as<unsigned int>(), sort_vals[1].as<unsigned int>());
    k_member_fill<<<t_tiles, kTileThreads, 0, stream>>>(d_xy, d_cnt, d_params, d_grid, tile_sums,
                                                        (unsigned int) t_tiles, dk.Current(), dv.Current(),
                                                        caps.members);
    size_t tmp = cub_tmp_bytes;
    const int end_bit = bits_for(caps.cells);
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, dk, dv, (int) caps.members, 0, end_bit, stream));
    const int m_tiles = ceil_div(caps.members, kTile);
    k_run_count<<<m_tiles, kTileThreads, 0, stream>>>(dk.Current(), &d_cnt->n_members, caps.members, tile_sums, d_cnt);
    k_db_place<<<m_tiles, kTileThreads, 0, stream>>>(dk.Current(), dv.Current(), d_cnt, caps.members, tile_sums,
                                                     (unsigned int) m_tiles, d_xy, d_params, d_grid,
                                                     pts_sorted.as<float4>(), db_id.as<unsigned int>(),
                                                     db_start.as<unsigned int>(), cell_db.as<unsigned int>(),
                                                     test_bits.as<unsigned int>());
    launches += 7 + 2 + (end_bit + 7) / 8;

    // BGKL: per-block training lists (hits + each ray once); the neighbour plan then indexes those
    if (hp.method == LA3DM_BGKL) enqueue_bgkl_lists(dk.Current(), dv.Current());
    // GP: sizes of the per-data-block regressors -- every capacity check must come before k_plan touches the map
    if (hp.method == LA3DM_GP) enqueue_gp_sizes();

    // test blocks, their slots in the map and their neighbour plans
    const int w_tiles = ceil_div(n_words, kThreads);
    const int prescanned = w_tiles > 1024 ? 1 : 0;
    k_test_count<<<w_tiles, kThreads, 0, stream>>>(test_bits.as<unsigned int>(), n_words, d_cnt, tile_sums);
    if (prescanned) { k_tiles_scan<<<1, 1024, 0, stream>>>(tile_sums, (unsigned int) w_tiles); ++launches; }
    k_test_place<<<w_tiles, kThreads, 0, stream>>>(test_bits.as<unsigned int>(), n_words, d_cnt, tile_sums,
                                                   (unsigned int) w_tiles, prescanned, test_id.as<unsigned int>(),
                                                   caps.tests);
    // slots of the test blocks: lookup + count of the new ones, then creation in test-block order
    const int p_tiles = ceil_div(caps.tests, kThreads);
    const int p_prescanned = p_tiles > 1024 ? 1 : 0;
    k_plan_find<<<p_tiles, kThreads, 0, stream>>>(test_id.as<unsigned int>(), d_cnt, d_grid, hkeys.as<long long>(),
                                                  hvals.as<int>(), hash_cap - 1, plan.as<NeighbourPlan>(), tile_sums);
    if (p_prescanned) { k_tiles_scan<<<1, 1024, 0, stream>>>(tile_sums, (unsigned int) p_tiles); ++launches; }
    k_plan<<<p_tiles, kThreads, 0, stream>>>(
        test_id.as<unsigned int>(), d_cnt, d_args, d_grid, cell_db.as<unsigned int>(),
        hp.method == LA3DM_BGKL ? seg_start.as<unsigned int>() : db_start.as<unsigned int>(),
        hkeys.as<long long>(), hvals.as<int>(), hash_cap - 1, keys.as<long long>(), plan.as<NeighbourPlan>(),
        hp.method == LA3DM_GP ? plan_db.as<unsigned int>() : nullptr,
        hp.method == LA3DM_BGK ? heavy_list.as<unsigned int>() : nullptr, light_list.as<unsigned int>(),
        mega_list.as<uint4>(), chunk_mega.as<unsigned int>(), dirty.as<unsigned char>(), touched.as<unsigned char>(),
        hp.method == LA3DM_GP ? cell_test.as<unsigned int>() : nullptr, tile_sums, (unsigned int) p_tiles,
        p_prescanned, pool.as<unsigned char>(), d_params, hp.method == LA3DM_BGK ? 1 : 0);
    ++launches;
    launches += 3;
}

}  // namespace la3dm_b200
