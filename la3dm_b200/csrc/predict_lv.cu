// la3dm_b200 -- BGKLVOctoMap: per-voxel training sets, prediction, node update and pruning.
//
// Replaces the single fused loop of BGKLVOctoMap::insert_pointcloud (src/bgklvoctomap/bgklvoctomap.cpp:133-273):
//   every block of the scan's bbox grid is created (:140-148); for every leaf AT BASE RESOLUTION (:157-160) the
//   training entries inside the closed box centre +- ell are collected with an R-tree query (:162-171): hits as points,
//   each ray that has a marker in the box once (:179-207); BGKLVInference::predict for the one voxel
//   (include/bgklvoctomap/bgklvinference.h:100-157: segment distance / ell clamped to 1, sparse kernel, no negative
//   clamp); node.update if kbar > 0.001f (:236-238, src/bgklvoctomap/bgklvoctree_node.cpp:29-77); blocks that had data
//   are pruned if original_size (:266-273).
//
// Here: the R-tree is a dense uniform grid of cell size ell over the training entries (sort by cell), a voxel's query
// visits the <= 4^3 cells its box overlaps and applies the reference's closed-box test; a ray is taken at its first
// marker inside the box (markers of a ray are collinear, so the ones inside a box are contiguous).  Voxels that have
// data are compacted into a list and then predicted by one warp each.
#include <cub/cub.cuh>

#include "engine.cuh"
#include "runs.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kThreads = 256;
constexpr unsigned int kPad = 0xFFFFFFFFu;
constexpr int kLvPruned = LA3DM_LV_PRUNED;

// uniform grid over the training entries
struct QGrid {
    double x0, y0, z0, inv;
    int n[3];
    unsigned int cells;
};

__device__ __forceinline__ int qcell(double x, double x0, double inv, int n) {
    const int c = (int) floor((x - x0) * inv);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void k_lv_qgrid(const unsigned int *__restrict__ mm, const DevParams *__restrict__ P, QGrid *g,
                           ScanCounters *c, unsigned int cells_cap) {
    if (c->overflow || c->n_train == 0) { g->cells = 0; g->n[0] = g->n[1] = g->n[2] = 0; return; }
    const double ell = (double) P->ell;
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = (double) float_unflip(mm[a]); hi[a] = (double) float_unflip(mm[3 + a]); }
    g->inv = 1.0 / ell;
    g->x0 = lo[0] - 1.5 * ell; g->y0 = lo[1] - 1.5 * ell; g->z0 = lo[2] - 1.5 * ell;
    unsigned long long cells = 1;
    for (int a = 0; a < 3; ++a) {
        const double span = (hi[a] - lo[a] + 3.0 * ell) * g->inv;
        const long long n = (long long) span + 2;
        g->n[a] = (int) (n > 0x7FFFFFF ? 0x7FFFFFF : n);
        cells *= (unsigned long long) g->n[a];
    }
    if (cells >= 0x7FFFFFF0ull) { atomicOr(&c->overflow, OVF_EXTENT); g->cells = 0; return; }
    g->cells = (unsigned int) cells;
    atomicMax(&c->n_cells, (unsigned int) cells);
    if (cells > (unsigned long long) cells_cap) atomicOr(&c->overflow, OVF_CELLS);
}

__device__ __forceinline__ unsigned int qcell_id(const QGrid &g, float x, float y, float z) {
    const int cx = qcell((double) x, g.x0, g.inv, g.n[0]), cy = qcell((double) y, g.y0, g.inv, g.n[1]),
              cz = qcell((double) z, g.z0, g.inv, g.n[2]);
    return ((unsigned int) cx * (unsigned int) g.n[1] + (unsigned int) cy) * (unsigned int) g.n[2] + (unsigned int) cz;
}

__global__ void k_lv_cellkeys(const float4 *__restrict__ xy, ScanCounters *c, const QGrid *__restrict__ g,
                              unsigned int *keys, unsigned int *vals, unsigned int cap) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned int n = c->overflow ? 0u : c->n_train;
    if (i == 0 && !c->overflow) {
        c->n_members = n;
        if (n > cap) atomicOr(&c->overflow, OVF_MEMBERS);
    }
    if (i >= n || n > cap) { keys[i] = kPad; vals[i] = 0; return; }
    const float4 p = xy[i];
    keys[i] = qcell_id(*g, p.x, p.y, p.z);
    vals[i] = i;
}

// per run head: cell -> run index + 1, run -> first sorted position (+ sentinel)
__global__ void k_lv_cell_place(const unsigned int *__restrict__ keys, ScanCounters *c, unsigned int cap,
                                const unsigned int *__restrict__ tile_sums, unsigned int n_tiles,
                                unsigned int *run_first, unsigned int *cell_run) {
    __shared__ unsigned int smem[66];
    const unsigned int n = c->overflow ? 0u : min(c->n_members, cap);
    unsigned int prefix, total;
    block_tile_prefix(tile_sums, blockIdx.x, n_tiles, smem, prefix, total);
    if (blockIdx.x == 0 && threadIdx.x == 0) { c->n_data_blocks = total; run_first[total] = n; }
    const unsigned int base = blockIdx.x * kTile + threadIdx.x * kTileItems;
    unsigned int flags = 0, cnt = 0;
    if (base < n) {
        unsigned int prev = base ? keys[base - 1] : 0u;
#pragma unroll
        for (int k = 0; k < kTileItems; ++k) {
            const unsigned int i = base + k;
            if (i < n) {
                const unsigned int key = keys[i];
                if (i == 0 || key != prev) { flags |= 1u << k; ++cnt; }
                prev = key;
            }
        }
    }
    unsigned int cta_total;
    unsigned int pos = prefix + block_exclusive_scan(cnt, smem, cta_total);
#pragma unroll
    for (int k = 0; k < kTileItems; ++k)
        if (flags & (1u << k)) {
            run_first[pos] = base + k;
            cell_run[keys[base + k]] = pos + 1;
            ++pos;
        }
}

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

struct BlockCell {
    int x, y, z;
    bool present;
    long long key;
};

__device__ __forceinline__ BlockCell block_cell(const GridDesc *g, unsigned int id) {
    BlockCell b;
    const unsigned int nz = (unsigned int) g->n[2], ny = (unsigned int) g->n[1];
    b.z = (int) (id % nz); b.y = (int) ((id / nz) % ny); b.x = (int) (id / (nz * ny));
    b.present = g->present[0][b.x] && g->present[1][b.y] && g->present[2][b.z];
    b.key = make_key(g->base[0] + b.x, g->base[1] + b.y, g->base[2] + b.z);
    return b;
}

// closed-box test of the R-tree query (rtree.h:1519-1532) for a point entry
__device__ __forceinline__ bool in_box(const float4 p, const float lo[3], const float hi[3]) {
    return !(p.x < lo[0] || p.x > hi[0] || p.y < lo[1] || p.y > hi[1] || p.z < lo[2] || p.z > hi[2]);
}

struct VoxelQuery {
    float p[3], lo[3], hi[3];
    int c0[3], c1[3];
};

// voxel centre (Block::get_loc), its query box centre -+ ell in fp32 (:161-163) and the grid cells the box overlaps
__device__ __forceinline__ VoxelQuery voxel_query(long long key, int node, const float3 *__restrict__ lut,
                                                  const DevParams &P, const QGrid &q) {
    VoxelQuery v;
    const float3 off = lut[node];
    v.p[0] = off.x + axis_center(key >> 40, P.block_size);
    v.p[1] = off.y + axis_center((key >> 20) & 0xFFFFF, P.block_size);
    v.p[2] = off.z + axis_center(key & 0xFFFFF, P.block_size);
    const double g0[3] = {q.x0, q.y0, q.z0};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        v.lo[a] = v.p[a] - P.ell;
        v.hi[a] = v.p[a] + P.ell;
        v.c0[a] = qcell((double) v.lo[a], g0[a], q.inv, q.n[a]);
        v.c1[a] = qcell((double) v.hi[a], g0[a], q.inv, q.n[a]);
    }
    return v;
}

// Active voxels: base-resolution leaves of the grid's blocks whose query box holds at least one training entry
// (has_gp_points_in_bbox, :165-166).  Read-only with respect to the map: a block that does not exist yet has all its
// finest nodes as leaves.
__global__ void k_lv_active(const GridDesc *__restrict__ g, const QGrid *__restrict__ qg, ScanCounters *c,
                            const DevParams *__restrict__ P, const float3 *__restrict__ lut,
                            const long long *__restrict__ hkeys, const int *__restrict__ hvals, size_t mask,
                            const unsigned char *__restrict__ pool, const float4 *__restrict__ xy,
                            const unsigned int *__restrict__ vals, const unsigned int *__restrict__ run_first,
                            const unsigned int *__restrict__ cell_run, uint2 *active, unsigned int active_cap,
                            unsigned int tests_cap) {
    if (c->overflow) return;
    const unsigned int n_bc = g->n_cells;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_bc > tests_cap) {
        c->n_test_blocks = n_bc;
        atomicOr(&c->overflow, OVF_TESTS);
    }
    if (n_bc > tests_cap) return;
    const QGrid q = *qg;
    const int finest = P->finest, f_off = P->layer_off[P->depth - 1];
    const unsigned long long total = (unsigned long long) n_bc * (unsigned long long) finest;
    for (unsigned long long w = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned int bc = (unsigned int) (w / (unsigned int) finest);
        const int j = (int) (w % (unsigned int) finest);
        const BlockCell b = block_cell(g, bc);
        if (!b.present) continue;
        const int slot = hash_find(hkeys, hvals, mask, b.key);
        if (slot >= 0) {   // only leaves at base resolution (:157-160): a pruned finest node is not a leaf
            const unsigned char st = pool[(size_t) slot * P->rec_bytes + P->st_off + f_off + j];
            if ((st & 7) == kLvPruned) continue;
        }
        const VoxelQuery v = voxel_query(b.key, f_off + j, lut, *P, q);
        bool found = false;
        for (int cx = v.c0[0]; cx <= v.c1[0] && !found; ++cx)
            for (int cy = v.c0[1]; cy <= v.c1[1] && !found; ++cy)
                for (int cz = v.c0[2]; cz <= v.c1[2] && !found; ++cz) {
                    const unsigned int cid = ((unsigned int) cx * (unsigned int) q.n[1] + (unsigned int) cy) *
                                                 (unsigned int) q.n[2] + (unsigned int) cz;
                    const unsigned int r = cell_run[cid];
                    if (!r) continue;
                    for (unsigned int i = run_first[r - 1]; i < run_first[r] && !found; ++i)
                        found = in_box(xy[vals[i]], v.lo, v.hi);
                }
        if (found) {
            const unsigned int pos = atomicAdd(&c->lv_active, 1u);
            if (pos < active_cap) active[pos] = make_uint2(bc, (unsigned int) j);
        }
    }
}

// Every block of the bbox grid exists after the scan (:140-148).  First kernel that touches the persistent map.
__global__ void k_lv_blocks(const GridDesc *__restrict__ g, ScanCounters *c, const ScanArgs *__restrict__ A,
                            long long *hkeys, int *hvals, size_t mask, long long *keys, unsigned char *touched, unsigned int *blk_slot,
                            unsigned char *blk_flags, unsigned int active_cap) {
    if (c->overflow) return;
    if (c->lv_active > active_cap) {      // checked here, before anything is written
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&c->overflow, OVF_LVACTIVE);
        return;
    }
    const unsigned int bc = blockIdx.x * blockDim.x + threadIdx.x;
    if (bc >= g->n_cells) return;
    const BlockCell b = block_cell(g, bc);
    blk_flags[bc] = 0;
    if (!b.present) { blk_slot[bc] = 0xFFFFFFFFu; return; }
    int slot = hash_find(hkeys, hvals, mask, b.key);
    if (slot < 0) {
        slot = (int) (A->n_blocks + atomicAdd(&c->n_new_blocks, 1u));
        keys[slot] = b.key;
        hash_insert(hkeys, hvals, mask, b.key, slot);
        blk_flags[bc] = 1;                // new: record to be initialised
    }
    blk_slot[bc] = (unsigned int) slot;
    touched[slot] = 1;                         // read side: la3dm_export_touched
}

// default nodes for the new blocks: (prior_A, prior_B), UNKNOWN, !classified  (bgklvoctree_node.h:35)
__global__ void k_lv_init_new(const GridDesc *__restrict__ g, const ScanCounters *__restrict__ c,
                              const DevParams *__restrict__ P, const unsigned int *__restrict__ blk_slot,
                              const unsigned char *__restrict__ blk_flags, unsigned char *pool, unsigned int active_cap) {
    if (c->overflow || c->lv_active > active_cap) return;
    const unsigned int n_bc = g->n_cells;
    const unsigned int words = (unsigned int) P->rec_bytes >> 2;          // 4-byte words per record
    const int nodes = P->nodes, st_off = P->st_off;
    const unsigned long long total = (unsigned long long) n_bc * words;
    for (unsigned long long w = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned int bc = (unsigned int) (w / words), k = (unsigned int) (w % words);
        if (!(blk_flags[bc] & 1)) continue;
        unsigned int v;
        const int byte = (int) k * 4;
        if (byte < st_off) v = __float_as_uint((k & 1) ? P->def_b : P->def_a);
        else {
            v = 0;
            for (int q = 0; q < 4; ++q)
                if (byte + q - st_off < nodes) v |= (unsigned int) LA3DM_UNKNOWN << (8 * q);
        }
        reinterpret_cast<unsigned int *>(pool + (size_t) blk_slot[bc] * P->rec_bytes)[k] = v;
    }
}

// covSparseLine element (bgklvinference.h:143-157): d / ell clamped to 1, sparse kernel, NO clamp of negative values
__device__ __forceinline__ float lv_kernel(float d, float ell, float sf2) {
    float x = d / ell;
    if (x > 1.0) x = 1.0f;
    const float t = x * 2.0f * 3.1415926f;
    float s, co;
    sincosf(t, &s, &co);
    return (((2.0f + co) * (1.0f - x) / 3.0f) + s / (2.0f * 3.1415926f)) * sf2;
}

// point3f::norm() narrowed to float (d(i,j) is a float matrix)
__device__ __forceinline__ float normf(float dx, float dy, float dz) {
    return (float) sqrt((double) (dx * dx + dy * dy + dz * dz));
}

// point_to_line_dist (bgklvinference.h:100-137) for one voxel centre and one segment
__device__ __forceinline__ float lv_seg_dist(const float *p, const float4 a, const float4 b) {
    const float vx = b.x - a.x, vy = b.y - a.y, vz = b.z - a.z;
    const float c2 = vx * vx + vy * vy + vz * vz;
    const float len = (float) sqrt((double) c2);
    const float px = p[0] - a.x, py = p[1] - a.y, pz = p[2] - a.z;
    if (len < 0.0001f) return normf(px, py, pz);
    const float c1 = px * vx + py * vy + pz * vz;
    if (c1 <= 0) return normf(px, py, pz);
    if (c2 <= c1) return normf(p[0] - b.x, p[1] - b.y, p[2] - b.z);
    const float t = c1 / c2;              // the double quotient narrowed to float equals the float quotient
    const float nx = a.x + vx * t, ny = a.y + vy * t, nz = a.z + vz * t;
    return normf(p[0] - nx, p[1] - ny, p[2] - nz);
}

// Occupancy::get_prob / get_var / update (src/bgklvoctomap/bgklvoctree_node.cpp:29-77)
__device__ __forceinline__ float lv_prob(float a, float b, float min_W) {
    const float W = (a + b < min_W) ? min_W : a + b;
    if (a > b) return (float) ((double) (a / (W - b)) + (double) (W - a - b) * 0.5 / (double) (W - b));
    return (float) (0.5 * (double) (W - b - a) / (double) (W - a));
}

__device__ __forceinline__ float lv_var(float a, float b, float min_W) {
    const float prob = lv_prob(a, b, min_W);
    const float W = (a + b < min_W) ? min_W : a + b;
    return (float) ((double) (a / W) * pow((double) (1 - prob), 2.0) + (double) ((W - a - b) / W) * pow(0.5 - (double) prob, 2.0) +
                    (double) (b / W) * pow((double) prob, 2.0));
}

__device__ __forceinline__ unsigned char lv_classify(float a, float b, const DevParams &P) {
    if (lv_var(a, b, P.min_W) > P.var_thresh) return LA3DM_LV_UNCERTAIN;
    const float p = lv_prob(a, b, P.min_W);
    return p > P.occupied_thresh ? LA3DM_OCCUPIED : (p < P.free_thresh ? LA3DM_FREE : LA3DM_UNKNOWN);
}

// one warp per active voxel
__global__ void __launch_bounds__(kThreads)
k_lv_predict(const uint2 *__restrict__ active, const GridDesc *__restrict__ g, const QGrid *__restrict__ qg,
             ScanCounters *c, const DevParams *__restrict__ Pg, const float3 *__restrict__ lut,
             const unsigned int *__restrict__ blk_slot, unsigned char *blk_flags, unsigned char *pool,
             const float4 *__restrict__ xy, const int *__restrict__ ray_of, const float4 *__restrict__ rays,
             const unsigned int *__restrict__ ray_first, const unsigned int *__restrict__ vals,
             const unsigned int *__restrict__ run_first, const unsigned int *__restrict__ cell_run,
             unsigned int active_cap) {
    if (c->overflow) return;
    const DevParams &P = *Pg;
    const QGrid q = *qg;
    const unsigned int n_act = min(c->lv_active, active_cap);
    const int lane = threadIdx.x & 31;
    const unsigned int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_w = (gridDim.x * blockDim.x) >> 5;
    const int f_off = P.layer_off[P.depth - 1];
    unsigned long long updates = 0, pairs = 0;
    for (unsigned int w = gw; w < n_act; w += n_w) {
        const uint2 av = active[w];
        const BlockCell b = block_cell(g, av.x);
        const int node = f_off + (int) av.y;
        const VoxelQuery v = voxel_query(b.key, node, lut, P, q);
        float yb = 0.f, kb = 0.f;
        unsigned int cnt = 0;
        // The query box covers 3 x 3 x 3 cells of edge ell with a handful of entries each: a lane looks up one cell,
        // a warp scan concatenates the cells' entry ranges, and the warp then walks the concatenation 32 entries at a
        // time (striding over one cell at a time would leave most lanes idle).
        const int ncy = v.c1[1] - v.c0[1] + 1, ncz = v.c1[2] - v.c0[2] + 1;
        const int ncell = (v.c1[0] - v.c0[0] + 1) * ncy * ncz;
        for (int cbase = 0; cbase < ncell; cbase += 32) {
            unsigned int c_start = 0, c_cnt = 0;
            const int ci = cbase + lane;
            if (ci < ncell) {
                const int cx = v.c0[0] + ci / (ncy * ncz), cy = v.c0[1] + (ci / ncz) % ncy, cz = v.c0[2] + ci % ncz;
                const unsigned int cid = ((unsigned int) cx * (unsigned int) q.n[1] + (unsigned int) cy) *
                                             (unsigned int) q.n[2] + (unsigned int) cz;
                const unsigned int r = cell_run[cid];
                if (r) { c_start = run_first[r - 1]; c_cnt = run_first[r] - c_start; }
            }
            unsigned int c_inc = c_cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, c_inc, d);
                if (lane >= d) c_inc += up;
            }
            const unsigned int n_ent = __shfl_sync(0xffffffffu, c_inc, 31);
            for (unsigned int e0 = 0; e0 < n_ent; e0 += 32) {
                const unsigned int ei = e0 + (unsigned int) lane;
                int own = 0;                                   // first cell whose inclusive count exceeds ei
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const unsigned int below = __shfl_sync(0xffffffffu, c_inc, own + step - 1);
                    if (below <= ei) own += step;
                }
                own = min(own, 31);
                const unsigned int o_inc = __shfl_sync(0xffffffffu, c_inc, own), o_cnt = __shfl_sync(0xffffffffu, c_cnt, own),
                                   o_start = __shfl_sync(0xffffffffu, c_start, own);
                {
                    if (ei < n_ent) {
                        const unsigned int i = o_start + (ei - (o_inc - o_cnt));
                        const unsigned int e = vals[i];
                        const float4 pt = xy[e];
                        if (!in_box(pt, v.lo, v.hi)) continue;
                        const int rid = ray_of[e];
                        float k;
                        if (rid < 0) {          // a hit: degenerate segment, label 1 (:180-188)
                            k = lv_kernel(normf(v.p[0] - pt.x, v.p[1] - pt.y, v.p[2] - pt.z), P.ell, P.sf2);
                            yb += k * 1.0f;
                        } else {                // a marker: its ray once per voxel (:189-199) -- at its first marker in the box
                            const unsigned int e0 = ray_first[rid];
                            if (e != e0) {
                                if (in_box(xy[e0], v.lo, v.hi)) continue;
                                if (e - 1 != e0 && in_box(xy[e - 1], v.lo, v.hi)) continue;
                            }
                            k = lv_kernel(lv_seg_dist(v.p, rays[2 * (size_t) rid], rays[2 * (size_t) rid + 1]), P.ell, P.sf2);
                            yb += k * 0.0f;
                        }
                        kb += k;
                        ++cnt;
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            yb += __shfl_xor_sync(0xffffffffu, yb, o);
            kb += __shfl_xor_sync(0xffffffffu, kb, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) {
            pairs += cnt;
            blk_flags[av.x] |= 2;            // Block_has_info (:240); one byte per block, benign same-value races
            if (kb > 0.001f) {               // :236-238
                unsigned char *rec = pool + (size_t) blk_slot[av.x] * P.rec_bytes;
                float2 *ab = reinterpret_cast<float2 *>(rec) + node;
                float2 val = *ab;
                val.x += yb;
                val.y += kb - yb;
                *ab = val;
                rec[P.st_off + node] = lv_classify(val.x, val.y, P) | 0x80;
                ++updates;
            }
        }
    }
    if (lane == 0 && (updates | pairs)) {
        atomicAdd(&c->updates, updates);
        atomicAdd(&c->pairs, pairs);
    }
}

// OcTree::prune (src/bgklvoctomap/bgklvoctree.cpp:101-148) for the blocks that had data, one CTA per block
__global__ void __launch_bounds__(kThreads)
k_lv_prune(const GridDesc *__restrict__ g, ScanCounters *c, const DevParams *__restrict__ Pg,
           const unsigned int *__restrict__ blk_slot, const unsigned char *__restrict__ blk_flags, unsigned char *pool) {
    if (c->overflow) return;
    const DevParams &P = *Pg;
    for (unsigned int bc = blockIdx.x; bc < g->n_cells; bc += gridDim.x) {
        if (!(blk_flags[bc] & 2)) continue;
        if (threadIdx.x == 0) atomicAdd(&c->n_test_blocks, 1u);
        if (!P.original_size) continue;
        unsigned char *rec = pool + (size_t) blk_slot[bc] * P.rec_bytes;
        float2 *ab = reinterpret_cast<float2 *>(rec);
        unsigned char *st = rec + P.st_off;
        for (int d = P.depth - 1; d > 0; --d) {
            const int off = P.layer_off[d], poff = P.layer_off[d - 1];
            const int groups = 1 << (3 * (d - 1));
            for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
                const unsigned char s0 = st[off + 8 * gi] & 7;
                if (s0 != LA3DM_UNKNOWN && s0 != kLvPruned) {        // FREE, OCCUPIED and UNCERTAIN groups collapse (:114-126)
                    bool same = true;
                    for (int i = 1; i < 8; ++i) same = same && ((st[off + 8 * gi + i] & 7) == s0);
                    if (same) {
                        ab[poff + gi] = ab[off + 8 * gi];
                        st[poff + gi] = (st[poff + gi] & 0x80) | s0;
                        for (int i = 0; i < 8; ++i) st[off + 8 * gi + i] = (st[off + 8 * gi + i] & 0x80) | kLvPruned;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void k_lv_finish(ScanCounters *c) {
    if (c->overflow) return;
    c->visits = c->lv_active;
}

inline int bits_for(unsigned int n) {
    int b = 1;
    while (b < 32 && (1ull << b) < (unsigned long long) n) ++b;
    return b;
}

}  // namespace

void Map::enqueue_lv() {
    const float4 *d_xy = xy.as<float4>();
    unsigned int *mm = d_mm + 12;
    unsigned int *tile_sums = tiles.as<unsigned int>();
    QGrid *qg = lv_qgrid.as<QGrid>();
    // block grid of the scan (get_blocks_in_bbox) and the uniform grid over the training entries
    enqueue_block_grid();
    LA3DM_CUDA(cudaMemsetAsync(cell_db.p, 0, (size_t) caps.cells * 4, stream));
    k_lv_qgrid<<<1, 1, 0, stream>>>(mm, d_params, qg, d_cnt, caps.cells);
    cub::DoubleBuffer<unsigned int> dk(sort_keys[0].as<unsigned int>(), sort_keys[1].as<unsigned int>());
    cub::DoubleBuffer<unsigned int> dv(sort_vals[0].as<unsigned int>(), sort_vals[1].as<unsigned int>());
    k_lv_cellkeys<<<ceil_div(caps.members, kThreads), kThreads, 0, stream>>>(d_xy, d_cnt, qg, dk.Current(), dv.Current(),
                                                                            caps.members);
    size_t tmp = cub_tmp_bytes;
    const int end_bit = bits_for(caps.cells);
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, dk, dv, (int) caps.members, 0, end_bit, stream));
    const int m_tiles = ceil_div(caps.members, kTile);
    k_run_count<<<m_tiles, kTileThreads, 0, stream>>>(dk.Current(), &d_cnt->n_members, caps.members, tile_sums, d_cnt);
    k_lv_cell_place<<<m_tiles, kTileThreads, 0, stream>>>(dk.Current(), d_cnt, caps.members, tile_sums,
                                                          (unsigned int) m_tiles, db_start.as<unsigned int>(),
                                                          cell_db.as<unsigned int>());
    // voxels that have data (read-only), then -- after every capacity check -- the blocks, then the prediction
    const int wide = num_sms * 8;
    k_lv_active<<<wide, kThreads, 0, stream>>>(d_grid, qg, d_cnt, d_params, d_lut, hkeys.as<long long>(),
                                               hvals.as<int>(), hash_cap - 1, pool.as<unsigned char>(), d_xy,
                                               dv.Current(), db_start.as<unsigned int>(), cell_db.as<unsigned int>(),
                                               lv_active.as<uint2>(), caps.lv_active, caps.tests);
    k_lv_blocks<<<ceil_div(caps.tests, kThreads), kThreads, 0, stream>>>(
        d_grid, d_cnt, d_args, hkeys.as<long long>(), hvals.as<int>(), hash_cap - 1, keys.as<long long>(),
        touched.as<unsigned char>(), lv_blk_slot.as<unsigned int>(), lv_blk_flags.as<unsigned char>(), caps.lv_active);
    k_lv_init_new<<<wide, kThreads, 0, stream>>>(d_grid, d_cnt, d_params, lv_blk_slot.as<unsigned int>(),
                                                 lv_blk_flags.as<unsigned char>(), pool.as<unsigned char>(),
                                                 caps.lv_active);
    record_event(ev_p0);
    k_lv_predict<<<wide, kThreads, 0, stream>>>(lv_active.as<uint2>(), d_grid, qg, d_cnt, d_params, d_lut,
                                                lv_blk_slot.as<unsigned int>(), lv_blk_flags.as<unsigned char>(),
                                                pool.as<unsigned char>(), d_xy, ray_of.as<int>(), rays.as<float4>(),
                                                ray_first.as<unsigned int>(), dv.Current(),
                                                db_start.as<unsigned int>(), cell_db.as<unsigned int>(),
                                                caps.lv_active);
    record_event(ev_p1);
    k_lv_prune<<<num_sms * 2, kThreads, 0, stream>>>(d_grid, d_cnt, d_params, lv_blk_slot.as<unsigned int>(),
                                                    lv_blk_flags.as<unsigned char>(), pool.as<unsigned char>());
    k_lv_finish<<<1, 1, 0, stream>>>(d_cnt);
    launches += 13 + 2 + (end_bit + 7) / 8;
}

size_t lv_qgrid_bytes() { return sizeof(QGrid); }

}  // namespace la3dm_b200
