// la3dm_b200 -- SURVEY.md section 8(f) rows next to the hot path:
//   * point query:  BGKOctoMap::search(point3f) / search(x, y, z)  (src/bgkoctomap/bgkoctomap.cpp:554-574, the -L/-LV/GP
//     copies) -> Block::search / get_index / get_node (src/bgkoctomap/bgkblock.cpp:132-156), batched on the GPU over
//     the device block table;
//   * import of a whole map in the reference's Block/OcTree layout (inverse of la3dm_export_blocks) and map
//     serialisation on top of it (the reference only has the unused node stream operators,
//     src/bgkoctomap/bgkoctree_node.cpp:46-58).
#include <cstdio>
#include <cstring>
#include <memory>

#include "engine.cuh"
#include "hash.cuh"
#include "leaf.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kThreads = 256;

// One thread per query point.
//   block:  block_to_hash_key(p) (bgkblock.cpp:73-77) -> key -> slot table; an unknown block answers like the reference's
//           `return OcTreeNode()` (a default node), marked depth = -1;
//   cell:   Block::get_index (bgkblock.cpp:139-147): int((p - centre) / resolution + cell_num / 2) per axis, clipped,
//           with cell_num = 2^(depth - 1) -- upstream freezes Block::cell_num at 8 during static initialisation
//           (bgkblock.cpp:105), which is only right for block_depth 4; this is the intended formula;
//   node:   Block::get_node -> index_map (init_index_map, bgkblock.cpp:34-67: finest nodes sorted by z, y, x) = the
//           finest node whose index interleaves the cell's bits, x -> 4, y -> 2, z -> 1 per level;
//   answer: finest_only != 0: that finest node even when it is PRUNED (what upstream's operator[] hands back);
//           else the LEAF that contains the point (walk up while PRUNED).
__global__ void k_search(const float *__restrict__ q, unsigned int n, int stride_f,
                         const long long *__restrict__ hkeys, const int *__restrict__ hvals, size_t mask,
                         const unsigned char *__restrict__ pool, const float3 *__restrict__ lut,
                         const DevParams *__restrict__ Pg, int finest_only, la3dm_leaf *out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevParams &P = *Pg;
    const float *p = q + (size_t) i * stride_f;
    const float x = p[0], y = p[1], z = p[2];
    const float bs = P.block_size;
    const long long ix = axis_index(x, bs), iy = axis_index(y, bs), iz = axis_index(z, bs);
    const long long key = make_key(ix, iy, iz);
    const float cx = axis_center(ix, bs), cy = axis_center(iy, bs), cz = axis_center(iz, bs);
    const int slot = hash_find(hkeys, hvals, mask, key);
    if (slot < 0) {
        la3dm_leaf L = make_leaf(P, key, 0, 0, make_float2(P.def_a, P.def_b), (unsigned char) LA3DM_UNKNOWN,
                                 make_float3(0.f, 0.f, 0.f), cx, cy, cz);
        L.depth = -1; L.index = -1;
        out[i] = L;
        return;
    }
    const int cells = 1 << (P.depth - 1);
    const int half = cells / 2;
    int c[3];
    c[0] = (int) ((x - cx) / P.resolution + (float) half);
    c[1] = (int) ((y - cy) / P.resolution + (float) half);
    c[2] = (int) ((z - cz) / P.resolution + (float) half);
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = max(0, min(c[a], cells - 1));
    int idx = 0;
    for (int l = P.depth - 2; l >= 0; --l)
        idx = (idx << 3) | (((c[0] >> l) & 1) << 2) | (((c[1] >> l) & 1) << 1) | ((c[2] >> l) & 1);
    const unsigned char *rec = pool + (size_t) slot * P.rec_bytes;
    const unsigned char *bst = rec + P.st_off;
    const float2 *bab = reinterpret_cast<const float2 *>(rec);
    int d = P.depth - 1;
    if (!finest_only)
        while (d > 0 && (bst[P.layer_off[d] + idx] & 7) == P.pruned_state) { --d; idx >>= 3; }
    const int node = P.layer_off[d] + idx;
    out[i] = make_leaf(P, key, d, idx, bab[node], bst[node], lut[node], cx, cy, cz);
}

// BGKOctoMap::RayCaster (include/bgkoctomap/bgkoctomap.h:91-214 and the -L/-LV/GP copies): the integer walk over the
// finest cells between two points, crossing blocks, one thread per ray.  Everything -- including the step accounting
// (`n -= 2` plus `n--` on an xy tie, a step without a move when no branch matches) and the unsigned-short wrap of a
// cell index stepping below 0 -- follows upstream statement by statement; cell indices use cell_num = 2^(depth-1) like
// la3dm_search (upstream mixes `lim` with the frozen Block::cell_num, identical at block_depth 4).
// Step i of ray r -> out[r * max_steps + i]: the finest node of the cell (what operator[] returns, PRUNED or not) as a
// leaf record with x y z = Block::get_point; where the block does not exist: depth = -1, x y z = the tracked position.
__global__ void k_raycast(const float *__restrict__ se, unsigned int n_rays, unsigned int max_steps,
                          const long long *__restrict__ hkeys, const int *__restrict__ hvals, size_t mask,
                          const unsigned char *__restrict__ pool, const float3 *__restrict__ lut,
                          const DevParams *__restrict__ Pg, la3dm_leaf *out, int *n_steps) {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const DevParams &P = *Pg;
    const float sx = se[6 * r], sy = se[6 * r + 1], sz = se[6 * r + 2];
    const float ex = se[6 * r + 3], ey = se[6 * r + 4], ez = se[6 * r + 5];
    const float bs = P.block_size, res = P.resolution;
    const int lim = 1 << (P.depth - 1), half = lim / 2;
    long long key = make_key(axis_index(sx, bs), axis_index(sy, bs), axis_index(sz, bs));
    int slot = hash_find(hkeys, hvals, mask, key);
    la3dm_leaf *o = out + (size_t) r * max_steps;
    int steps = 0;
    if (slot < 0) { n_steps[r] = 0; return; }                    // upstream: n = 0 when the start block does not exist
    // hash_key_to_block of the start block = block->get_center()
    float blx = axis_center(key >> 40, bs), bly = axis_center((key >> 20) & 0xFFFFF, bs), blz = axis_center(key & 0xFFFFF, bs);
    float bcx = blx, bcy = bly, bcz = blz;                       // centre of the CURRENT block (get_point)
    // Block::get_index (bgkblock.cpp:139-147)
    auto clip = [&](int a) { return max(0, min(a, lim - 1)); };
    unsigned short x = (unsigned short) clip((int) ((sx - blx) / res + (float) half));
    unsigned short y = (unsigned short) clip((int) ((sy - bly) / res + (float) half));
    unsigned short z = (unsigned short) clip((int) ((sz - blz) / res + (float) half));
    float cpx = sx, cpy = sy, cpz = sz;
    const int x0 = (int) (sx / res), y0 = (int) (sy / res), z0 = (int) (sz / res);
    const int x1 = (int) (ex / res), y1 = (int) (ey / res), z1 = (int) (ez / res);
    int dx = abs(x1 - x0), dy = abs(y1 - y0), dz = abs(z1 - z0);
    int n = 1 + dx + dy + dz;
    const int x_inc = x1 > x0 ? 1 : (x1 == x0 ? 0 : -1), y_inc = y1 > y0 ? 1 : (y1 == y0 ? 0 : -1),
              z_inc = z1 > z0 ? 1 : (z1 == z0 ? 0 : -1);
    int xy_error = dx - dy, xz_error = dx - dz, yz_error = dy - dz;
    dx *= 2; dy *= 2; dz *= 2;
    while (n > 0 && steps < (int) max_steps) {
        // index_map[x + y lim + z lim^2]: the finest node whose index interleaves the cell's bits (bgkblock.cpp:34-67)
        int idx = 0;
        for (int l = P.depth - 2; l >= 0; --l)
            idx = (idx << 3) | (((x >> l) & 1) << 2) | (((y >> l) & 1) << 1) | ((z >> l) & 1);
        const int node = P.layer_off[P.depth - 1] + idx;
        la3dm_leaf L;
        if (slot >= 0) {
            const unsigned char *rec = pool + (size_t) slot * P.rec_bytes;
            L = make_leaf(P, key, P.depth - 1, idx, reinterpret_cast<const float2 *>(rec)[node], rec[P.st_off + node],
                          lut[node], bcx, bcy, bcz);
            cpx = L.x; cpy = L.y; cpz = L.z;                      // current_p = block->get_point(x, y, z)
        } else {
            L = make_leaf(P, key, P.depth - 1, idx, make_float2(P.def_a, P.def_b), (unsigned char) LA3DM_UNKNOWN,
                          make_float3(0.f, 0.f, 0.f), 0.f, 0.f, 0.f);
            L.depth = -1;
            L.x = cpx; L.y = cpy; L.z = cpz;
        }
        o[steps++] = L;
        auto cross = [&](float &bl, int inc, unsigned short &c) {
            bl += (float) inc * bs;
            key = make_key(axis_index(blx, bs), axis_index(bly, bs), axis_index(blz, bs));
            slot = hash_find(hkeys, hvals, mask, key);
            bcx = axis_center(key >> 40, bs); bcy = axis_center((key >> 20) & 0xFFFFF, bs); bcz = axis_center(key & 0xFFFFF, bs);
            c = inc > 0 ? 0 : (unsigned short) (lim - 1);
        };
        if (xy_error > 0 && xz_error > 0) {
            x = (unsigned short) (x + x_inc); cpx += (float) x_inc * res; xy_error -= dy; xz_error -= dz;
            if (x >= lim) cross(blx, x_inc, x);
        } else if (xy_error < 0 && yz_error > 0) {
            y = (unsigned short) (y + y_inc); cpy += (float) y_inc * res; xy_error += dx; yz_error -= dz;
            if (y >= lim) cross(bly, y_inc, y);
        } else if (yz_error < 0 && xz_error < 0) {
            z = (unsigned short) (z + z_inc); cpz += (float) z_inc * res; xz_error += dx; yz_error += dy;
            if (z >= lim) cross(blz, z_inc, z);
        } else if (xy_error == 0) {
            x = (unsigned short) (x + x_inc); y = (unsigned short) (y + y_inc);
            n -= 2;
            cpx += (float) x_inc * res; cpy += (float) y_inc * res;
            if (x >= lim) cross(blx, x_inc, x);
            if (y >= lim) cross(bly, y_inc, y);
        }
        --n;
    }
    n_steps[r] = steps;
}

// nodes in the reference's Occupancy layout -> block records; one thread per node, the block's first thread also fills
// the spare bytes behind the states (leaf count read by k_predict_bgk's early-out, zero padding)
__global__ void k_unpack_nodes(const la3dm_node *__restrict__ in, unsigned int n_blocks, const DevParams *__restrict__ Pg,
                               unsigned char *pool) {
    const DevParams &P = *Pg;
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) n_blocks * P.nodes) return;
    const unsigned int b = (unsigned int) (i / P.nodes), n = (unsigned int) (i % P.nodes);
    unsigned char *rec = pool + (size_t) b * P.rec_bytes;
    const la3dm_node v = in[i];
    reinterpret_cast<float2 *>(rec)[n] = make_float2(v.a, v.b);
    rec[P.st_off + n] = (unsigned char) ((v.state & 7) | (v.classified ? 0x80 : 0));
    if (n == 0) {
        const la3dm_node *blk = in + (size_t) b * P.nodes;
        unsigned int leaves = 0;
        for (int d = 0; d < P.depth; ++d) {
            const int cnt = 1 << (3 * d);
            for (int k = 0; k < cnt; ++k) {
                if ((blk[P.layer_off[d] + k].state & 7) == P.pruned_state) continue;
                if (d + 1 == P.depth || (blk[P.layer_off[d + 1] + 8 * k].state & 7) == P.pruned_state) ++leaves;
            }
        }
        for (int k = P.st_off + P.nodes; k < P.rec_bytes; ++k)
            rec[k] = k == P.st_off + P.nodes ? (unsigned char) (leaves > 255u ? 255u : leaves) : 0;
    }
}

// on-disk header of la3dm_save (little endian, the struct as it lies in memory on x86-64 / aarch64)
struct FileHeader {
    char magic[8];            // "LA3DMB2\0"
    uint32_t version;         // 1
    uint32_t method;
    la3dm_params params;
    uint64_t n_blocks;
    uint32_t nodes_per_block;
    uint32_t node_bytes;      // sizeof(la3dm_node)
};
const char kMagic[8] = {'L', 'A', '3', 'D', 'M', 'B', '2', '\0'};

struct FileCloser {
    void operator()(FILE *f) const { if (f) fclose(f); }
};

}  // namespace

void Map::search(const float *xyz, size_t n, size_t stride_bytes, bool device_ptr, int finest_only, la3dm_leaf *out) {
    if (stride_bytes < 12 || stride_bytes % 4 != 0) throw StatusError{LA3DM_ERR_INVALID, "stride_bytes must be a multiple of 4, >= 12"};
    if (n > 0x7FFFFFF0ull) throw StatusError{LA3DM_ERR_INVALID, "too many query points"};
    if (n == 0) return;
    check_synced();
    if (!xyz || !out) throw StatusError{LA3DM_ERR_INVALID, "null query / output"};
    LA3DM_CUDA(cudaSetDevice(device));
    const float *d_q = xyz;
    if (!device_ptr) {
        export_tmp.reserve(n * stride_bytes, stream);
        LA3DM_CUDA(cudaMemcpyAsync(export_tmp.p, xyz, n * stride_bytes, cudaMemcpyHostToDevice, stream));
        d_q = export_tmp.as<float>();
    }
    leaf_out.reserve(n * sizeof(la3dm_leaf), stream);
    k_search<<<ceil_div((long long) n, kThreads), kThreads, 0, stream>>>(
        d_q, (unsigned int) n, (int) (stride_bytes / 4), hkeys.as<long long>(), hvals.as<int>(), hash_cap - 1,
        pool.as<unsigned char>(), d_lut, d_params, finest_only, leaf_out.as<la3dm_leaf>());
    LA3DM_CUDA(cudaMemcpyAsync(out, leaf_out.p, n * sizeof(la3dm_leaf), cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
}

void Map::raycast(const float *start_end, size_t n_rays, size_t max_steps, la3dm_leaf *out, int32_t *n_steps) {
    if (n_rays == 0) return;
    if (!start_end || !out || !n_steps || max_steps == 0) throw StatusError{LA3DM_ERR_INVALID, "raycast: null argument"};
    if (n_rays > 0x7FFFFFF0ull || n_rays * max_steps > 0x7FFFFFF0ull) throw StatusError{LA3DM_ERR_INVALID, "raycast: too many steps"};
    check_synced();
    LA3DM_CUDA(cudaSetDevice(device));
    export_tmp.reserve(n_rays * 6 * sizeof(float) + n_rays * sizeof(int32_t), stream);
    float *d_se = export_tmp.as<float>();
    int *d_n = reinterpret_cast<int *>(d_se + n_rays * 6);
    LA3DM_CUDA(cudaMemcpyAsync(d_se, start_end, n_rays * 6 * sizeof(float), cudaMemcpyHostToDevice, stream));
    leaf_out.reserve(n_rays * max_steps * sizeof(la3dm_leaf), stream);
    k_raycast<<<ceil_div((long long) n_rays, 128), 128, 0, stream>>>(
        d_se, (unsigned int) n_rays, (unsigned int) max_steps, hkeys.as<long long>(), hvals.as<int>(), hash_cap - 1,
        pool.as<unsigned char>(), d_lut, d_params, leaf_out.as<la3dm_leaf>(), d_n);
    LA3DM_CUDA(cudaMemcpyAsync(n_steps, d_n, n_rays * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaMemcpyAsync(out, leaf_out.p, n_rays * max_steps * sizeof(la3dm_leaf), cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
}

void Map::import_blocks(const int64_t *in_keys, const la3dm_node *in_nodes, size_t n) {
    if (n_blocks != 0) throw StatusError{LA3DM_ERR_INVALID, "import_blocks: the map must be empty"};
    if (n == 0) return;
    if (!in_keys || !in_nodes) throw StatusError{LA3DM_ERR_INVALID, "import_blocks: null input"};
    if (n > 0x7FFFFFF0ull / (size_t) hp.nodes) throw StatusError{LA3DM_ERR_INVALID, "import_blocks: too many blocks"};
    // keys strictly increasing (sorted, no duplicates anywhere) and inside the 3 x 20-bit key space; node states valid
    for (size_t i = 0; i < n; ++i) {
        if (in_keys[i] < 0 || in_keys[i] >= ((int64_t) 1 << 60)) throw StatusError{LA3DM_ERR_INVALID, "import_blocks: block key out of range"};
        if (i && in_keys[i] <= in_keys[i - 1]) throw StatusError{LA3DM_ERR_INVALID, "import_blocks: block keys must be strictly increasing"};
    }
    for (size_t i = 0; i < n * (size_t) hp.nodes; ++i)
        if (in_nodes[i].state > (uint8_t) hp.pruned_state) throw StatusError{LA3DM_ERR_INVALID, "import_blocks: invalid node state"};
    LA3DM_CUDA(cudaSetDevice(device));
    ensure_pool(n + caps.tests);
    const size_t total = n * (size_t) hp.nodes;
    export_buf.reserve(total * sizeof(la3dm_node), stream);
    LA3DM_CUDA(cudaMemcpyAsync(keys.p, in_keys, n * 8, cudaMemcpyHostToDevice, stream));
    LA3DM_CUDA(cudaMemcpyAsync(export_buf.p, in_nodes, total * sizeof(la3dm_node), cudaMemcpyHostToDevice, stream));
    k_unpack_nodes<<<ceil_div((long long) total, kThreads), kThreads, 0, stream>>>(
        export_buf.as<la3dm_node>(), (unsigned int) n, d_params, pool.as<unsigned char>());
    LA3DM_CUDA(cudaMemsetAsync(touched.p, 1, n, stream));     // an imported block is news to an incremental mirror
    n_blocks = (long long) n;
    try {
        rebuild_hash();
        LA3DM_CUDA(cudaStreamSynchronize(stream));
    } catch (...) {
        n_blocks = 0;          // a failed import leaves an empty map, not a half-filled one
        throw;
    }
}

void Map::save(const char *path) {
    if (!path) throw StatusError{LA3DM_ERR_INVALID, "save: null path"};
    const size_t n = (size_t) n_blocks;
    std::vector<int64_t> k(n);
    std::vector<la3dm_node> nd(n * (size_t) hp.nodes);
    size_t got = 0;
    export_blocks(k.data(), nd.data(), n, &got);
    std::unique_ptr<FILE, FileCloser> f(fopen(path, "wb"));
    if (!f) throw StatusError{LA3DM_ERR_INVALID, std::string("save: cannot open ") + path};
    FileHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, kMagic, 8);
    h.version = 1; h.method = (uint32_t) hp.method; h.params = api_params; h.n_blocks = n;
    h.nodes_per_block = (uint32_t) hp.nodes; h.node_bytes = (uint32_t) sizeof(la3dm_node);
    bool ok = fwrite(&h, sizeof(h), 1, f.get()) == 1;
    ok = ok && (n == 0 || fwrite(k.data(), 8, n, f.get()) == n);
    ok = ok && (n == 0 || fwrite(nd.data(), sizeof(la3dm_node), nd.size(), f.get()) == nd.size());
    ok = ok && fflush(f.get()) == 0;
    if (!ok) throw StatusError{LA3DM_ERR_INVALID, std::string("save: short write to ") + path};
}

void Map::load(const char *path) {
    if (!path) throw StatusError{LA3DM_ERR_INVALID, "load: null path"};
    if (n_blocks != 0) throw StatusError{LA3DM_ERR_INVALID, "load: the map must be empty"};
    std::unique_ptr<FILE, FileCloser> f(fopen(path, "rb"));
    if (!f) throw StatusError{LA3DM_ERR_INVALID, std::string("load: cannot open ") + path};
    FileHeader h;
    if (fread(&h, sizeof(h), 1, f.get()) != 1 || std::memcmp(h.magic, kMagic, 8) != 0 || h.version != 1)
        throw StatusError{LA3DM_ERR_INVALID, "load: not a la3dm_b200 map file"};
    // the node arrays only mean something under the parameters they were built with (thresholds, priors, depth)
    if (h.method != (uint32_t) hp.method || std::memcmp(&h.params, &api_params, sizeof(la3dm_params)) != 0 ||
        h.nodes_per_block != (uint32_t) hp.nodes || h.node_bytes != sizeof(la3dm_node))
        throw StatusError{LA3DM_ERR_INVALID, "load: the file was written by a map with a different method / parameters"};
    const size_t n = (size_t) h.n_blocks;
    if (n > 0x7FFFFFF0ull / (size_t) hp.nodes) throw StatusError{LA3DM_ERR_INVALID, "load: corrupt block count"};
    std::vector<int64_t> k(n);
    std::vector<la3dm_node> nd(n * (size_t) hp.nodes);
    if (n && (fread(k.data(), 8, n, f.get()) != n || fread(nd.data(), sizeof(la3dm_node), nd.size(), f.get()) != nd.size()))
        throw StatusError{LA3DM_ERR_INVALID, "load: truncated file"};
    import_blocks(k.data(), nd.data(), n);
}

}  // namespace la3dm_b200
