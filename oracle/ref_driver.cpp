// TEST INFRASTRUCTURE ONLY.  C-callable driver around the UNMODIFIED reference map classes, compiled together with
// the reference's own sources straight from /root/reference (see oracle/Makefile) into oracle/_ref/libla3dm_ref_<m>.so.
// Nothing here is linked into, or called by, the product library (la3dm_b200/csrc); it is the checker for tests/,
// the golden-vector generator (tests/golden/make_golden.py) and the `--impl reference` / cpu_baseline arm of bench.py.
//
// One shared object per method, because all four reference methods reuse the class names la3dm::Block / OcTree /
// Occupancy.  Select with -DREF_BGK | -DREF_BGKL | -DREF_BGKLV | -DREF_GP.
//
// `#define private public` lets the dump read m_A/m_B (private in the reference); it does not change any reference
// translation unit, only this driver's view of the headers.
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public
#define protected public
#if defined(REF_BGK)
#include "bgkoctomap.h"
typedef la3dm::BGKOctoMap MapT;
#elif defined(REF_BGKL)
#include "bgkloctomap.h"
typedef la3dm::BGKLOctoMap MapT;
#elif defined(REF_BGKLV)
#include "bgklvoctomap.h"
typedef la3dm::BGKLVOctoMap MapT;
#elif defined(REF_GP)
#include "gpoctomap.h"
typedef la3dm::GPOctoMap MapT;
#else
#error "select a method"
#endif
#undef private
#undef protected

using la3dm::point3f;

extern "C" {

// params layout (floats), mirroring the reference constructors:
//  BGK / BGKL : resolution, block_depth, sf2, ell, free_thresh, occupied_thresh, var_thresh, prior_A, prior_B
//               (include/bgkoctomap/bgkoctomap.h:50-58, include/bgkloctomap/bgkloctomap.h:53-61)
//  BGKLV      : the nine above + original_size, min_W          (include/bgklvoctomap/bgklvoctomap.h:52-62)
//  GP         : resolution, block_depth, sf2, ell, noise, l, min_var, max_var, max_known_var, free_thresh,
//               occupied_thresh                                   (include/gpoctomap/gpoctomap.h:50-52)
void *ref_create(const float *p, int n) {
#if defined(REF_BGK) || defined(REF_BGKL)
    if (n < 9) return nullptr;
    return new MapT(p[0], (unsigned short) p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
#elif defined(REF_BGKLV)
    if (n < 11) return nullptr;
    return new MapT(p[0], (unsigned short) p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9] != 0.0f, p[10]);
#else
    if (n < 11) return nullptr;
    return new MapT(p[0], (unsigned short) p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10]);
#endif
}

void ref_destroy(void *h) { delete static_cast<MapT *>(h); }

void ref_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}

int ref_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void fill_cloud(la3dm::PCLPointCloud &cloud, const float *xyz, int64_t n) {
    cloud.points.resize((size_t) n);
    for (int64_t i = 0; i < n; ++i) cloud.points[i] = la3dm::PCLPointType(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    cloud.width = (uint32_t) n;
    cloud.height = 1;
    cloud.is_dense = true;
}

// The call the static nodes / servers make: map.insert_pointcloud(cloud, origin, ds_resolution, free_res, max_range)
// (src/bgkoctomap/bgkoctomap_static_node.cpp:95).
void ref_insert_pointcloud(void *h, const float *xyz, int64_t n, const float *origin, float ds_resolution,
                           float free_res, float max_range) {
    la3dm::PCLPointCloud cloud;
    fill_cloud(cloud, xyz, n);
    point3f o(origin[0], origin[1], origin[2]);
    // the -L/-LV reference prints "Sampled points: N" to std::cout on every call; silence it
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    static_cast<MapT *>(h)->insert_pointcloud(cloud, o, ds_resolution, free_res, max_range);
    std::cout.rdbuf(old);
}

// insert_training_data(const GPPointCloud &xy) (bgkoctomap.h:86, gpoctomap.h): BGK and GP only.  NOTE upstream
// dereferences a null Block* for a test block that does not exist yet (bgkoctomap.cpp:155-160): callers of this
// checker only pass points whose test blocks were created by an earlier insert_pointcloud.
int ref_insert_training_data(void *h, const float *xyzy, int64_t n) {
#if defined(REF_BGK) || defined(REF_GP)
    MapT::GPPointCloud xy;
    for (int64_t i = 0; i < n; ++i)
        xy.emplace_back(point3f(xyzy[4 * i], xyzy[4 * i + 1], xyzy[4 * i + 2]), xyzy[4 * i + 3]);
    static_cast<MapT *>(h)->insert_training_data(xy);
    return 0;
#else
    (void) h; (void) xyzy; (void) n;
    return -1;
#endif
}

int64_t ref_num_blocks(void *h) { return (int64_t) static_cast<MapT *>(h)->block_arr.size(); }

int64_t ref_num_leaves(void *h) {
    MapT *m = static_cast<MapT *>(h);
    int64_t n = 0;
    for (auto it = m->block_arr.cbegin(); it != m->block_arr.cend(); ++it)
        for (auto l = it->second->begin_leaf(); l != it->second->end_leaf(); ++l) ++n;
    return n;
}

// Per leaf: block key, node (depth,index), centre xyz + size, the two node floats (m_A,m_B | m_ivar,ivar), state,
// classified, get_prob(), get_var().  Order = unordered_map order x reference DFS; callers sort.
void ref_dump_leaves(void *h, int64_t *block_key, int32_t *depth, int32_t *index, float *loc_size, float *ab,
                     uint8_t *state, uint8_t *classified, float *prob_var) {
    MapT *m = static_cast<MapT *>(h);
    int64_t n = 0;
    for (auto it = m->block_arr.cbegin(); it != m->block_arr.cend(); ++it) {
        la3dm::Block *b = it->second;
        for (auto l = b->begin_leaf(); l != b->end_leaf(); ++l, ++n) {
            la3dm::OcTreeHashKey k = l.get_hash_key();
            block_key[n] = it->first;
#if defined(REF_BGKLV)
            depth[n] = (int32_t) (k >> 28);
            index[n] = (int32_t) (k & 0xFFFFFFF);
#else
            depth[n] = (int32_t) (k >> 16);
            index[n] = (int32_t) (k & 0xFFFF);
#endif
            point3f p = b->get_loc(l);
            loc_size[4 * n + 0] = p.x();
            loc_size[4 * n + 1] = p.y();
            loc_size[4 * n + 2] = p.z();
            loc_size[4 * n + 3] = b->get_size(l);
            la3dm::OcTreeNode &node = l.get_node();
#if defined(REF_GP)
            ab[2 * n] = node.m_ivar;
            ab[2 * n + 1] = node.ivar;
#else
            ab[2 * n] = node.m_A;
            ab[2 * n + 1] = node.m_B;
#endif
            state[n] = (uint8_t) node.get_state();
            classified[n] = node.classified ? 1 : 0;
            prob_var[2 * n] = node.get_prob();
            prob_var[2 * n + 1] = node.get_var();
        }
    }
}

// search(x, y, z) of the reference map (src/bgkoctomap/bgkoctomap.cpp:554-574 -> Block::search, bgkblock.cpp:132-156),
// n query points: the two node floats, state, classified of the node upstream returns (a default node where the block
// does not exist).  NOTE upstream's Block::cell_num is frozen at 8 (bgkblock.cpp:105): only meaningful for block_depth 4.
void ref_search(void *h, const float *xyz, int64_t n, float *ab, uint8_t *state, uint8_t *classified) {
    MapT *m = static_cast<MapT *>(h);
    for (int64_t i = 0; i < n; ++i) {
        la3dm::OcTreeNode node = m->search(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
#if defined(REF_GP)
        ab[2 * i] = node.m_ivar;
        ab[2 * i + 1] = node.ivar;
#else
        ab[2 * i] = node.m_A;
        ab[2 * i + 1] = node.m_B;
#endif
        state[i] = (uint8_t) node.get_state();
        classified[i] = node.classified ? 1 : 0;
    }
}

// RayCaster(map, start, end) walked to its end (at most max_steps): per step the point, block key, node key
// (depth << 16 | index; LV: << 28), valid flag and the node's two floats / state.  Returns the number of steps.
int64_t ref_raycast(void *h, const float *start, const float *end, int64_t max_steps, float *p_out, int64_t *block_key,
                    int64_t *node_key, uint8_t *valid, float *ab, uint8_t *state) {
    MapT *m = static_cast<MapT *>(h);
    MapT::RayCaster rc(m, point3f(start[0], start[1], start[2]), point3f(end[0], end[1], end[2]));
    int64_t i = 0;
    while (!rc.end() && i < max_steps) {
        point3f p;
        la3dm::OcTreeNode node;
        la3dm::BlockHashKey bk;
        la3dm::OcTreeHashKey nk;
        const bool ok = rc.next(p, node, bk, nk);
        p_out[3 * i] = p.x(); p_out[3 * i + 1] = p.y(); p_out[3 * i + 2] = p.z();
        block_key[i] = (int64_t) bk;
        node_key[i] = (int64_t) nk;
        valid[i] = ok ? 1 : 0;
#if defined(REF_GP)
        ab[2 * i] = node.m_ivar; ab[2 * i + 1] = node.ivar;
#else
        ab[2 * i] = node.m_A; ab[2 * i + 1] = node.m_B;
#endif
        state[i] = (uint8_t) node.get_state();
        ++i;
    }
    return i;
}

void ref_get_bbox(void *h, float *mn, float *mx) {
    point3f a, b;
    static_cast<MapT *>(h)->get_bbox(a, b);
    mn[0] = a.x(); mn[1] = a.y(); mn[2] = a.z();
    mx[0] = b.x(); mx[1] = b.y(); mx[2] = b.z();
}

// Front-end only (get_training_data is a private const member of the reference map).  Returns the number of
// training entries; when out != nullptr writes 7 floats per entry: x0 y0 z0 x1 y1 z1 label  (for BGK/GP x1..z1 repeat
// x0..z0).  For -L/-LV also the ray table (6 floats per ray) and ray_idx per entry.
static thread_local std::vector<float> g_xy, g_rays;
static thread_local std::vector<int32_t> g_ray_idx;

int64_t ref_training_data(void *h, const float *xyz, int64_t n, const float *origin, float ds_resolution,
                          float free_res, float max_range, int64_t *n_rays) {
    MapT *m = static_cast<MapT *>(h);
    la3dm::PCLPointCloud cloud;
    fill_cloud(cloud, xyz, n);
    point3f o(origin[0], origin[1], origin[2]);
    g_xy.clear(); g_rays.clear(); g_ray_idx.clear();
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
#if defined(REF_BGK) || defined(REF_GP)
    MapT::GPPointCloud xy;
    m->get_training_data(cloud, o, ds_resolution, free_res, max_range, xy);
    for (auto &e : xy) {
        const float v[7] = {e.first.x(), e.first.y(), e.first.z(), e.first.x(), e.first.y(), e.first.z(), e.second};
        g_xy.insert(g_xy.end(), v, v + 7);
        g_ray_idx.push_back(-1);
    }
#else
    MapT::GPLineCloud xy, rays;
    std::vector<int> ray_idx;
#if defined(REF_BGKLV)
    if (ds_resolution > m->resolution) ds_resolution = m->resolution;   // bgklvoctomap.cpp:102-104
#endif
    m->get_training_data(cloud, o, ds_resolution, free_res, max_range, xy, rays, ray_idx);
    for (size_t i = 0; i < xy.size(); ++i) {
        const float v[7] = {xy[i].first.x0(), xy[i].first.y0(), xy[i].first.z0(),
                            xy[i].first.x1(), xy[i].first.y1(), xy[i].first.z1(), xy[i].second};
        g_xy.insert(g_xy.end(), v, v + 7);
        g_ray_idx.push_back(ray_idx[i]);
    }
    for (auto &r : rays) {
        const float v[6] = {r.first.x0(), r.first.y0(), r.first.z0(), r.first.x1(), r.first.y1(), r.first.z1()};
        g_rays.insert(g_rays.end(), v, v + 6);
    }
#endif
    std::cout.rdbuf(old);
    if (n_rays) *n_rays = (int64_t) (g_rays.size() / 6);
    return (int64_t) (g_xy.size() / 7);
}

void ref_training_data_copy(float *xy7, int32_t *ray_idx, float *rays6) {
    if (xy7) std::memcpy(xy7, g_xy.data(), g_xy.size() * sizeof(float));
    if (ray_idx) std::memcpy(ray_idx, g_ray_idx.data(), g_ray_idx.size() * sizeof(int32_t));
    if (rays6) std::memcpy(rays6, g_rays.data(), g_rays.size() * sizeof(float));
}

// Key helpers (src/bgkoctomap/bgkblock.cpp:69-101) for known-answer tests.
int64_t ref_block_to_hash_key(float x, float y, float z) { return la3dm::block_to_hash_key(x, y, z); }

void ref_hash_key_to_block(int64_t key, float *c) {
    point3f p = la3dm::hash_key_to_block(key);
    c[0] = p.x(); c[1] = p.y(); c[2] = p.z();
}

void ref_extended_block(int64_t key, int64_t *out7) {
    la3dm::ExtendedBlock e = la3dm::get_extended_block(key);
    for (int i = 0; i < 7; ++i) out7[i] = e[i];
}

// LUT entry: centre offset of node (depth,index) (src/bgkoctomap/bgkblock.cpp:7-32)
int ref_key_loc(int depth, int index, float *c) {
#if defined(REF_BGKLV)
    la3dm::OcTreeHashKey k = la3dm::node_to_hash_key((unsigned short) depth, (unsigned long) index);
#else
    la3dm::OcTreeHashKey k = la3dm::node_to_hash_key((unsigned short) depth, (unsigned short) index);
#endif
    auto it = la3dm::Block::key_loc_map.find(k);
    if (it == la3dm::Block::key_loc_map.end()) return 0;
    c[0] = it->second.x(); c[1] = it->second.y(); c[2] = it->second.z();
    return 1;
}

}  // extern "C"
