/*
 * la3dm_b200 -- C++ facade with the reference's class and method names over the C ABI (include/la3dm_b200.h).
 *
 * Drop-in for the part of
 *   la3dm::BGKOctoMap    (include/bgkoctomap/bgkoctomap.h:50-58, 82-84, 89, 217-319)
 *   la3dm::BGKLOctoMap   (include/bgkloctomap/bgkloctomap.h:53-61)
 *   la3dm::BGKLVOctoMap  (include/bgklvoctomap/bgklvoctomap.h:52-62)
 *   la3dm::GPOctoMap     (include/gpoctomap/gpoctomap.h:50-52)
 * that the reference's nodes use: constructor, insert_pointcloud, get_bbox, begin_leaf()/end_leaf() with
 * get_loc / get_size / get_node / get_pruned_locs, get_resolution / get_block_depth / get_block_size, search.
 * Header only; link with -lla3dm_b200.  All map state lives on the GPU; the leaf iterator walks a host mirror that is
 * refreshed lazily (one la3dm_export_leaves call) the first time it is used after an insert.
 *
 * The cloud / point types are template parameters so that the header does not depend on PCL:
 *   Cloud  : has `.points` (contiguous container of structs that start with float x, y, z), e.g.
 *            pcl::PointCloud<pcl::PointXYZ> (16-byte stride);
 *   Point  : has x(), y(), z() and a (float, float, float) constructor, e.g. la3dm::point3f.
 * In a tree that still has the reference's headers, define LA3DM_B200_NAMESPACE to something else than `la3dm` to
 * keep both implementations side by side.
 */
#ifndef LA3DM_B200_OCTOMAP_H
#define LA3DM_B200_OCTOMAP_H

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../la3dm_b200.h"

#ifndef LA3DM_B200_NAMESPACE
#define LA3DM_B200_NAMESPACE la3dm
#endif

namespace LA3DM_B200_NAMESPACE {

/// enum class State (include/bgkoctomap/bgkoctree_node.h:10-12; BGKLV: bgklvoctree_node.h:11-13)
enum class State : char { FREE, OCCUPIED, UNKNOWN, PRUNED };
enum class LVState : char { FREE, OCCUPIED, UNKNOWN, UNCERTAIN, PRUNED };

typedef int64_t BlockHashKey;

/// minimal stand-in for la3dm::point3f (include/common/point3f.h) for trees that do not have it
struct vec3f {
    float v[3];
    vec3f() : v{0.f, 0.f, 0.f} {}
    vec3f(float x, float y, float z) : v{x, y, z} {}
    float &x() { return v[0]; }
    float &y() { return v[1]; }
    float &z() { return v[2]; }
    const float &x() const { return v[0]; }
    const float &y() const { return v[1]; }
    const float &z() const { return v[2]; }
};

/// what LeafIterator::get_node() returns: read-only view of one Occupancy (include/bgkoctomap/bgkoctree_node.h:17-81)
class OcTreeNode {
public:
    explicit OcTreeNode(const la3dm_leaf *l = nullptr) : l_(l) {}
    explicit OcTreeNode(const la3dm_leaf &copy) : own_(copy), l_(&own_) {}
    OcTreeNode(const OcTreeNode &o) : own_(o.own_), l_(o.l_ == &o.own_ ? &own_ : o.l_) {}
    OcTreeNode &operator=(const OcTreeNode &o) { own_ = o.own_; l_ = o.l_ == &o.own_ ? &own_ : o.l_; return *this; }
    float get_prob() const { return l_ ? l_->prob : 0.5f; }
    float get_var() const { return l_ ? l_->var : 0.f; }
    State get_state() const { return l_ ? static_cast<State>(l_->state) : State::UNKNOWN; }
    int get_state_raw() const { return l_ ? l_->state : (int) LA3DM_UNKNOWN; }   /* BGKLV numbering: LVState */
    bool classified() const { return l_ && l_->classified; }
    float get_a() const { return l_ ? l_->a : 0.f; }   /* m_A | GP m_ivar */
    float get_b() const { return l_ ? l_->b : 0.f; }   /* m_B | GP ivar   */
private:
    la3dm_leaf own_ = la3dm_leaf();
    const la3dm_leaf *l_;
};

template <int METHOD>
class OctoMapT {
public:
    ~OctoMapT() { if (h_) la3dm_destroy(h_); }
    OctoMapT(const OctoMapT &) = delete;
    OctoMapT &operator=(const OctoMapT &) = delete;

    float get_resolution() const { return params_.resolution; }
    float get_block_depth() const { return (float) params_.block_depth; }   /* float upstream too (bgkoctomap.h:70) */
    float get_block_size() const { return (float) std::pow(2, params_.block_depth - 1) * params_.resolution; }

    /// insert_pointcloud (include/bgkoctomap/bgkoctomap.h:82-84)
    template <class Cloud, class Point>
    void insert_pointcloud(const Cloud &cloud, const Point &origin, float ds_resolution, float free_res = 2.0f,
                           float max_range = -1) {
        const float o[3] = {origin.x(), origin.y(), origin.z()};
        const size_t n = cloud.points.size();
        const float *xyz = n ? reinterpret_cast<const float *>(&cloud.points[0]) : nullptr;
        check(la3dm_insert_pointcloud(h_, xyz, n, sizeof(cloud.points[0]), o, ds_resolution, free_res, max_range));
        dirty_ = true;
    }
    /// same, from a raw host array of n records of stride_bytes that start with float x, y, z
    void insert_pointcloud(const float *xyz, size_t n, size_t stride_bytes, const float origin[3], float ds_resolution,
                           float free_res = 2.0f, float max_range = -1) {
        check(la3dm_insert_pointcloud(h_, xyz, n, stride_bytes, origin, ds_resolution, free_res, max_range));
        dirty_ = true;
    }

    /// insert_training_data (include/bgkoctomap/bgkoctomap.h:86; BGKOctoMap and GPOctoMap): GPPointCloud is
    /// std::vector<std::pair<point3f, float>> upstream -- any container of pairs (point with x() y() z(), label) works
    template <class GPPointCloud>
    void insert_training_data(const GPPointCloud &xy) {
        std::vector<float> buf;
        buf.reserve(4 * xy.size());
        for (auto it = xy.begin(); it != xy.end(); ++it) {
            buf.push_back(it->first.x()); buf.push_back(it->first.y()); buf.push_back(it->first.z());
            buf.push_back(it->second);
        }
        check(la3dm_insert_training_data(h_, buf.data(), buf.size() / 4, 16));
        dirty_ = true;
    }

    /// cloudHandler's cloud path (src/bgkoctomap/bgkoctomap_server.cpp:70-86) in one call: sensor-frame cloud, the
    /// map <- sensor transform as 3 x 4 row-major floats, VoxelGrid prefilter, insert if more than min_points are left
    template <class Cloud, class Point>
    void insert_pointcloud_ingest(const Cloud &cloud, const float tf[12], float prefilter_ds, const Point &origin,
                                  float ds_resolution, float free_res = 2.0f, float max_range = -1, int min_points = 5) {
        const float o[3] = {origin.x(), origin.y(), origin.z()};
        const size_t n = cloud.points.size();
        const float *xyz = n ? reinterpret_cast<const float *>(&cloud.points[0]) : nullptr;
        check(la3dm_insert_pointcloud_ingest(h_, xyz, n, sizeof(cloud.points[0]), tf, prefilter_ds, min_points, o,
                                             ds_resolution, free_res, max_range));
        dirty_ = true;
    }

    /// multi-GPU replicas (include/la3dm_b200.h, "multi-GPU, BGKOctoMap"); one map object per GPU
    void reserve_blocks(size_t n) { check(la3dm_reserve_blocks(h_, n)); }
    void peer_set_deferred(bool on) { check(la3dm_peer_set_deferred(h_, on ? 1 : 0)); }
    void peer_attach(int world, int rank, void *const *pool_bases, void *const *flags) {
        check(la3dm_peer_attach(h_, world, rank, pool_bases, flags));
    }
    void peer_sync() { check(la3dm_peer_sync(h_)); dirty_ = true; }

    /// get_bbox (src/bgkoctomap/bgkoctomap.cpp:368-381)
    template <class Point>
    void get_bbox(Point &lim_min, Point &lim_max) const {
        float mn[3], mx[3];
        check(la3dm_get_bbox(h_, mn, mx));
        lim_min = Point(mn[0], mn[1], mn[2]);
        lim_max = Point(mx[0], mx[1], mx[2]);
    }

    la3dm_scan_stats last_stats() const {
        la3dm_scan_stats s;
        la3dm_last_stats(h_, &s);
        return s;
    }
    size_t num_blocks() const { return (size_t) la3dm_num_blocks(h_); }
    la3dm_map *handle() const { return h_; }

    /// LeafIterator (include/bgkoctomap/bgkoctomap.h:217-307).  Order: blocks by key, leaves by (depth, index).
    class LeafIterator {
    public:
        LeafIterator(const OctoMapT *m, size_t i) : m_(m), i_(i) {}
        bool operator==(const LeafIterator &o) const { return i_ == o.i_; }
        bool operator!=(const LeafIterator &o) const { return i_ != o.i_; }
        LeafIterator &operator++() { ++i_; return *this; }
        LeafIterator operator++(int) { LeafIterator r(*this); ++i_; return r; }
        OcTreeNode operator*() const { return OcTreeNode(&m_->leaves_[i_]); }
        OcTreeNode get_node() const { return OcTreeNode(&m_->leaves_[i_]); }
        vec3f get_loc() const { const la3dm_leaf &l = m_->leaves_[i_]; return vec3f(l.x, l.y, l.z); }
        float get_size() const { return m_->leaves_[i_].size; }
        const la3dm_leaf &raw() const { return m_->leaves_[i_]; }
        /// centres of the finest voxels a pruned leaf covers (bgkoctomap.h:265-283, same float stepping)
        std::vector<vec3f> get_pruned_locs() const {
            std::vector<vec3f> out;
            const la3dm_leaf &l = m_->leaves_[i_];
            const float res = m_->params_.resolution, size = l.size;
            const float x0 = l.x - size * 0.5 + res * 0.5, y0 = l.y - size * 0.5 + res * 0.5, z0 = l.z - size * 0.5 + res * 0.5;
            const float x1 = l.x + size * 0.5, y1 = l.y + size * 0.5, z1 = l.z + size * 0.5;
            for (float x = x0; x < x1; x += res)
                for (float y = y0; y < y1; y += res)
                    for (float z = z0; z < z1; z += res) out.emplace_back(x, y, z);
            return out;
        }
    private:
        const OctoMapT *m_;
        size_t i_;
    };
    LeafIterator begin_leaf() const { refresh(); return LeafIterator(this, 0); }
    LeafIterator end_leaf() const { refresh(); return LeafIterator(this, leaves_.size()); }
    size_t num_leaves() const { refresh(); return leaves_.size(); }

    /// The server loop without the full-map walk (src/bgkoctomap/bgkoctomap_server.cpp:94-144 visits every leaf of the
    /// map after every scan and keeps the OCCUPIED / FREE ones): the leaves in the wanted states of the blocks the scans
    /// since the previous call touched, plus the keys of those blocks -- replace what the marker arrays hold for the
    /// listed blocks (la3dm_export_touched).  state_mask: bit s = state s.
    void touched_leaves(unsigned int state_mask, std::vector<la3dm_leaf> &leaves, std::vector<int64_t> &block_keys) const {
        size_t nl = 0, nb = 0;
        check(la3dm_export_touched(h_, state_mask, nullptr, 0, &nl, nullptr, 0, &nb, 0));
        leaves.resize(nl ? nl : 1);
        block_keys.resize(nb);
        if (nb) check(la3dm_export_touched(h_, state_mask, leaves.data(), leaves.size(), &nl, block_keys.data(), nb, &nb, 1));
        leaves.resize(nl);
    }

    /// search(x, y, z) (include/bgkoctomap/bgkoctomap.h:315-319): the leaf that holds the point, looked up on the
    /// device (la3dm_search); an UNKNOWN default node if the block does not exist, like upstream's `OcTreeNode()`.
    /// The reference's Block::search is only right for block_depth 4; this one is right for any depth.
    OcTreeNode search(float x, float y, float z) const {
        const float q[3] = {x, y, z};
        la3dm_leaf l;
        check(la3dm_search(h_, q, 1, sizeof(q), 0, &l));
        return l.depth < 0 ? OcTreeNode() : OcTreeNode(l);
    }
    /// batch form: n points (x y z at the start of each stride_bytes record) -> n leaf records (depth -1: no block)
    void search(const float *xyz, size_t n, size_t stride_bytes, la3dm_leaf *out, bool finest_only = false) const {
        check(la3dm_search(h_, xyz, n, stride_bytes, finest_only ? 1 : 0, out));
    }
    /// RayCaster (include/bgkoctomap/bgkoctomap.h:91-214): `RayCaster ray(&map, start, end); while (!ray.end()) {
    /// if (ray.next(p, node, block_key, node_key)) ... }` -- the whole walk is done on the device at construction
    /// (la3dm_raycast), next() hands its steps out one by one.  node_key = (depth << 16) + index like OcTreeHashKey.
    class RayCaster {
    public:
        template <class Point>
        RayCaster(const OctoMapT *map, const Point &start, const Point &end, size_t max_steps = 4096) : i_(0) {
            const float se[6] = {start.x(), start.y(), start.z(), end.x(), end.y(), end.z()};
            steps_.resize(max_steps);
            int32_t n = 0;
            map->check(la3dm_raycast(map->h_, se, 1, max_steps, steps_.data(), &n));
            steps_.resize((size_t) n);
        }
        bool end() const { return i_ >= steps_.size(); }
        template <class Point>
        bool next(Point &p, OcTreeNode &node, BlockHashKey &block_key, uint32_t &node_key) {
            const la3dm_leaf &l = steps_[i_++];
            p = Point(l.x, l.y, l.z);
            block_key = l.block_key;
            const bool valid = l.depth >= 0;
            node_key = (uint32_t) ((valid ? l.depth : 0) << 16) + (uint32_t) l.index;
            if (valid) node = OcTreeNode(l);
            return valid;
        }
    private:
        std::vector<la3dm_leaf> steps_;
        size_t i_;
    };

    /// checkpoint / resume (la3dm_save / la3dm_load; load needs an empty map with the same parameters)
    void save(const std::string &path) const { check(la3dm_save(h_, path.c_str())); }
    void load(const std::string &path) { check(la3dm_load(h_, path.c_str())); dirty_ = true; }
    template <class Point>
    OcTreeNode search(const Point &p) const { return search(p.x(), p.y(), p.z()); }

protected:
    explicit OctoMapT(const la3dm_params &p, int device = 0) : params_(p) {
        const int rc = la3dm_create(METHOD, &params_, device, &h_);
        if (rc != LA3DM_OK) {
            const char *msg = la3dm_last_error(nullptr);
            throw std::runtime_error(std::string("la3dm_b200: ") + (msg && *msg ? msg : la3dm_status_string(rc)));
        }
    }
    void check(int rc) const {
        if (rc != LA3DM_OK) throw std::runtime_error(std::string("la3dm_b200: ") + la3dm_last_error(h_));
    }
    void refresh() const {
        if (!dirty_) return;
        size_t n = 0;
        check(la3dm_export_leaves(h_, nullptr, 0, &n));
        leaves_.resize(n);
        if (n) check(la3dm_export_leaves(h_, leaves_.data(), n, &n));
        block_range_.clear();
        for (size_t i = 0; i < n;) {
            size_t j = i;
            while (j < n && leaves_[j].block_key == leaves_[i].block_key) ++j;
            block_range_[leaves_[i].block_key] = std::make_pair(i, j);
            i = j;
        }
        dirty_ = false;
    }

    la3dm_params params_;
    la3dm_map *h_ = nullptr;
    mutable bool dirty_ = true;
    mutable std::vector<la3dm_leaf> leaves_;
    mutable std::unordered_map<BlockHashKey, std::pair<size_t, size_t>> block_range_;
};

inline la3dm_params make_bgk_params(float resolution, unsigned short block_depth, float sf2, float ell,
                                    float free_thresh, float occupied_thresh, float var_thresh, float prior_A,
                                    float prior_B) {
    la3dm_params p = la3dm_params();
    p.resolution = resolution; p.block_depth = block_depth; p.sf2 = sf2; p.ell = ell;
    p.free_thresh = free_thresh; p.occupied_thresh = occupied_thresh; p.var_thresh = var_thresh;
    p.prior_A = prior_A; p.prior_B = prior_B;
    return p;
}

/// la3dm::BGKOctoMap (include/bgkoctomap/bgkoctomap.h:50-58); defaults of the no-argument constructor: bgkoctomap.cpp:22-31
class BGKOctoMap : public OctoMapT<LA3DM_BGK> {
public:
    BGKOctoMap(float resolution, unsigned short block_depth, float sf2, float ell, float free_thresh,
               float occupied_thresh, float var_thresh, float prior_A, float prior_B, int device = 0)
        : OctoMapT(make_bgk_params(resolution, block_depth, sf2, ell, free_thresh, occupied_thresh, var_thresh, prior_A,
                                   prior_B), device) {}
    BGKOctoMap() : BGKOctoMap(0.1f, 4, 1.0f, 1.0f, 0.3f, 0.7f, 1.0f, 1.0f, 1.0f) {}
};

/// la3dm::BGKLOctoMap (include/bgkloctomap/bgkloctomap.h:53-61)
class BGKLOctoMap : public OctoMapT<LA3DM_BGKL> {
public:
    BGKLOctoMap(float resolution, unsigned short block_depth, float sf2, float ell, float free_thresh,
                float occupied_thresh, float var_thresh, float prior_A, float prior_B, int device = 0)
        : OctoMapT(make_bgk_params(resolution, block_depth, sf2, ell, free_thresh, occupied_thresh, var_thresh, prior_A,
                                   prior_B), device) {}
    BGKLOctoMap() : BGKLOctoMap(0.1f, 4, 1.0f, 1.0f, 0.3f, 0.7f, 1.0f, 1.0f, 1.0f) {}
};

/// la3dm::BGKLVOctoMap (include/bgklvoctomap/bgklvoctomap.h:52-62)
class BGKLVOctoMap : public OctoMapT<LA3DM_BGKLV> {
public:
    BGKLVOctoMap(float resolution, unsigned short block_depth, float sf2, float ell, float free_thresh,
                 float occupied_thresh, float var_thresh, float prior_A, float prior_B, bool original_size, float MIN_W,
                 int device = 0)
        : OctoMapT(lv_params(resolution, block_depth, sf2, ell, free_thresh, occupied_thresh, var_thresh, prior_A,
                             prior_B, original_size, MIN_W), device) {}
private:
    static la3dm_params lv_params(float resolution, unsigned short block_depth, float sf2, float ell, float free_thresh,
                                  float occupied_thresh, float var_thresh, float prior_A, float prior_B,
                                  bool original_size, float MIN_W) {
        la3dm_params p = make_bgk_params(resolution, block_depth, sf2, ell, free_thresh, occupied_thresh, var_thresh,
                                         prior_A, prior_B);
        p.original_size = original_size ? 1 : 0;
        p.min_W = MIN_W;
        return p;
    }
};

/// la3dm::GPOctoMap (include/gpoctomap/gpoctomap.h:50-52)
class GPOctoMap : public OctoMapT<LA3DM_GP> {
public:
    GPOctoMap(float resolution, unsigned short block_depth, float sf2, float ell, float noise, float l, float min_var,
              float max_var, float max_known_var, float free_thresh, float occupied_thresh, int device = 0)
        : OctoMapT(gp_params(resolution, block_depth, sf2, ell, noise, l, min_var, max_var, max_known_var, free_thresh,
                             occupied_thresh), device) {}
    GPOctoMap() : GPOctoMap(0.1f, 4, 1.0f, 1.0f, 0.01f, 100.f, 0.001f, 1000.f, 0.02f, 0.3f, 0.7f) {}
private:
    static la3dm_params gp_params(float resolution, unsigned short block_depth, float sf2, float ell, float noise,
                                  float l, float min_var, float max_var, float max_known_var, float free_thresh,
                                  float occupied_thresh) {
        la3dm_params p = la3dm_params();
        p.resolution = resolution; p.block_depth = block_depth; p.sf2 = sf2; p.ell = ell; p.noise = noise; p.l = l;
        p.min_var = min_var; p.max_var = max_var; p.max_known_var = max_known_var;
        p.free_thresh = free_thresh; p.occupied_thresh = occupied_thresh;
        return p;
    }
};

}  // namespace LA3DM_B200_NAMESPACE

#endif /* LA3DM_B200_OCTOMAP_H */
