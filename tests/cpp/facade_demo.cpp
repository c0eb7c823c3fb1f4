// Test program for the C++ facade (include/la3dm_b200/octomap.h): reads like the reference's static node
// (src/bgkoctomap/bgkoctomap_static_node.cpp:86-139): construct, insert scans, walk the leaves.
// usage: facade_demo <scan file: int32 n_scans, int32 n_pts, float origins[n_scans][3], float pts[n_scans][n_pts][3]>
// prints: leaves free occupied unknown sum_prob bbox(6) probe_prob ray_steps ray_valid ray_bad
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <map>
#include <vector>

#include "la3dm_b200/octomap.h"

struct PointXYZ { float x, y, z, pad; };          // same layout as pcl::PointXYZ (16 bytes)
struct Cloud { std::vector<PointXYZ> points; };   // same member name as pcl::PointCloud

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int n_scans = 0, n_pts = 0;
    if (fread(&n_scans, 4, 1, f) != 1 || fread(&n_pts, 4, 1, f) != 1) return 2;
    std::vector<float> org(3 * n_scans), pts((size_t) 3 * n_scans * n_pts);
    if (fread(org.data(), 4, org.size(), f) != org.size() || fread(pts.data(), 4, pts.size(), f) != pts.size()) return 2;
    fclose(f);
    // config/methods/bgkoctomap.yaml
    la3dm::BGKOctoMap map(0.1f, 3, 1.0f, 0.2f, 0.3f, 0.7f, 100.0f, 0.001f, 0.001f);
    std::map<int64_t, std::vector<la3dm_leaf>> mirror;
    for (int s = 0; s < n_scans; ++s) {
        Cloud cloud;
        cloud.points.resize(n_pts);
        for (int i = 0; i < n_pts; ++i) {
            const float *p = &pts[((size_t) s * n_pts + i) * 3];
            cloud.points[i] = PointXYZ{p[0], p[1], p[2], 1.0f};
        }
        la3dm::vec3f origin(org[3 * s], org[3 * s + 1], org[3 * s + 2]);
        map.insert_pointcloud(cloud, origin, map.get_resolution(), 0.5f, 8.0f);   // static node: ds = resolution
        // the server loop's marker arrays (bgkoctomap_server.cpp:94-144) kept from the blocks this scan touched
        std::vector<la3dm_leaf> lv;
        std::vector<int64_t> bk;
        map.touched_leaves((1u << LA3DM_FREE) | (1u << LA3DM_OCCUPIED), lv, bk);
        for (int64_t k : bk) mirror.erase(k);
        for (const la3dm_leaf &l : lv) mirror[l.block_key].push_back(l);
    }
    long n = 0, nf = 0, no = 0, nu = 0;
    double sp = 0;
    for (auto it = map.begin_leaf(); it != map.end_leaf(); ++it) {
        ++n;
        const la3dm::OcTreeNode node = it.get_node();
        if (node.get_state() == la3dm::State::FREE) ++nf;
        else if (node.get_state() == la3dm::State::OCCUPIED) ++no;
        else ++nu;
        sp += node.get_prob();
    }
    // RayCaster like the (commented-out) use in src/bgkloctomap/bgkloctomap_static_node.cpp:117-129: walk a ray from the
    // first sensor origin; every valid step must be the node search() names for the step's own point
    long ray_steps = 0, ray_valid = 0, ray_bad = 0;
    {
        la3dm::vec3f start(org[0], org[1], org[2]), end(org[0] + 3.0f, org[1] + 1.0f, org[2] + 0.2f);
        la3dm::BGKOctoMap::RayCaster ray(&map, start, end);
        while (!ray.end()) {
            la3dm::vec3f p;
            la3dm::OcTreeNode node;
            la3dm::BlockHashKey bk;
            uint32_t nk;
            ++ray_steps;
            if (ray.next(p, node, bk, nk)) {
                ++ray_valid;
                la3dm_leaf l;
                const float q[3] = {p.x(), p.y(), p.z()};
                map.search(q, 1, sizeof(q), &l, true);
                if (l.block_key != bk || ((uint32_t) (l.depth << 16) + (uint32_t) l.index) != nk || l.a != node.get_a()) ++ray_bad;
            }
        }
    }
    // insert_training_data: three labelled points next to the first origin
    {
        std::vector<std::pair<la3dm::vec3f, float>> xy;
        xy.emplace_back(la3dm::vec3f(org[0] + 0.5f, org[1], org[2]), 1.0f);
        xy.emplace_back(la3dm::vec3f(org[0] + 0.3f, org[1], org[2]), 0.0f);
        xy.emplace_back(la3dm::vec3f(org[0] + 0.1f, org[1], org[2]), 0.0f);
        la3dm::BGKOctoMap side(0.1f, 3, 1.0f, 0.2f, 0.3f, 0.7f, 100.0f, 0.001f, 0.001f);
        side.insert_training_data(xy);
        if (side.num_blocks() == 0 || side.search(org[0] + 0.5f, org[1], org[2]).get_prob() <= 0.9f) ++ray_bad;
    }
    {   // the incremental mirror must hold exactly the FREE / OCCUPIED leaves of the full walk above
        long mf = 0, mo = 0;
        for (const auto &kv : mirror)
            for (const la3dm_leaf &l : kv.second) (l.state == LA3DM_FREE ? mf : mo) += 1;
        if (mf != nf || mo != no) ++ray_bad;
    }
    la3dm::vec3f mn, mx;
    map.get_bbox(mn, mx);
    const la3dm::OcTreeNode probe = map.search(5.05f, 0.15f, 1.35f);    // SURVEY.md 8c known answer after scan 1
    printf("%ld %ld %ld %ld %.4f %.9g %.9g %.9g %.9g %.9g %.9g %.7f %ld %ld %ld\n", n, nf, no, nu, sp, mn.x(), mn.y(), mn.z(),
           mx.x(), mx.y(), mx.z(), probe.get_prob(), ray_steps, ray_valid, ray_bad);
    return 0;
}
