// Test program for the C++ facade (include/la3dm_b200/octomap.h): reads like the reference's static node
// (src/bgkoctomap/bgkoctomap_static_node.cpp:86-139): construct, insert scans, walk the leaves.
// usage: facade_demo <scan file: int32 n_scans, int32 n_pts, float origins[n_scans][3], float pts[n_scans][n_pts][3]>
// prints: leaves free occupied unknown sum_prob bbox(6)
#include <cstdio>
#include <cstdlib>
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
    for (int s = 0; s < n_scans; ++s) {
        Cloud cloud;
        cloud.points.resize(n_pts);
        for (int i = 0; i < n_pts; ++i) {
            const float *p = &pts[((size_t) s * n_pts + i) * 3];
            cloud.points[i] = PointXYZ{p[0], p[1], p[2], 1.0f};
        }
        la3dm::vec3f origin(org[3 * s], org[3 * s + 1], org[3 * s + 2]);
        map.insert_pointcloud(cloud, origin, map.get_resolution(), 0.5f, 8.0f);   // static node: ds = resolution
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
    la3dm::vec3f mn, mx;
    map.get_bbox(mn, mx);
    const la3dm::OcTreeNode probe = map.search(5.05f, 0.15f, 1.35f);    // SURVEY.md 8c known answer after scan 1
    printf("%ld %ld %ld %ld %.4f %.9g %.9g %.9g %.9g %.9g %.9g %.7f\n", n, nf, no, nu, sp, mn.x(), mn.y(), mn.z(), mx.x(),
           mx.y(), mx.z(), probe.get_prob());
    return 0;
}
