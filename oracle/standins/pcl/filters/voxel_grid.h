// TEST INFRASTRUCTURE ONLY -- stand-in for <pcl/filters/voxel_grid.h>.
//
// PCL is a third-party dependency of la3dm that is NOT vendored under /root/reference and is not installed here
// (package.xml only names `pcl_ros`; version unpinned -> "parity unpinned" at this boundary).
// This header restates the published algorithm of pcl::VoxelGrid<PointT>::applyFilter (PCL 1.8-1.12,
// filters/include/pcl/filters/impl/voxel_grid.hpp) for the default configuration the reference uses
// (call sites: src/bgkoctomap/bgkoctomap.cpp:427-430, src/gpoctomap/gpoctomap.cpp:411-414,
//  src/bgkloctomap/bgkloctomap.cpp:354-357, src/bgklvoctomap/bgklvoctomap.cpp:433-436):
//   * no filter field, min_points_per_voxel = 0, downsample_all_data = true (CentroidPoint accumulator),
//   * getMinMax3D over all points (non-finite points skipped only when !is_dense),
//   * inverse_leaf_size = 1/leaf (fp32), min_b/max_b = floor(min*inv), div_b = max_b-min_b+1,
//   * int64 overflow guard: dx*dy*dz > INT32_MAX  =>  output = input (PCL prints a warning and copies),
//   * idx = ijk0 + ijk1*div0 + ijk2*div0*div1 with ijk = int(floor(p*inv) - float(min_b)),
//   * sort by idx, one output point per idx run in ascending idx order,
//   * centroid = (sequential fp32 sum of the run) / float(count).
// PCL sorts with std::sort (unstable); the order of equal-idx points is therefore unspecified in the real library.
// We fix it to ascending input index (stable sort) so that the oracle is deterministic.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include "../point_cloud.h"

namespace pcl {
template <typename PointT>
class VoxelGrid {
public:
    typedef typename PointCloud<PointT>::Ptr PointCloudPtr;
    void setInputCloud(const PointCloudPtr &cloud) { input_ = cloud; }
    void setLeafSize(float lx, float ly, float lz) {
        leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz;
        for (int i = 0; i < 3; ++i) inv_[i] = 1.0f / leaf_[i];
    }
    void filter(PointCloud<PointT> &output) {
        output.points.clear();
        output.height = 1;
        output.is_dense = true;
        output.width = 0;
        if (!input_ || input_->points.empty()) return;
        const std::vector<PointT> &in = input_->points;
        const bool dense = input_->is_dense;

        float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(),
                       std::numeric_limits<float>::max()};
        float mx[3] = {-mn[0], -mn[1], -mn[2]};
        for (size_t k = 0; k < in.size(); ++k) {
            const float v[3] = {in[k].x, in[k].y, in[k].z};
            if (!dense && (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2]))) continue;
            for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], v[i]); mx[i] = std::max(mx[i], v[i]); }
        }
        const std::int64_t dx = static_cast<std::int64_t>((mx[0] - mn[0]) * inv_[0]) + 1;
        const std::int64_t dy = static_cast<std::int64_t>((mx[1] - mn[1]) * inv_[1]) + 1;
        const std::int64_t dz = static_cast<std::int64_t>((mx[2] - mn[2]) * inv_[2]) + 1;
        if (dx * dy * dz > static_cast<std::int64_t>(std::numeric_limits<std::int32_t>::max())) {
            output = *input_;   // "Leaf size is too small for the input dataset. Integer indices would overflow."
            return;
        }
        int min_b[3], max_b[3], div_b[3];
        for (int i = 0; i < 3; ++i) {
            min_b[i] = static_cast<int>(std::floor(mn[i] * inv_[i]));
            max_b[i] = static_cast<int>(std::floor(mx[i] * inv_[i]));
            div_b[i] = max_b[i] - min_b[i] + 1;
        }
        const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};

        struct Entry { unsigned int idx; unsigned int src; };
        std::vector<Entry> iv;
        iv.reserve(in.size());
        for (size_t k = 0; k < in.size(); ++k) {
            const float v[3] = {in[k].x, in[k].y, in[k].z};
            if (!dense && (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2]))) continue;
            const int i0 = static_cast<int>(std::floor(v[0] * inv_[0]) - static_cast<float>(min_b[0]));
            const int i1 = static_cast<int>(std::floor(v[1] * inv_[1]) - static_cast<float>(min_b[1]));
            const int i2 = static_cast<int>(std::floor(v[2] * inv_[2]) - static_cast<float>(min_b[2]));
            const int idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
            iv.push_back(Entry{static_cast<unsigned int>(idx), static_cast<unsigned int>(k)});
        }
        std::stable_sort(iv.begin(), iv.end(), [](const Entry &a, const Entry &b) { return a.idx < b.idx; });

        size_t first = 0;
        while (first < iv.size()) {
            size_t last = first + 1;
            while (last < iv.size() && iv[last].idx == iv[first].idx) ++last;
            float s[3] = {0.f, 0.f, 0.f};
            for (size_t li = first; li < last; ++li) {
                const PointT &p = in[iv[li].src];
                s[0] += p.x; s[1] += p.y; s[2] += p.z;
            }
            const float n = static_cast<float>(last - first);
            output.points.push_back(PointT(s[0] / n, s[1] / n, s[2] / n));
            first = last;
        }
        output.width = static_cast<std::uint32_t>(output.points.size());
    }

private:
    PointCloudPtr input_;
    float leaf_[3] = {0.f, 0.f, 0.f};
    float inv_[3] = {0.f, 0.f, 0.f};
};
}  // namespace pcl
