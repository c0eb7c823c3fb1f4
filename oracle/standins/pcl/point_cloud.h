// TEST INFRASTRUCTURE ONLY -- stand-in for <pcl/point_cloud.h>: the subset of pcl::PointCloud<T> the reference map
// classes use (points/width/height/is_dense, begin/end/size/push_back/clear, Ptr).  See point_types.h.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    typedef typename std::vector<PointT>::iterator iterator;
    typedef typename std::vector<PointT>::const_iterator const_iterator;

    std::vector<PointT> points;
    std::uint32_t width = 0;
    std::uint32_t height = 0;
    bool is_dense = true;

    iterator begin() { return points.begin(); }
    iterator end() { return points.end(); }
    const_iterator begin() const { return points.begin(); }
    const_iterator end() const { return points.end(); }
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = 0; height = 0; }
    // pcl::PointCloud::push_back: append and make the cloud unorganised (width = size, height = 1)
    void push_back(const PointT &p) {
        points.push_back(p);
        width = static_cast<std::uint32_t>(points.size());
        height = 1;
    }
    PointT &operator[](size_t i) { return points[i]; }
    const PointT &operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl
