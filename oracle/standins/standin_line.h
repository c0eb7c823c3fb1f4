// TEST INFRASTRUCTURE ONLY -- point-to-segment distance shared by the -L and -LV stand-ins.
// Restates point_to_line_dist (include/bgkloctomap/bgklinference.h:106-141, identical copy at
// include/bgklvoctomap/bgklvinference.h:98-135) using the reference's own point3f arithmetic
// (include/common/point3f.h: float component math, dot()/norm() evaluated in float then widened to double).
#pragma once
#include "point3f.h"
namespace la3dm_standin {
static inline float point_to_segment(const la3dm::point3f &p, const float *seg) {
    const float EPSILON = 0.0001f;
    la3dm::point3f p0(seg[0], seg[1], seg[2]);
    la3dm::point3f p1(seg[3], seg[4], seg[5]);
    la3dm::point3f line_vec = p1 - p0;
    float line_len = line_vec.norm();
    la3dm::point3f pnt_vec = p - p0;
    if (line_len < EPSILON) return (float) (p - p0).norm();
    double c1 = pnt_vec.dot(line_vec);
    double c2 = line_vec.dot(line_vec);
    if (c1 <= 0) return (float) (p - p0).norm();
    if (c2 <= c1) return (float) (p - p1).norm();
    double b = c1 / c2;
    la3dm::point3f nearest = p0 + (line_vec * b);
    return (float) (p - nearest).norm();
}
}  // namespace la3dm_standin
