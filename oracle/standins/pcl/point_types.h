// TEST INFRASTRUCTURE ONLY -- stand-in for <pcl/point_types.h> (PCL is not installed in this image and is not
// part of /root/reference).  Only what la3dm's map classes touch: pcl::PointXYZ with public x,y,z and a 16-byte
// footprint like the real (SSE-padded) type.  Used solely to compile the UNMODIFIED reference sources into
// oracle/_ref (see oracle/Makefile).  Never linked into the product library.
#pragma once
namespace pcl {
struct PointXYZ {
    float x, y, z, _pad;
    PointXYZ() : x(0.f), y(0.f), z(0.f), _pad(1.f) {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), _pad(1.f) {}
};
}  // namespace pcl
