// ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for <pcl/point_types.h> so that the reference's
// include/ikd-Tree/ikd_Tree.{h,cpp} compile unmodified in a container that has
// no PCL and no Eigen.  Only the three point structs that ikd_Tree.cpp
// instantiates (ikd_Tree.cpp:1725-1727) and the Eigen::aligned_allocator name
// used by KD_TREE::PointVector (ikd_Tree.h:55) are provided.
//
// Layout follows PCL's documented layout: 16-byte aligned, xyz + pad,
// normal_xyz + pad, intensity, curvature + pad (48 bytes for PointXYZINormal).
#pragma once
#include <memory>
#include <vector>

namespace pcl {

struct alignas(16) PointXYZ {
    float x = 0.f, y = 0.f, z = 0.f, _pad0 = 1.f;
};

struct alignas(16) PointXYZI {
    float x = 0.f, y = 0.f, z = 0.f, _pad0 = 1.f;
    float intensity = 0.f, _pad1[3] = {0.f, 0.f, 0.f};
};

struct alignas(16) PointXYZINormal {
    float x = 0.f, y = 0.f, z = 0.f, _pad0 = 1.f;
    float normal_x = 0.f, normal_y = 0.f, normal_z = 0.f, _pad1 = 0.f;
    float intensity = 0.f, curvature = 0.f, _pad2[2] = {0.f, 0.f};
};

}  // namespace pcl

namespace Eigen {
template <class T>
using aligned_allocator = std::allocator<T>;
}
