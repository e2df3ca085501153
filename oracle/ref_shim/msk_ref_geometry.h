// Stand-in for the one geometry value type Shape / Mesh / Emitter hold but the pinned code paths never compute with:
// include/misaki/core/transform.h needs Eigen::Affine3f, AngleAxisf and 4x4 inverses.  The bounding box / sphere types
// are the reference's own (include/misaki/core/{bbox,bsphere}.h).  TEST INFRASTRUCTURE.
#pragma once
#include <Eigen/Core>
#include <misaki/core/bbox.h>
namespace misaki {
struct Transform4f { // identity only: the pinned loaders and shapes are given world-space data
    Transform4f() {}
    Eigen::Vector3f apply_point(const Eigen::Vector3f &p) const { return p; }
    Eigen::Vector3f apply_normal(const Eigen::Vector3f &n) const { return n; }
    Eigen::Vector3f apply_vector(const Eigen::Vector3f &v) const { return v; }
};
} // namespace misaki
