// Stand-ins for the geometry value types Shape / Mesh hold but the pinned code paths never compute with
// (include/misaki/core/{transform,bbox}.h need Eigen::Affine3f, 4x4 inverses, ...).  TEST INFRASTRUCTURE.
#pragma once
#include <Eigen/Core>
namespace misaki {
struct Transform4f { Transform4f() {} };
struct BoundingBox3f {
    Eigen::Vector3f pmin, pmax;
    BoundingBox3f() { reset(); }
    BoundingBox3f(const Eigen::Vector3f &a, const Eigen::Vector3f &b) : pmin(a), pmax(b) {}
    void reset() { pmin = Eigen::Vector3f::Constant(1e30f); pmax = Eigen::Vector3f::Constant(-1e30f); }
    void expand(const Eigen::Vector3f &p) { pmin = pmin.cwiseMin(p); pmax = pmax.cwiseMax(p); }
    void expand(const BoundingBox3f &b) { pmin = pmin.cwiseMin(b.pmin); pmax = pmax.cwiseMax(b.pmax); }
};
} // namespace misaki
