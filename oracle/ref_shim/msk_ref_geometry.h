// The reference's own transform type (include/misaki/core/transform.h, compiled from where it lies against the Eigen stand-in's
// Affine3f / AngleAxisf / 4x4 inverse) and its bounding box / sphere types (include/misaki/core/{bbox,bsphere}.h).
// Shapes and meshes of the pinned build are given world-space data, i.e. the identity transform -- whose apply_point /
// apply_normal are exact -- and the perspective camera (ref_camera_wrap.cpp) gets real matrices through
// Properties::transform (the flat property set of ref_shim/misaki/core/properties.h).  TEST INFRASTRUCTURE.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <misaki/core/bbox.h>
#include <misaki/core/transform.h>
namespace misaki {
template <> inline Transform4f msk_ref_make_transform<Transform4f>(const float *m) { // row-major 4x4
    Eigen::Matrix4f M;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M(i, j) = m[4 * i + j];
    return Transform4f(M);
}
} // namespace misaki
