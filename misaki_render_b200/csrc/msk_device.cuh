// Device-side data layout shared by the BVH builder, the traversal kernels and the
// wavefront path tracer.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/misaki_b200.h"

#define MSK_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) return msk::cuda_fail(err__, #expr, __FILE__, __LINE__);     \
    } while (0)

namespace msk {

int cuda_fail(cudaError_t err, const char *expr, const char *file, int line);
int fail(int code, const char *fmt, ...);

// ---- wide BVH node: 80 bytes = five 16-byte loads (quantised boxes as in Ylitie, Karras, Laine 2017) ----
//  n0: p.x p.y p.z | ex ey ez imask          origin of the quantisation grid, exponents, internal-child mask
//  n1: child_base | tri_base | valid | -
//  n2: qlo_x[0..7]            | qlo_y[0..7]
//  n3: qlo_z[0..7]            | qhi_x[0..7]
//  n4: qhi_y[0..7]            | qhi_z[0..7]
// Child slot j (0..7) is either empty, an inner node -- bit j of imask; its node is child_base + the number of inner
// slots below j -- or a leaf of 1..kMaxLeafTris triangles stored at triangle slots tri_base + 3j + k.  `valid` =
// imask << 24 | bit 3j + k for every triangle present, so a traversal step ORs the STATIC word
// 1 << (24 + j) | 7 << 3j for every child box the ray enters and masks with `valid` once.  Triangle storage is
// therefore padded: a node owns the slot range [3 jmin, 3 jmax + 3) of its leaf slots (tri_base is biased by
// -3 jmin, modulo 2^32); the builder packs the leaves of a node into the tightest run of free slots.
constexpr int kNodeFloat4s = 5;
constexpr int kTriFloat4s  = 3; // v0.xyz, prim | v1.xyz, geom | v2.xyz, -
#ifndef MSK_MAX_LEAF_TRIS
#define MSK_MAX_LEAF_TRIS 3 /* 1..3: three static triangle bits per child slot */
#endif
constexpr int kMaxLeafTris = MSK_MAX_LEAF_TRIS;

struct DMeshInfo {
    uint32_t vert_offset; // first vertex in DScene::verts (units of vertices)
    uint32_t tri_offset;  // first triangle in DScene::indices (units of triangles)
    uint32_t ntris;
    int32_t  bsdf;
    int32_t  emitter;
    uint32_t flags;       // 1 = vertex normals, 2 = texcoords; bits 8-15 / 16-23: interior / exterior medium + 1
    float    inv_area;    // 1 / Mesh::m_surface_area (float, sequential sum as mesh.cpp:39-48)
    uint32_t cdf_offset;  // emitter meshes: first entry of the (ntris+1)-entry area CDF in DScene::cdfs
};

struct DSpectrum { // MSK_SPEC_CHECKERBOARD reuses the fields: see texture_resolve (msk_shading.cuh)
    int32_t  kind;
    float    c0, c1, c2;
    float    value;
    uint32_t table_offset, table_size;
    float    lambda_min, inv_interval;
};

struct DCamera {
    float s2c[16];
    float c2w[16];
    float near_clip, far_clip;
    uint32_t width, height;
    float filter_radius;
    float filter_scale; // MSK_FILTER_RESOLUTION / radius
};

struct DScene {
    const float4    *nodes;
    const float4    *tris;
    const float4    *verts;   // 2 float4 per vertex: px py pz nx | ny nz u v
    const uint32_t  *indices; // 3 per triangle, mesh-local vertex ids
    const DMeshInfo *meshes;
    const MskBsdf   *bsdfs;
    const MskEmitter *emitters;
    const DSpectrum *spectra;
    const MskMedium *media;   // media/homogeneous.cpp (volpath only)
    const float     *tables;
    const float     *cdfs;
    const float     *filter_table; // 33 entries
    const float4    *cie;          // 95 rows: xbar ybar zbar d65
    uint32_t nemitters;
    int32_t  environment;
    float    env_radius;
    uint32_t nmeshes;
    uint32_t bsdf_type_mask; // bit t: some mesh uses a BSDF of MskBsdfType t
    int32_t  sensor_medium;  // medium the camera sits in, or -1 (sensor.cpp:12-18)
    uint32_t has_textures;   // some spectrum is uv-dependent (MSK_SPEC_CHECKERBOARD): resolve ids per surface point
    uint32_t nnodes;         // wide BVH nodes
    float    bb_lo[3], bb_scale[3]; // scene bounds: (p - bb_lo) * bb_scale in [0, 1) -- ray reordering keys (msk_render.cu: k_ray_keys)
    uint32_t k47;            // 0x47000000: the byte -> float bias of the node decode, opaque to ptxas (msk_traverse.cuh: qfloat)
    DCamera  cam;
};

} // namespace msk
