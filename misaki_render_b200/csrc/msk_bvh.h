// GPU BVH build (msk_bvh.cu): host-callable entry points.
#pragma once
#include "msk_device.cuh"
#include <vector>

namespace msk {

struct BvhResult {
    float4  *nodes = nullptr;  // nnodes x 5 float4 (80-byte wide nodes), device
    float4  *tris  = nullptr;  // tri_slots x 3 float4: the padded per-node triangle blocks (msk_device.cuh), device
    uint64_t nnodes = 0, ntris = 0, tri_slots = 0;
    uint32_t depth = 0;        // levels of the wide tree
    int      builder = 0;      // MSK_BVH_PLOC / MSK_BVH_LBVH
    uint32_t build_rounds = 0; // PLOC: clustering rounds
    float    ms_build = 0.f;
    float    sah_cost = 0.f;   // sum of binary-node areas / root area
    float    lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
};

// d_verts: 2 float4 per vertex; d_indices: 3 mesh-local indices per triangle; both device.
enum { MSK_BVH_PLOC = 0, MSK_BVH_LBVH = 1 };
int  bvh_build(cudaStream_t stream, const float4 *d_verts, const uint32_t *d_indices, const std::vector<DMeshInfo> &meshes,
               BvhResult *out, int builder = MSK_BVH_LBVH);
void bvh_free(BvhResult *r);
void bvh_print_shape(cudaStream_t stream, const BvhResult &r); // MSK_DEBUG_SETUP: child / leaf histograms of the wide tree on stderr

} // namespace msk
