// Wavefront renderer + batch ray queries (msk_render.cu): host-callable interface.
#pragma once
#include "msk_device.cuh"

namespace msk {

class Renderer {
public:
    Renderer();
    ~Renderer();
    int  init(int sm_count);
    void release();
    // d_film: device, height x width x 5 floats (X,Y,Z,A,W); with `aov` (the AOV integrator, aov.cpp) the pixel
    // stride is 5 + nch and the AOV channels follow XYZAW
    int render(cudaStream_t stream, const DScene &sc, const MskRenderDesc &rd, float *d_film, MskStats *stats,
               const MskAovDesc *aov = nullptr);
    int reserve(const DScene &sc, const MskRenderDesc &rd);    // allocate the path pool of such a render now
    static int aov_plan(const MskAovDesc &aov, uint32_t *nch); // validates, counts channels
    int intersect(cudaStream_t stream, const DScene &sc, const MskRay *d_rays, MskHit *d_hits, size_t n);
    int intersect_stats(cudaStream_t stream, const DScene &sc, const MskRay *d_rays, size_t n, uint32_t *d_nodes, uint32_t *d_tris);
    int occluded(cudaStream_t stream, const DScene &sc, const MskRay *d_rays, uint8_t *d_occ, size_t n);

private:
    int ensure_pool(uint32_t capacity, int count);
    struct Impl;
    Impl *impl_;
};

} // namespace msk
