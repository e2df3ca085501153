// C entry point over the reference's OWN SamplingIntegrator::render / render_block / render_sample
// (src/librender/integrator.cpp, with utils.cpp), #included from where it lies: the tile loop over BlockGenerator's
// spiral, the order of random draws per camera sample (position 2, wavelength 1, aperture 2), spectrum_to_xyz, the
// XYZAW channels, ImageBlock::put and HDRFilm::put (films/hdrfilm.cpp) -- driving the reference's own PathTracer / Scene of ref_path_wrap.cpp.
// tbb::parallel_for is a serial stand-in (ONE task: the sampler is cloned once, integrator.cpp:57, and never re-seeded --
// SURVEY F6); the camera ray of a sample is supplied by a callback (perspective.cpp is not part of the pinned build).
// TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "ref_path_scene.h"
#include <misaki/render/sensor.h>
#include <filters/gaussian.cpp>     // class GaussianFilter (in-class members only)
#include <samplers/independent.cpp> // class IndependentSampler (ditto)
#include <utils.cpp>
#include <integrator.cpp>

using namespace misaki;
misaki::SamplingIntegrator *msk_ref_path_tracer(RefPathScene *s);
misaki::Scene *msk_ref_scene(RefPathScene *s);
misaki::Film *msk_ref_make_hdrfilm(int W, int H, const misaki::ReconstructionFilter *filter); // ref_hdrfilm_wrap.cpp
const float *msk_ref_hdrfilm_storage(misaki::Film *film);

misaki::SamplingIntegrator *msk_ref_make_aov(RefPathScene *s); // ref_path_wrap.cpp

static int render_with(SamplingIntegrator *integrator, RefPathScene *s, int W, int H, int spp, float stddev, Sensor::RayCallback cb, int nch,
                       float *film_out) {
    try {
        Properties fp;
        fp.floats["stddev"] = stddev;
        GaussianFilter *filter = new GaussianFilter(fp);
        Film *film = msk_ref_make_hdrfilm(W, H, filter); // the reference's own HDRFilm: prepare / put under its mutex
        Properties sp;
        sp.ints["sample_count"] = spp;
        IndependentSampler *sampler = new IndependentSampler(sp);
        Sensor *sensor = new Sensor(film, sampler, cb);
        if (!integrator->render(msk_ref_scene(s), sensor)) return -1;
        memcpy(film_out, msk_ref_hdrfilm_storage(film), sizeof(float) * (size_t) W * H * nch);
        return 0;
    } catch (...) { return -2; }
}
// film_out: H x W x 5 (X, Y, Z, A, W).  block_size: SamplingIntegrator "block_size" (default 32) is fixed when the tracer is
// constructed, so the scene's tracer is used as is.
extern "C" int ref_render(void *handle, int W, int H, int spp, float stddev, Sensor::RayCallback cb, float *film_out) {
    RefPathScene *s = (RefPathScene *) handle;
    return render_with(msk_ref_path_tracer(s), s, W, H, spp, stddev, cb, 5, film_out);
}
// The same loop driving AOVIntegrator (integrators/aov.cpp: depth, position, uv, geometric / shading normal, nested path
// tracer RGBA): film_out is H x W x 21 -- X, Y, Z, A, W and the 16 AOV channels in aov_names() order (integrator.cpp:36-41).
extern "C" int ref_render_aov(void *handle, int W, int H, int spp, float stddev, Sensor::RayCallback cb, float *film_out) {
    RefPathScene *s = (RefPathScene *) handle;
    try {
        return render_with(msk_ref_make_aov(s), s, W, H, spp, stddev, cb, 21, film_out);
    } catch (...) { return -2; }
}
