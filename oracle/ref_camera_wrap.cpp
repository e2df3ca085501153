// C entry points over the reference's OWN camera: include/misaki/core/transform.h (Transform4f: scale / translate /
// perspective / lookat, products that carry their inverses), src/librender/sensor.cpp (Sensor, ProjectiveCamera:
// near / far clip, aspect from the film) and src/librender/sensors/perspective.cpp (camera_to_sample, sample_ray),
// #included from where they lie.  Eigen::Affine3f, AngleAxisf and the 4x4 inverse come from the stand-in
// (ref_shim/Eigen/Geometry, ref_shim/Eigen/Core: cofactor inverse -- Eigen's packed SSE routine orders the float operations
// differently, so the pin is to 1e-6 relative, not bit equality).  The film and sampler children are the reference's own
// HDRFilm / IndependentSampler (ref_hdrfilm_wrap.cpp, ref_plugins_wrap.cpp).
// The render-loop wrappers keep their Sensor stand-in (a callback); the real class is renamed in this translation unit so
// that the two do not collide, and ref_camera_callback() is such a callback backed by the real camera.
// TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "msk_ref_prelude.h"
#define MSK_REF_REAL_SENSOR
#define Sensor MskRefRealSensor
#include <misaki/core/logger.h>
#include <misaki/core/manager.h>
#include <misaki/core/properties.h>
#include <misaki/render/sensor.h>
#include <misaki/render/medium.h>
#include <sensor.cpp>
#include <sensors/perspective.cpp>
#include <samplers/independent.cpp> // class IndependentSampler (in-class members only)

using namespace misaki;
misaki::Film *msk_ref_make_hdrfilm(int W, int H, const misaki::ReconstructionFilter *filter); // ref_hdrfilm_wrap.cpp

namespace {
PerspectiveCamera *g_bound = nullptr;
void bound_callback(float ws, float px, float py, float *out16) {
    auto [ray, weight] = g_bound->sample_ray(ws, Eigen::Vector2f(px, py), Eigen::Vector2f(0.f, 0.f));
    for (int i = 0; i < 3; ++i) { out16[i] = ray.o[i]; out16[3 + i] = ray.d[i]; }
    out16[6] = ray.mint; out16[7] = ray.maxt;
    for (int i = 0; i < 4; ++i) { out16[8 + i] = ray.wavelengths[i]; out16[12 + i] = weight[i]; }
}
} // namespace

// to_world: row-major 4x4 camera-to-world, or null for the identity.  fov in degrees (perspective.cpp:11).
extern "C" void *ref_camera_create(int W, int H, float fov, float near_clip, float far_clip, const float *to_world) {
    try {
        Properties props("perspective");
        props.floats["fov"] = fov; props.floats["near_clip"] = near_clip; props.floats["far_clip"] = far_clip;
        if (to_world) { std::array<float, 16> m; for (int i = 0; i < 16; ++i) m[i] = to_world[i]; props.matrices["to_world"] = m; }
        Properties sp;
        sp.ints["sample_count"] = 1;
        props.children.push_back({ "film", ref<Object>(msk_ref_make_hdrfilm(W, H, nullptr)) });
        props.children.push_back({ "sampler", ref<Object>(new IndependentSampler(sp)) });
        return new PerspectiveCamera(props); // (the stand-in's ref<T> never frees; ref_camera_destroy deletes)
    } catch (...) { return nullptr; }
}
extern "C" void ref_camera_destroy(void *h) { if (g_bound == h) g_bound = nullptr; delete (PerspectiveCamera *) h; }

// samples: n x 3 (wavelength sample, pixel-unit x, pixel-unit y); out: n x 16 = o[3] d[3] mint maxt | wavelengths[4] | weight[4]
extern "C" int ref_camera_sample_rays(void *h, const float *samples, size_t n, float *out) {
    try {
        const PerspectiveCamera *cam = (const PerspectiveCamera *) h;
        for (size_t k = 0; k < n; ++k) {
            auto [ray, weight] = cam->sample_ray(samples[3 * k], Eigen::Vector2f(samples[3 * k + 1], samples[3 * k + 2]), Eigen::Vector2f(0.f, 0.f));
            float *o = out + 16 * k;
            for (int i = 0; i < 3; ++i) { o[i] = ray.o[i]; o[3 + i] = ray.d[i]; }
            o[6] = ray.mint; o[7] = ray.maxt;
            for (int i = 0; i < 4; ++i) { o[8 + i] = ray.wavelengths[i]; o[12 + i] = weight[i]; }
        }
        return 0;
    } catch (...) { return -2; }
}

// The matrices perspective.cpp:12-19 builds: camera_to_sample and its carried inverse (row-major), for film W x H.
extern "C" int ref_camera_matrices(int W, int H, float fov, float near_clip, float far_clip, float *camera_to_sample, float *sample_to_camera) {
    try {
        const float aspect = W / (float) H; // sensor.cpp:44
        Transform4f c2s = Transform4f::scale(Eigen::Vector3f((float) W, (float) H, 1.f)) * Transform4f::scale(Eigen::Vector3f(-0.5f, -0.5f * aspect, 1.f)) *
                          Transform4f::translate(Eigen::Vector3f(-1.f, -1.f / aspect, 0.f)) * Transform4f::perspective(fov, near_clip, far_clip);
        Transform4f s2c = c2s.inverse();
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { camera_to_sample[4 * i + j] = c2s.matrix()(i, j); sample_to_camera[4 * i + j] = s2c.matrix()(i, j); }
        return 0;
    } catch (...) { return -2; }
}

// Transform4f::lookat (transform.h:189-199), row-major
extern "C" void ref_transform_lookat(const float *origin, const float *target, const float *up, float *out16) {
    Transform4f t = Transform4f::lookat(Eigen::Vector3f(origin[0], origin[1], origin[2]), Eigen::Vector3f(target[0], target[1], target[2]),
                                        Eigen::Vector3f(up[0], up[1], up[2]));
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out16[4 * i + j] = t.matrix()(i, j);
}
// scale / translate / rotate (axis, angle as passed: the XML loader converts degrees) composed as the loader composes them:
// kind 0 translate, 1 scale, 2 rotate; result = T_kind(v[, angle]) (row-major), and its carried inverse
extern "C" void ref_transform_make(int kind, const float *v, float angle, float *out16, float *inv16) {
    Eigen::Vector3f a(v[0], v[1], v[2]);
    Transform4f t = kind == 0 ? Transform4f::translate(a) : (kind == 1 ? Transform4f::scale(a) : Transform4f::rotate(a, angle));
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { out16[4 * i + j] = t.matrix()(i, j); inv16[4 * i + j] = t.inverse_matrix()(i, j); }
}

// a Sensor::RayCallback for ref_render / ref_render_aov (ref_render_wrap.cpp) backed by this camera: the render loop then
// runs on the reference's own camera (one camera bound at a time; the loop's stand-in for tbb is serial or joins its threads)
extern "C" void *ref_camera_callback(void *h) { g_bound = (PerspectiveCamera *) h; return (void *) &bound_callback; }
