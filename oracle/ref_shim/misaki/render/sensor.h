// Stand-in for include/misaki/render/sensor.h (the real one pulls transform.h): the members SamplingIntegrator::render and
// Scene use, with the same signatures.  The camera itself (perspective.cpp) is not part of the pinned build: Sensor forwards
// sample_ray_differential to a callback that supplies the ray (an INPUT, as in the single-path comparisons).  The film IS the
// reference's own (film.h / film.cpp / films/hdrfilm.cpp, ref_hdrfilm_wrap.cpp).
// TEST INFRASTRUCTURE.
#pragma once
#if defined(MSK_REF_REAL_SENSOR)
// ref_camera_wrap.cpp: the reference's own include/misaki/render/sensor.h (gen copy, oracle/Makefile.ref), with the class
// renamed by that translation unit so that it does not collide with the stand-in the render-loop wrappers use
#include <misaki/render/sensor_real.h>
#else
#include "msk_ref_prelude.h"
#include <misaki/core/object.h>
#include <misaki/core/ray.h>
#include <misaki/render/film.h>
#include <misaki/render/imageblock.h>
#include <misaki/render/sampler.h>
#include <string>
#include <vector>
namespace misaki {
class Sensor : public Object {
public:
    // out: o[3] d[3] mint maxt | wavelengths[4] | ray_weight[4]
    typedef void (*RayCallback)(float wavelength_sample, float px, float py, float *out16);
    Sensor(Film *film, Sampler *sampler, RayCallback cb) : m_film(film), m_sampler(sampler), m_cb(cb) {}
    Film *film() { return m_film; }
    const Film *film() const { return m_film; }
    Sampler *sampler() { return m_sampler; }
    const Sampler *sampler() const { return m_sampler; }
    const Medium *medium() const { return nullptr; }
    std::pair<RayDifferential, Spectrum> sample_ray_differential(const float wavelength_sample, const Eigen::Vector2f &sample2,
                                                                 const Eigen::Vector2f & /*aperture sample: consumed, unused*/) const {
        float r[16];
        m_cb(wavelength_sample, sample2.x(), sample2.y(), r);
        Ray ray(Eigen::Vector3f(r[0], r[1], r[2]), Eigen::Vector3f(r[3], r[4], r[5]), r[6], r[7], 0.f, Wavelength(r[8], r[9], r[10], r[11]));
        RayDifferential rd(ray);
        rd.wavelengths = ray.wavelengths;
        rd.o_x = rd.o_y = ray.o; rd.d_x = rd.d_y = ray.d; // no compiled BSDF consumes differentials (bsdf.cpp:18)
        return { rd, Spectrum(r[12], r[13], r[14], r[15]) };
    }
private:
    Film *m_film;
    Sampler *m_sampler;
    RayCallback m_cb;
};
} // namespace misaki
#endif
