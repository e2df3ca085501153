// C entry points over further plugin sources the reference's build compiles, #included from where they lie:
//   src/librender/rfilter.cpp + filters/gaussian.cpp        the discretised reconstruction filter (33-entry table)
//   src/librender/sampler.cpp + samplers/independent.cpp    IndependentSampler::seed / next1d / next2d
//   src/librender/spectra/regular.cpp, uniform.cpp          the tabulated and the constant spectrum
// over the stand-ins under oracle/ref_shim/ (oracle/Makefile.ref).  TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "msk_ref_prelude.h"
#include <misaki/render/interaction.h>
#include <misaki/render/texture.h>
#include <rfilter.cpp>
#include <filters/gaussian.cpp>
#include <sampler.cpp>
#include <samplers/independent.cpp>
#include <spectra/regular.cpp>
#include <spectra/uniform.cpp>

using namespace misaki;

// for ref_spectra_wrap.cpp: regular.cpp defines a non-inline operator<<, so it can only be compiled into one translation unit
misaki::Object *msk_ref_make_regular(const misaki::Properties &p) { return new RegularSpectrum(p); }

extern "C" {

// out: radius, border_size, then the 33 table entries read back through eval_discretized (rfilter.h:13-16)
void ref_gaussian_filter(float stddev, float out[35]) {
    Properties p;
    p.floats["stddev"] = stddev;
    GaussianFilter f(p);
    out[0] = f.radius(); out[1] = (float) f.border_size();
    const float scale = float(MSK_FILTER_RESOLUTION) / f.radius();
    for (int i = 0; i <= MSK_FILTER_RESOLUTION; ++i) out[2 + i] = f.eval_discretized((float(i) + 0.5f) / scale);
}
// seed(seed_value) with Sampler "base_seed", then n1 x next1d followed by n2 x next2d (independent.cpp:20-35)
void ref_independent_sampler(uint64_t base_seed, uint64_t seed_value, int n1, int n2, float *out) {
    Properties p;
    p.ints["base_seed"] = (long long) base_seed;
    IndependentSampler s(p);
    s.seed(seed_value);
    for (int i = 0; i < n1; ++i) *out++ = s.next1d();
    for (int i = 0; i < n2; ++i) { Eigen::Vector2f v = s.next2d(); *out++ = v.x(); *out++ = v.y(); }
}
void ref_regular_spectrum(float lambda_min, float lambda_max, const float *values, size_t size, const float wl[4], float out[4], float *mean) {
    Properties p;
    p.floats["lambda_min"] = lambda_min; p.floats["lambda_max"] = lambda_max;
    p.ints["size"] = (long long) size; p.pointers["values"] = values;
    RegularSpectrum r(p);
    SceneInteraction si;
    si.wavelengths = Wavelength(wl[0], wl[1], wl[2], wl[3]);
    Spectrum v = r.eval(si);
    for (int i = 0; i < 4; ++i) out[i] = v.coeff(i);
    *mean = r.mean();
}
void ref_uniform_spectrum(float value, const float wl[4], float out[4]) {
    Properties p;
    p.floats["value"] = value;
    UniformSpectrum u(p);
    SceneInteraction si;
    si.wavelengths = Wavelength(wl[0], wl[1], wl[2], wl[3]);
    Spectrum v = u.eval(si);
    for (int i = 0; i < 4; ++i) out[i] = v.coeff(i);
}

} // extern "C"
