// C entry points over the reference's OWN math headers, compiled from where they lie under /root/reference against a
// minimal Eigen stand-in (oracle/ref_shim/) by oracle/Makefile.ref into oracle/_ref/libmisaki_ref_math.so.
// TEST INFRASTRUCTURE: tools/gen_golden_ref_math.py calls these to produce tests/golden/ref_math.json, which pins the
// oracle's restatements (tests/test_oracle_ref_math.py).  Only calls and type conversions live here.
#include "msk_ref_prelude.h"
#include <misaki/core/warp.h>       // the copies under oracle/_ref/gen with the fwd.h include replaced (Makefile.ref)
#include <misaki/render/fresnel.h>
#include <misaki/render/microfacet.h>
#include <misaki/render/srgb.h>

using namespace misaki;
using Eigen::Vector2f;
using Eigen::Vector3f;

extern "C" {

void ref_pcg32_uints(uint64_t initstate, uint64_t initseq, uint32_t *out, size_t n) { // mathutils.h:89-143
    math::PCG32 rng(initstate, initseq);
    for (size_t i = 0; i < n; ++i) out[i] = rng.next_uint32();
}
void ref_pcg32_floats(uint64_t initstate, uint64_t initseq, float *out, size_t n) {
    math::PCG32 rng(initstate, initseq);
    for (size_t i = 0; i < n; ++i) out[i] = rng.next_float32();
}
// which: 0 uniform_triangle 1 disk_concentric 2 cosine_hemisphere 3 uniform_sphere (warp.h:11-53)
void ref_warp(int which, float u, float v, float out[3]) {
    Vector2f s(u, v);
    out[0] = out[1] = out[2] = 0.f;
    if (which == 0) { Vector2f r = warp::square_to_uniform_triangle(s); out[0] = r.x(); out[1] = r.y(); }
    else if (which == 1) { Vector2f r = warp::square_to_uniform_disk_concentric(s); out[0] = r.x(); out[1] = r.y(); }
    else if (which == 2) { Vector3f r = warp::square_to_cosine_hemisphere(s); out[0] = r.x(); out[1] = r.y(); out[2] = r.z(); }
    else { Vector3f r = warp::square_to_uniform_sphere(s); out[0] = r.x(); out[1] = r.y(); out[2] = r.z(); }
}
void ref_coordinate_system(const float n[3], float s[3], float t[3]) { // mathutils.h:196-203, frame.h:16-18
    Frame f(Vector3f(n[0], n[1], n[2]));
    for (int i = 0; i < 3; ++i) { s[i] = f.s[i]; t[i] = f.t[i]; }
}
void ref_frame_roundtrip(const float n[3], const float v[3], float local[3], float world[3]) { // frame.h:20-26
    Frame f(Vector3f(n[0], n[1], n[2]));
    Vector3f l = f.to_local(Vector3f(v[0], v[1], v[2])), w = f.to_world(Vector3f(v[0], v[1], v[2]));
    for (int i = 0; i < 3; ++i) { local[i] = l[i]; world[i] = w[i]; }
}
void ref_fresnel(float cos_theta_i, float eta, float out[4]) { // fresnel.h:37-63: F, cos_theta_t, eta_it, eta_ti
    auto [r, ct, it, ti] = fresnel<float>(cos_theta_i, eta);
    out[0] = r; out[1] = ct; out[2] = it; out[3] = ti;
}
void ref_fresnel_conductor(float cos_theta_i, const float eta[3], const float k[3], float out[3]) { // fresnel.h:65-88
    Color3 r = fresnel_conductor(cos_theta_i, Color3(eta[0], eta[1], eta[2]), Color3(k[0], k[1], k[2]));
    for (int i = 0; i < 3; ++i) out[i] = r.coeff(i);
}
void ref_reflect_refract(const float wi[3], const float m[3], float cos_theta_t, float eta_ti, float refl[3], float refr[3]) { // fresnel.h:16-34
    Vector3f a(wi[0], wi[1], wi[2]), b(m[0], m[1], m[2]);
    Vector3f r = reflect<float>(a, b), t = refract<float>(a, b, cos_theta_t, eta_ti);
    for (int i = 0; i < 3; ++i) { refl[i] = r[i]; refr[i] = t[i]; }
}
// which: 0 eval(a) 1 pdf(wi = a, m = b) 2 sample(wi = a, sample = b.xy) -> m, pdf 3 G(wi = a, wo = b, m = c) 4 smith_g1(v = a, m = b)
void ref_ggx(int which, float au, float av, const float a[3], const float b[3], const float c[3], float out[4]) { // microfacet.h:11-43,108-175
    MicrofacetDistribution d(MicrofacetDistribution::Type::GGX, au, av, false);
    Vector3f A(a[0], a[1], a[2]), B(b[0], b[1], b[2]), C(c[0], c[1], c[2]);
    out[0] = out[1] = out[2] = out[3] = 0.f;
    if (which == 0) out[0] = d.eval(A);
    else if (which == 1) out[0] = d.pdf(A, B);
    else if (which == 2) { auto [m, pdf] = d.sample(A, Vector2f(b[0], b[1])); out[0] = m.x(); out[1] = m.y(); out[2] = m.z(); out[3] = pdf; }
    else if (which == 3) out[0] = d.G(A, B, C);
    else out[0] = d.smith_g1(A, B);
}
void ref_sample_wavelength(float u, float wl[4], float weight[4]) { // spectrum.h:152-181, mathutils.h:167-182
    auto [w, p] = sample_wavelength<float, 4>(u);
    for (int i = 0; i < 4; ++i) { wl[i] = w.coeff(i); weight[i] = p.coeff(i); }
}
void ref_spectrum_to_xyz(const float value[4], const float wl[4], float xyz[3]) { // spectrum.h:83-115 (table: src/librender/spectrum.cpp)
    Spectrum v(value[0], value[1], value[2], value[3]);
    Wavelength w(wl[0], wl[1], wl[2], wl[3]);
    auto r = spectrum_to_xyz(v, w);
    for (int i = 0; i < 3; ++i) xyz[i] = r[i];
}
void ref_xyz_to_srgb(const float xyz[3], float rgb[3]) { // spectrum.h:138-143
    Vector3f r = xyz_to_srgb(Vector3f(xyz[0], xyz[1], xyz[2]));
    for (int i = 0; i < 3; ++i) rgb[i] = r[i];
}
void ref_srgb_model_eval(const float c[3], const float wl[4], float out[4]) { // srgb.h:8-19
    Spectrum r = srgb_model_eval(Color3(c[0], c[1], c[2]), Wavelength(wl[0], wl[1], wl[2], wl[3]));
    for (int i = 0; i < 4; ++i) out[i] = r.coeff(i);
}
// Distribution1D built from `pdf` (distribution.h:14-117): out = index, reused sample, for each u
void ref_distribution_sample_reuse(const float *pdf, size_t n, const float *u, size_t nu, uint32_t *index, float *reused, float *cdf_out) {
    Distribution1D d;
    d.init(pdf, (int) n); // as Mesh::area_distr_build does, mesh.cpp:39-48
    for (size_t i = 0; i < nu; ++i) { auto [idx, s] = d.sample_reuse(u[i]); index[i] = idx; reused[i] = s; }
    if (cdf_out) for (size_t i = 0; i < d.cdf().size(); ++i) cdf_out[i] = d.cdf()[i];
}

} // extern "C"
