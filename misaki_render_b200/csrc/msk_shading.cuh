// Device-side restatement of the per-vertex work of PathTracer::sample
// (reference src/librender/integrators/path.cpp:23-131): sampler, camera, spectra,
// hit reconstruction, emitters and the BSDFs.  Each block cites the reference
// file:line whose result it must reproduce (paths relative to the reference root).
// Arithmetic is float32 with FMA contraction; transcendental functions are CUDA's
// (<= 2 ulp), so parity with the CPU oracle is to tolerance, not bit-exact.
#pragma once
#include "msk_device.cuh"

namespace msk {

// ------------------------------------------------------------------ small vector helpers
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 normalize(V3 a) { // Eigen normalized(): v / sqrt(squaredNorm), true division
    float z = dot(a, a);
    if (z > 0.f) { float n = sqrtf(z); return v3(a.x / n, a.y / n, a.z / n); }
    return a;
}
__device__ __forceinline__ float max_abs(V3 a) { return fmaxf(fabsf(a.x), fmaxf(fabsf(a.y), fabsf(a.z))); }

// 4 hero wavelengths in one float4
__device__ __forceinline__ float4 f4(float s) { return make_float4(s, s, s, s); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator/(float4 a, float4 b) { return make_float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 operator*(float s, float4 a) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 operator/(float4 a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }
__device__ __forceinline__ float4 operator+(float4 a, float s) { return make_float4(a.x + s, a.y + s, a.z + s, a.w + s); }
__device__ __forceinline__ float4 operator-(float4 a, float s) { return make_float4(a.x - s, a.y - s, a.z - s, a.w - s); }
__device__ __forceinline__ float4 sqrt4(float4 a) { return make_float4(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z), sqrtf(a.w)); }
__device__ __forceinline__ float hmax(float4 a) { return fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)); }
__device__ __forceinline__ bool is_zero(float4 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f && a.w == 0.f; }

constexpr float kPi = 3.14159265358979323846f, kInvPi = 0.31830988618379067154f, kInvFourPi = 0.07957747154594766788f;
constexpr float kEpsilon = 5.9604644775390625e-08f;          // mathutils.h:16-17 (epsilon/2)
constexpr float kRayEpsilon = kEpsilon * 1500, kShadowEpsilon = kRayEpsilon * 10; // mathutils.h:19-20
// tmin of a visibility ray spawned at o (interaction.h:40-44 scaled as scene.cpp:91-93).  A function of the origin alone: the
// shadow queue stores the path index in the ray record's tmin field and k_shadow recomputes this value (ShadowIO::tag).
__device__ __forceinline__ float shadow_tmin(float ox, float oy, float oz) { return kRayEpsilon * (1.f + max_abs(v3(ox, oy, oz))); }
#define MSK_INF __int_as_float(0x7f800000)

__device__ __forceinline__ float sqr(float a) { return a * a; }
__device__ __forceinline__ float safe_sqrt(float a) { return sqrtf(fmaxf(a, 0.f)); }

// ------------------------------------------------------------------ PCG32 (mathutils.h:85-143) + IndependentSampler
constexpr uint64_t kPcgMult = 0x5851f42d4c957f2dULL;
constexpr uint64_t kPcgInc  = (0xda3e39cb94b95bdbULL << 1) | 1ull; // seed(.., PCG32_DEFAULT_STREAM), independent.cpp:25

__device__ __forceinline__ uint32_t pcg_next(uint64_t &state) {
    uint64_t old = state;
    state = old * kPcgMult + kPcgInc;
    uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
__device__ __forceinline__ uint64_t pcg_seed(uint64_t initstate) { // mathutils.h:95-101
    uint64_t s = 0;
    pcg_next(s);
    s += initstate;
    pcg_next(s);
    return s;
}
__device__ __forceinline__ float next1d(uint64_t &state) { // mathutils.h:111-120
    return __uint_as_float((pcg_next(state) >> 9) | 0x3f800000u) - 1.0f;
}

// ------------------------------------------------------------------ frames / warps
__device__ __forceinline__ void coordinate_system(V3 n, V3 &s, V3 &t) { // mathutils.h:196-203
    float sign = copysignf(1.f, n.z);
    float a = -1.f / (sign + n.z);
    float b = n.x * n.y * a;
    s = v3(1.f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    t = v3(b, sign + n.y * n.y * a, -n.y);
}
struct Frame { V3 s, t, n; };
__device__ __forceinline__ V3 to_local(const Frame &f, V3 v) { return v3(dot(v, f.s), dot(v, f.t), dot(v, f.n)); }
__device__ __forceinline__ V3 to_world(const Frame &f, V3 v) { return f.s * v.x + f.t * v.y + f.n * v.z; }

// sin / cos of an angle that is known to lie within a few multiples of pi (every call site below): CUDA's sincosf carries a
// Payne-Hanek slow path for huge arguments -- never taken here, but ~150 instructions of every shade kernel's code
// (profiles/r01g_ncu_k_shade.txt: stall_no_instruction).  sincospif reduces its argument exactly and has no slow path; the
// division by pi costs one rounding of the angle (<= 1 ulp), the same size as the rounding of the angle itself.
__device__ __forceinline__ void sincos_of(float phi, float *sn, float *cs) { sincospif(phi * kInvPi, sn, cs); }

__device__ __forceinline__ V3 square_to_cosine_hemisphere(float sx, float sy) { // warp.h:17-43
    float x = 2.f * sx - 1.f, y = 2.f * sy - 1.f;
    float phi, r;
    if (x == 0 && y == 0) { r = phi = 0; }
    else if (x * x > y * y) { r = x; phi = (kPi / 4.f) * (y / x); }
    else { r = y; phi = (kPi / 2.f) - (x / y) * (kPi / 4.f); }
    float sn, cs;
    sincos_of(phi, &sn, &cs);
    float px = r * cs, py = r * sn;
    return v3(px, py, safe_sqrt(1.f - (px * px + py * py)));
}
__device__ __forceinline__ V3 square_to_uniform_sphere(float sx, float sy) { // warp.h:46-53
    float z = -2.f * sy + 1.f, r = safe_sqrt(-z * z + 1.f);
    float sn, cs;
    sincospif(2.f * sx, &sn, &cs); // sin / cos (2 pi sx)
    return v3(r * cs, r * sn, z);
}

// ------------------------------------------------------------------ wavelengths (spectrum.h:152-181)
__device__ __forceinline__ void sample_wavelength(float sample, float4 &wl, float4 &weight) {
    float l[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float value = sample + (float) i / 4.f;
        float u = (value <= 1.f) ? value : value - 1.f;
        l[i] = 538.f - atanhf(0.8569106254698279f - 1.8275019724092267f * u) * 138.88888888888889f;
        float tmp = coshf(0.0072f * (l[i] - 538.f));
        w[i] = 253.82f * tmp * tmp;
    }
    wl = make_float4(l[0], l[1], l[2], l[3]);
    weight = make_float4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------ spectra (src/librender/spectra/*.cpp, srgb.h:8-19)
__device__ __forceinline__ float srgb1(float c0, float c1, float c2, float l) {
    float v = (c0 * l + c1) * l + c2;
    return fmaxf(.5f * v * (1.f / sqrtf(v * v + 1.f)) + .5f, 0.f);
}
__device__ __forceinline__ float4 srgb_model_eval(const DSpectrum &s, float4 wl) {
    if (isinf(s.c2)) return f4(copysignf(1.f, s.c2) * .5f + .5f);
    return make_float4(srgb1(s.c0, s.c1, s.c2, wl.x), srgb1(s.c0, s.c1, s.c2, wl.y), srgb1(s.c0, s.c1, s.c2, wl.z),
                       srgb1(s.c0, s.c1, s.c2, wl.w));
}
__device__ __forceinline__ float regular1(const DSpectrum &s, const float *__restrict__ tables, float l) { // regular.cpp:73-91
    float x = (l - s.lambda_min) * s.inv_interval;
    uint32_t idx = min((uint32_t) x, s.table_size - 2u);
    const float *t = tables + s.table_offset + idx;
    float y0 = __ldg(t), y1 = __ldg(t + 1);
    float w1 = x - (float) idx, w0 = 1.f - w1;
    return w0 * y0 + w1 * y1;
}
__device__ __forceinline__ float4 regular_eval(const DSpectrum &s, const float *__restrict__ tables, float4 wl) {
    return make_float4(regular1(s, tables, wl.x), regular1(s, tables, wl.y), regular1(s, tables, wl.z), regular1(s, tables, wl.w));
}
#ifndef MSK_SPEC_NOINLINE
#define MSK_SPEC_NOINLINE 0
#endif
#if MSK_SPEC_NOINLINE
__device__ __noinline__ float4 spectrum_eval(const DScene &sc, int id, float4 wl) {
#else
__device__ __forceinline__ float4 spectrum_eval(const DScene &sc, int id, float4 wl) {
#endif
    const DSpectrum s = sc.spectra[id];
    switch (s.kind) {
        case MSK_SPEC_UNIFORM: { // uniform.cpp:19-26
            bool in = wl.x >= 360.f && wl.y >= 360.f && wl.z >= 360.f && wl.w >= 360.f && wl.x <= 830.f && wl.y <= 830.f &&
                      wl.z <= 830.f && wl.w <= 830.f;
            return in ? f4(s.value) : f4(0.f);
        }
        case MSK_SPEC_SRGB: return srgb_model_eval(s, wl);
        case MSK_SPEC_SRGB_D65: return regular_eval(s, sc.tables, wl) * srgb_model_eval(s, wl);
        case MSK_SPEC_REGULAR: return regular_eval(s, sc.tables, wl);
        case MSK_SPEC_SRGB_UNBOUNDED: return srgb_model_eval(s, wl) * s.value;
    }
    return f4(0.f);
}

// textures/checkerboard.cpp:16-31: a uv-dependent texture only SELECTS one of its children, so it is resolved to
// the id of a plain spectrum once per surface point (children have smaller ids: the loop terminates).
// DSpectrum of a checkerboard: table_offset/table_size = child ids, (c0 c1 c2 | value lambda_min inv_interval) = the
// two rows of Transform3f "to_uv"; the products are kept un-contracted like the oracle's.
__device__ __forceinline__ int texture_resolve(const DScene &sc, int id, float u, float v) {
    while (id >= 0) {
        const DSpectrum &s = sc.spectra[id];
        if (s.kind != MSK_SPEC_CHECKERBOARD) break;
        float tu = __fadd_rn(__fadd_rn(__fmul_rn(s.c0, u), __fmul_rn(s.c1, v)), s.c2);
        float tv = __fadd_rn(__fadd_rn(__fmul_rn(s.value, u), __fmul_rn(s.lambda_min, v)), s.inv_interval);
        tu -= floorf(tu); tv -= floorf(tv);
        id = ((tu > .5f) == (tv > .5f)) ? (int) s.table_offset : (int) s.table_size;
    }
    return id;
}
__device__ __forceinline__ void bsdf_resolve_textures(const DScene &sc, MskBsdf &b, float u, float v) {
    b.reflectance = texture_resolve(sc, b.reflectance, u, v);
    b.transmittance = texture_resolve(sc, b.transmittance, u, v);
    b.eta = texture_resolve(sc, b.eta, u, v);
    b.k = texture_resolve(sc, b.k, u, v);
}

// spectrum.h:83-115; the 4-wide mean adds (v0+v2)+(v1+v3) like Eigen's packet reduction
__device__ __forceinline__ void spectrum_to_xyz(const DScene &sc, float4 value, float4 wl, float &X, float &Y, float &Z) {
    float l[4] = { wl.x, wl.y, wl.z, wl.w }, v[4] = { value.x, value.y, value.z, value.w };
    float x[4], y[4], z[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float t = (l[s] - 360.f) * (94.f / 470.f);
        uint32_t i0 = min((uint32_t) t, 93u);
        float4 r0 = __ldg(sc.cie + i0), r1 = __ldg(sc.cie + i0 + 1);
        float w1 = t - (float) i0, w0 = 1.f - w1;
        x[s] = (w0 * r0.x + w1 * r1.x) * v[s];
        y[s] = (w0 * r0.y + w1 * r1.y) * v[s];
        z[s] = (w0 * r0.z + w1 * r1.z) * v[s];
    }
    X = ((x[0] + x[2]) + (x[1] + x[3])) / 4.f;
    Y = ((y[0] + y[2]) + (y[1] + y[3])) / 4.f;
    Z = ((z[0] + z[2]) + (z[1] + z[3])) / 4.f;
}

// ------------------------------------------------------------------ camera (perspective.cpp:22-41)
__device__ __forceinline__ void camera_ray(const DCamera &cam, float px, float py, V3 &o, V3 &d, float &mint, float &maxt) {
    const float *M = cam.s2c;
    float r0 = M[0] * px + M[1] * py + M[2] * 0.f + M[3], r1 = M[4] * px + M[5] * py + M[6] * 0.f + M[7],
          r2 = M[8] * px + M[9] * py + M[10] * 0.f + M[11], r3 = M[12] * px + M[13] * py + M[14] * 0.f + M[15];
    V3 dl = normalize(v3(r0 / r3, r1 / r3, r2 / r3));
    float inv_z = 1.f / dl.z;
    mint = cam.near_clip * inv_z;
    maxt = cam.far_clip * inv_z;
    const float *T = cam.c2w;
    float w = T[15];
    o = v3(T[3] / w, T[7] / w, T[11] / w);
    d = v3(T[0] * dl.x + T[1] * dl.y + T[2] * dl.z, T[4] * dl.x + T[5] * dl.y + T[6] * dl.z, T[8] * dl.x + T[9] * dl.y + T[10] * dl.z);
}

// ------------------------------------------------------------------ Fresnel (fresnel.h) / GGX (microfacet.h)
__device__ __forceinline__ V3 reflect_m(V3 wi, V3 m) { return m * (2.f * dot(wi, m)) - wi; } // fresnel.h:16-20
__device__ __forceinline__ V3 refract_m(V3 wi, V3 m, float cos_theta_t, float eta_ti) {       // fresnel.h:29-34
    return m * (dot(wi, m) * eta_ti + cos_theta_t) - wi * eta_ti;
}
struct FresnelR { float F, cos_theta_t, eta_it, eta_ti; };
__device__ __forceinline__ FresnelR fresnel_dielectric(float cos_theta_i, float eta) { // fresnel.h:37-63
    float eta_it = cos_theta_i >= 0.f ? eta : 1.f / eta, eta_ti = cos_theta_i >= 0.f ? 1.f / eta : eta;
    float ctt2 = 1.f - eta_ti * eta_ti * (1.f - cos_theta_i * cos_theta_i);
    float ci = fabsf(cos_theta_i), ct = safe_sqrt(ctt2);
    float a_s = (ci - eta_it * ct) / (ci + eta_it * ct), a_p = (ct - eta_it * ci) / (ct + eta_it * ci);
    float r = (eta == 1.f || ci == 0.f) ? (eta == 1.f ? 0.f : 1.f) : 0.5f * (a_s * a_s + a_p * a_p);
    return FresnelR{ r, ct * copysignf(1.f, -cos_theta_i), eta_it, eta_ti };
}
__device__ __forceinline__ float4 fresnel_conductor(float cos_theta_i, float4 eta, float4 k) { // fresnel.h:65-88
    float c2 = cos_theta_i * cos_theta_i, s2 = 1.f - c2, s4 = s2 * s2;
    float4 temp_1 = eta * eta - k * k - s2;
    float4 a2pb2 = sqrt4(temp_1 * temp_1 + 4.f * k * k * eta * eta);
    float4 a = sqrt4(.5f * (a2pb2 + temp_1));
    float4 term_1 = a2pb2 + c2, term_2 = 2.f * cos_theta_i * a;
    float4 r_s = (term_1 - term_2) / (term_1 + term_2);
    float4 term_3 = a2pb2 * c2 + s4, term_4 = term_2 * s2;
    float4 r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return .5f * (r_s + r_p);
}

struct Ggx { float au, av; };
__device__ __forceinline__ Ggx ggx_make(float au, float av) { return Ggx{ fmaxf(au, 1e-4f), fmaxf(av, 1e-4f) }; } // microfacet.h:190-193
__device__ __forceinline__ float ggx_eval(const Ggx &g, V3 m) { // microfacet.h:11-18,108-125
    if (m.z <= 0.f) return 0.f;
    float c2 = m.z * m.z;
    float e = ((m.x * m.x / (g.au * g.au)) + (m.y * m.y) / (g.av * g.av)) / c2;
    float root = (1.f + e) * c2;
    float result = 1.f / (kPi * g.au * g.av * root * root);
    return result * m.z > 1e-20f ? result : 0.f;
}
__device__ __forceinline__ float ggx_pdf(const Ggx &g, V3 m) { return ggx_eval(g, m) * m.z; } // microfacet.h:127-129
__device__ __forceinline__ V3 ggx_sample(const Ggx &g, float sx, float sy, float &pdf) {   // microfacet.h:20-40
    float sin_phi, cos_phi;
    if (g.au == g.av) {
        // isotropic: phi_m = atan(tan(pi + 2 pi sy)) + pi floor(2 sy + 1/2) is 2 pi sy up to a multiple of 2 pi, so its sine and
        // cosine are those of 2 pi sy -- no tan / atan round trip (~150 instructions per vertex of the C2 / C3 materials)
        sincospif(2.f * sy, &sin_phi, &cos_phi);
    } else {
        // tan(pi + 2 pi sy) = tan(2 pi sy) = sinpi(2 sy) / cospi(2 sy): exact argument reduction, no Payne-Hanek path
        float st, ct;
        sincospif(2.f * sy, &st, &ct);
        float phi_m = atanf(g.au / g.av * (st / ct)) + kPi * floorf(2 * sy + 0.5f);
        sincos_of(phi_m, &sin_phi, &cos_phi);
    }
    float c = cos_phi / g.au, s = sin_phi / g.av;
    float alpha_sqr = 1.f / (c * c + s * s);
    float tan2 = alpha_sqr * sx / (1.f - sx);
    float cos_m = 1.f / sqrtf(1.f + tan2);
    float tmp = 1 + tan2 / alpha_sqr;
    pdf = kInvPi / (g.au * g.av * cos_m * cos_m * cos_m * tmp * tmp);
    if (pdf < 1e-20f) pdf = 0;
    float sin_m = safe_sqrt(1 - cos_m * cos_m);
    return v3(sin_m * cos_phi, sin_m * sin_phi, cos_m);
}
__device__ __forceinline__ float ggx_g1(const Ggx &g, V3 v, V3 m) { // microfacet.h:150-175
    float xy = sqr(g.au * v.x) + sqr(g.av * v.y), t2 = xy / sqr(v.z);
    if (xy == 0.f) return 1.f;
    if (dot(v, m) * v.z <= 0.f) return 0.f;
    return 2.f / (1.f + sqrtf(1.f + t2));
}
__device__ __forceinline__ float ggx_G(const Ggx &g, V3 wi, V3 wo, V3 m) { return ggx_g1(g, wi, m) * ggx_g1(g, wo, m); }

// ------------------------------------------------------------------ BSDFs (src/librender/bsdfs/*.cpp)
enum : uint32_t {
    BF_Null = 0x1, BF_DiffuseReflection = 0x2, BF_GlossyReflection = 0x8, BF_GlossyTransmission = 0x10,
    BF_DeltaReflection = 0x20, BF_DeltaTransmission = 0x40, BF_Delta = 0x61
};
__device__ __forceinline__ bool bsdf_is_smooth(int type) { // has_flag(flags, Smooth), path.cpp:56
    return type == MSK_BSDF_DIFFUSE || type == MSK_BSDF_ROUGHCONDUCTOR || type == MSK_BSDF_ROUGHDIELECTRIC;
}

struct BsdfSample { V3 wo; float pdf, eta; uint32_t type; float4 weight; };

// The spectra a BSDF reads (Texture::eval(si) of "reflectance"/"specular_reflectance", "specular_transmittance",
// "eta", "k"), evaluated ONCE per path vertex by eval_vertex_spectra() below: sample() and eval() of the same vertex
// used to re-evaluate them at ten inlined call sites, which made the rough-conductor shade kernel 8840 SASS
// instructions (141 KB) and instruction-fetch bound (stall no_instruction = 4.5 per issue, profiles/r01f_ncu_k_shade.txt).
struct BsdfSpectra { float4 R, Tr, eta, k; };

// sample(): wi is the local incident direction; returns weight = f*cos/pdf
// TYPE >= 0: the BSDF type is known at compile time (k_shade specialised per material key) and the switch folds
// to one case; TYPE < 0: dispatch on b.type.
template <int TYPE>
__device__ __forceinline__ BsdfSample bsdf_sample_1(const BsdfSpectra &sp, const MskBsdf &b, V3 wi, float s1, float s2x, float s2y) {
    BsdfSample bs;
    bs.wo = v3(0, 0, 0); bs.pdf = 0.f; bs.eta = 1.f; bs.type = 0; bs.weight = f4(0.f);
    const float ci = wi.z;
    switch (TYPE >= 0 ? TYPE : b.type) {
        case MSK_BSDF_DIFFUSE: { // diffuse.cpp:19-32
            if (ci <= 0.f) return bs;
            bs.wo = square_to_cosine_hemisphere(s2x, s2y);
            bs.pdf = kInvPi * bs.wo.z;
            bs.type = BF_DiffuseReflection;
            if (bs.pdf > 0.f) bs.weight = sp.R;
            return bs;
        }
        case MSK_BSDF_CONDUCTOR: { // conductor.cpp:22-39
            if (ci <= 0.f) return bs;
            bs.wo = v3(-wi.x, -wi.y, wi.z); bs.pdf = 1.f; bs.type = BF_DeltaReflection;
            bs.weight = sp.R * fresnel_conductor(ci, sp.eta, sp.k);
            return bs;
        }
        case MSK_BSDF_ROUGHCONDUCTOR: { // roughconductor.cpp:53-80
            if (ci <= 0.f) return bs;
            Ggx g = ggx_make(b.alpha_u, b.alpha_v);
            V3 m = ggx_sample(g, s2x, s2y, bs.pdf);
            bs.wo = reflect_m(wi, m);
            bs.type = BF_GlossyReflection;
            if (!(bs.pdf != 0.f && bs.wo.z > 0.f)) return bs;
            float weight = b.sample_visible ? ggx_g1(g, bs.wo, m) : ggx_G(g, wi, bs.wo, m) * dot(wi, m) / (ci * m.z);
            bs.pdf /= 4.f * dot(bs.wo, m);
            bs.weight = fresnel_conductor(dot(wi, m), sp.eta, sp.k) * weight;
            return bs;
        }
        case MSK_BSDF_ROUGHDIELECTRIC: { // roughdielectric.cpp:58-114
            float eta = b.int_ior / b.ext_ior;
            Ggx g = ggx_make(b.alpha_u, b.alpha_v), gs = g;
            if (!b.sample_visible) { float sc_ = 1.2f - .2f * sqrtf(fabsf(ci)); gs.au *= sc_; gs.av *= sc_; }
            V3 m = ggx_sample(gs, s2x, s2y, bs.pdf);
            if (bs.pdf == 0) return bs;
            FresnelR fr = fresnel_dielectric(dot(wi, m), eta);
            bool sel_r = s1 <= fr.F;
            float4 weight = f4(1.f);
            bs.pdf *= sel_r ? fr.F : (1.f - fr.F);
            bs.eta = sel_r ? 1.f : fr.eta_it;
            bs.type = sel_r ? BF_GlossyReflection : BF_GlossyTransmission;
            float dwh_dwo;
            if (sel_r) {
                bs.wo = reflect_m(wi, m);
                weight = weight * sp.R;
                dwh_dwo = 1.f / (4.f * dot(bs.wo, m));
            } else {
                bs.wo = refract_m(wi, m, fr.cos_theta_t, fr.eta_ti);
                weight = weight * sqr(fr.eta_ti);
                dwh_dwo = sqr(bs.eta) * dot(bs.wo, m) / sqr(dot(wi, m) + bs.eta * dot(bs.wo, m));
            }
            weight = weight * (b.sample_visible ? ggx_g1(g, bs.wo, m) : ggx_G(g, wi, bs.wo, m) * dot(wi, m) / (ci * m.z));
            bs.pdf *= fabsf(dwh_dwo);
            bs.weight = weight;
            return bs;
        }
        case MSK_BSDF_DIELECTRIC: { // dielectric.cpp:26-72
            FresnelR fr = fresnel_dielectric(ci, b.int_ior / b.ext_ior);
            bool sel_r = s2x <= fr.F;
            bs.pdf = sel_r ? fr.F : 1.f - fr.F;
            bs.type = sel_r ? BF_DeltaReflection : BF_DeltaTransmission;
            bs.wo = sel_r ? v3(-wi.x, -wi.y, wi.z) : v3(-fr.eta_ti * wi.x, -fr.eta_ti * wi.y, fr.cos_theta_t);
            bs.eta = sel_r ? 1.f : fr.eta_it;
            bs.weight = sel_r ? sp.R : sp.Tr * fr.eta_ti * fr.eta_ti;
            return bs;
        }
    }
    return bs;
}

// eval() and pdf() of the NEE direction in one pass (path.cpp:61-62)
template <int TYPE>
__device__ __forceinline__ void bsdf_eval_pdf_1(const BsdfSpectra &sp, const MskBsdf &b, V3 wi, V3 wo, float4 &val, float &pdf) {
    val = f4(0.f); pdf = 0.f;
    const float ci = wi.z, co = wo.z;
    switch (TYPE >= 0 ? TYPE : b.type) {
        case MSK_BSDF_DIFFUSE: // diffuse.cpp:34-57
            if (ci > 0.f && co > 0.f) { val = sp.R * kInvPi * co; pdf = kInvPi * co; }
            return;
        case MSK_BSDF_ROUGHCONDUCTOR: { // roughconductor.cpp:82-120
            if (!(ci > 0.f && co > 0.f)) return;
            V3 H = normalize(wo + wi);
            Ggx g = ggx_make(b.alpha_u, b.alpha_v);
            float D = ggx_eval(g, H);
            if (D != 0.f) {
                float G = ggx_G(g, wi, wo, H);
                float result = D * G / (4.f * ci);
                float4 F = fresnel_conductor(dot(wi, H), sp.eta, sp.k);
                val = F * sp.R * result;
            }
            if (dot(wi, H) > 0.f && dot(wo, H) > 0.f)
                pdf = b.sample_visible ? ggx_eval(g, H) * ggx_g1(g, wi, H) / (4.f * ci) : ggx_pdf(g, H) / (4.f * dot(wo, H));
            return;
        }
        case MSK_BSDF_ROUGHDIELECTRIC: { // roughdielectric.cpp:116-190
            if (ci == 0.f) return;
            float m_eta = b.int_ior / b.ext_ior, m_inv_eta = b.ext_ior / b.int_ior;
            bool refl = ci * co > 0.f;
            float eta = ci > 0.f ? m_eta : m_inv_eta, inv_eta = ci > 0.f ? m_inv_eta : m_eta;
            V3 m = normalize(wi + wo * (refl ? 1.f : eta));
            m = m * copysignf(1.f, m.z);
            Ggx g = ggx_make(b.alpha_u, b.alpha_v);
            float D = ggx_eval(g, m);
            float F = fresnel_dielectric(dot(wi, m), m_eta).F;
            float G = ggx_G(g, wi, wo, m);
            if (refl) val = F * D * G * sp.R / (4.f * fabsf(ci));
            else {
                float scale = sqr(inv_eta);
                val = sp.Tr *
                      fabsf((scale * (1.f - F) * D * G * eta * eta * dot(wi, m) * dot(wo, m)) / (ci * sqr(dot(wi, m) + eta * dot(wo, m))));
            }
            if (dot(wi, m) * ci <= 0.f || dot(wo, m) * co <= 0.f) return;
            float dwh_dwo = refl ? 1.f / (4.f * dot(wo, m)) : (eta * eta * dot(wo, m)) / sqr(dot(wi, m) + eta * dot(wo, m));
            Ggx gs = g;
            if (!b.sample_visible) { float sc_ = 1.2f - .2f * sqrtf(fabsf(ci)); gs.au *= sc_; gs.av *= sc_; }
            float prob = ggx_pdf(gs, m);
            prob *= refl ? F : 1.f - F;
            pdf = prob * fabsf(dwh_dwo);
            return;
        }
        default: return;
    }
}

// twosided.cpp:38-101 (same BRDF on both sides)
// (one call site of the inner function: flip wi, sample, flip wo back)
template <int TYPE>
__device__ __forceinline__ BsdfSample bsdf_sample(const BsdfSpectra &sp, const MskBsdf &b, V3 wi, float s1, float s2x, float s2y) {
    const bool two = b.twosided != 0, flip = two && wi.z < 0.f;
    if (two && wi.z == 0.f) {
        BsdfSample bs;
        bs.wo = v3(0, 0, 0); bs.pdf = 0.f; bs.eta = 1.f; bs.type = 0; bs.weight = f4(0.f);
        return bs;
    }
    if (flip) wi.z = -wi.z;
    BsdfSample bs = bsdf_sample_1<TYPE>(sp, b, wi, s1, s2x, s2y);
    if (flip) bs.wo.z = -bs.wo.z;
    return bs;
}
template <int TYPE>
__device__ __forceinline__ void bsdf_eval_pdf(const BsdfSpectra &sp, const MskBsdf &b, V3 wi, V3 wo, float4 &val, float &pdf) {
    if (b.twosided) {
        if (wi.z == 0.f) { val = f4(0.f); pdf = 0.f; return; }
        if (wi.z < 0.f) { wi.z = -wi.z; wo.z = -wo.z; }
    }
    bsdf_eval_pdf_1<TYPE>(sp, b, wi, wo, val, pdf);
}

__device__ __forceinline__ float mis_weight(float pdf_a, float pdf_b) { // path.cpp:127-131
    pdf_a *= pdf_a; pdf_b *= pdf_b;
    return pdf_a > 0.f ? pdf_a / (pdf_a + pdf_b) : 0.f;
}

// ------------------------------------------------------------------ hit reconstruction (mesh.cpp:51-101, interaction.h:55-60)
struct Surface {
    V3 p, n;      // position (barycentric), geometric normal
    Frame sh;     // shading frame
    float uvx, uvy; // si.uv: interpolated texcoords, or the barycentrics when the mesh has none (mesh.cpp:65-72)
};

__device__ __forceinline__ Surface make_surface(const DScene &sc, const DMeshInfo &mi, uint32_t prim, float bu, float bv) {
    const uint32_t *ip = sc.indices + 3 * (size_t) (mi.tri_offset + prim);
    uint32_t i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
    const float4 *vp = sc.verts + 2 * (size_t) mi.vert_offset;
    float4 a0 = __ldg(vp + 2 * (size_t) i0), b0 = __ldg(vp + 2 * (size_t) i1), c0 = __ldg(vp + 2 * (size_t) i2);
    V3 p0 = v3(a0.x, a0.y, a0.z), p1 = v3(b0.x, b0.y, b0.z), p2 = v3(c0.x, c0.y, c0.z);
    float b1 = bu, b2 = bv, bb0 = 1.f - b1 - b2;
    V3 dp0 = p1 - p0, dp1 = p2 - p0;
    Surface s;
    s.p = p0 * bb0 + p1 * b1 + p2 * b2;
    s.n = normalize(cross(dp0, dp1));
    V3 dp_du, dp_dv;
    coordinate_system(s.n, dp_du, dp_dv);
    float4 a1, b1v, c1;
    if (mi.flags) { a1 = __ldg(vp + 2 * (size_t) i0 + 1); b1v = __ldg(vp + 2 * (size_t) i1 + 1); c1 = __ldg(vp + 2 * (size_t) i2 + 1); }
    s.uvx = bu; s.uvy = bv;
    if (mi.flags & 2u) {
        s.uvx = a1.z * bb0 + b1v.z * b1 + c1.z * b2; s.uvy = a1.w * bb0 + b1v.w * b1 + c1.w * b2;
        float du0 = b1v.z - a1.z, dv0 = b1v.w - a1.w, du1 = c1.z - a1.z, dv1 = c1.w - a1.w;
        float det = du0 * dv1 - dv0 * du1, inv_det = 1.f / det;
        if (det != 0.f) dp_du = (dv1 * dp0 - dv0 * dp1) * inv_det;
    }
    if (mi.flags & 1u) {
        V3 n0 = v3(a0.w, a1.x, a1.y), n1 = v3(b0.w, b1v.x, b1v.y), n2 = v3(c0.w, c1.x, c1.y);
        s.sh.n = normalize(n0 * bb0 + n1 * b1 + n2 * b2);
    } else {
        s.sh.n = s.n;
    }
    V3 ff = -s.sh.n * dot(s.sh.n, dp_du) + dp_du;
    s.sh.s = normalize(ff);
    s.sh.t = cross(s.sh.n, s.sh.s);
    return s;
}

// ------------------------------------------------------------------ emitters (emitters/area.cpp, emitters/constant.cpp, shape.cpp:64-86)
struct NeeSample {
    V3 d;            // direction towards the light (world)
    float dist, pdf; // pdf == 0: no contribution
    int radiance;    // spectrum id of the sampled emitter's radiance at the sampled point (textures resolved), -1: none
    float pdf0;      // value = spectrum_eval(radiance) / pdf0 [* scale]: emitter_sample_direct's division, then scene.cpp:86-87
    float scale;     //   (number of emitters when a light was selected at random, else 1)
    float stale_pdf; // pdf_emitter_direct() of this record, consumed only by the env-miss quirk (q8)
};

__device__ __forceinline__ float4 nee_value(const NeeSample &ns, float4 radiance) { // radiance / pdf (before the visibility test)
    float4 v = radiance / ns.pdf0;
    return ns.scale != 1.f ? v * ns.scale : v;
}
__device__ __forceinline__ NeeSample sample_emitter_direct(const DScene &sc, V3 ref_p, float sx, float sy) { // scene.cpp:69-89
    NeeSample r;
    r.pdf = 0.f; r.radiance = -1; r.pdf0 = 1.f; r.scale = 1.f; r.stale_pdf = 0.f; r.dist = 0.f; r.d = v3(0, 0, 0);
    uint32_t ne = sc.nemitters;
    if (!ne) return r;
    uint32_t index = 0;
    float sel = 1.f;
    if (ne > 1) {
        sel = 1.f / (float) ne;
        index = min((uint32_t) (sx * (float) ne), ne - 1u);
        sx = (sx - (float) index * sel) * (float) ne;
    }
    const MskEmitter em = sc.emitters[index];
    if (em.type == MSK_EMITTER_AREA) { // area.cpp:33-45 + shape.cpp:64-78 + mesh.cpp:103-133
        const DMeshInfo mi = sc.meshes[em.shape];
        const float *cdf = sc.cdfs + mi.cdf_offset;
        // Distribution1D::sample_reuse(sample.y): upper_bound over ntris+1 entries, distribution.h:114-123
        uint32_t lo = 0, hi = mi.ntris + 1;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (__ldg(cdf + mid) <= sy) lo = mid + 1; else hi = mid; }
        int face = min(max((int) lo - 1, 0), (int) mi.ntris - 1);
        float c0 = __ldg(cdf + face), c1 = __ldg(cdf + face + 1);
        sy = (sy - c0) / (c1 - c0);
        const uint32_t *ip = sc.indices + 3 * (size_t) (mi.tri_offset + face);
        uint32_t i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
        const float4 *vp = sc.verts + 2 * (size_t) mi.vert_offset;
        float4 a0 = __ldg(vp + 2 * (size_t) i0), b0 = __ldg(vp + 2 * (size_t) i1), c0v = __ldg(vp + 2 * (size_t) i2);
        V3 p0 = v3(a0.x, a0.y, a0.z), p1 = v3(b0.x, b0.y, b0.z), p2 = v3(c0v.x, c0v.y, c0v.z);
        V3 e0 = p1 - p0, e1 = p2 - p0;
        float t = safe_sqrt(1.f - sx);
        float bx = 1.f - t, by = t * sy; // warp.h:11-15
        V3 p = p0 + e0 * bx + e1 * by;
        V3 ns = normalize(cross(e0, e1));
        if (mi.flags & 1u) {
            float4 a1 = __ldg(vp + 2 * (size_t) i0 + 1), b1 = __ldg(vp + 2 * (size_t) i1 + 1), c1v = __ldg(vp + 2 * (size_t) i2 + 1);
            V3 n0 = v3(a0.w, a1.x, a1.y), n1 = v3(b0.w, b1.x, b1.y), n2 = v3(c0v.w, c1v.x, c1v.y);
            ns = normalize(n0 * (1.f - bx - by) + n1 * bx + n2 * by);
        }
        V3 d = p - ref_p;
        float dist2 = dot(d, d);
        float dist = sqrtf(dist2);
        d = v3(d.x / dist, d.y / dist, d.z / dist);
        float dp = fabsf(dot(d, ns));
        float pdf = mi.inv_area * ((dp != 0.f) ? dist2 / dp : 0.f);
        r.d = d; r.dist = dist;
        int radiance = em.radiance;
        if (sc.has_textures) { // ps.uv, mesh.cpp:114-119: the warped sample, or the interpolated texcoords
            float u = bx, v = by;
            if (mi.flags & 2u) {
                float4 a1 = __ldg(vp + 2 * (size_t) i0 + 1), b1 = __ldg(vp + 2 * (size_t) i1 + 1), c1v = __ldg(vp + 2 * (size_t) i2 + 1);
                float w0 = 1.f - bx - by;
                u = a1.z * w0 + b1.z * bx + c1v.z * by; v = a1.w * w0 + b1.w * bx + c1v.w * by;
            }
            radiance = texture_resolve(sc, radiance, u, v);
        }
        // pdf_emitter_direct(ds) of this record: shape.cpp:80-86
        r.stale_pdf = mi.inv_area * ((dp != 0.f) ? (dist * dist) / dp : 0.f) * (ne > 1 ? 1.f / (float) ne : 1.f);
        if (dot(d, ns) < 0.f && pdf != 0.f) {
            r.radiance = radiance; r.pdf0 = pdf;
            r.pdf = pdf;
        }
    } else { // constant.cpp:55-73
        V3 d = square_to_uniform_sphere(sx, sy);
        r.d = d; r.dist = 2.f * sc.env_radius; r.pdf = kInvFourPi;
        r.radiance = sc.has_textures ? texture_resolve(sc, em.radiance, 0.f, 0.f) : em.radiance; r.pdf0 = r.pdf;
        r.stale_pdf = kInvFourPi * (ne > 1 ? 1.f / (float) ne : 1.f);
    }
    if (ne > 1) { r.pdf *= sel; r.scale = (float) ne; }
    return r;
}

// Every spectrum one path vertex needs -- the BSDF's (by type), the radiance of the emitter that was hit (id_le) and of
// the emitter NEE sampled (id_ln); ids < 0 are skipped.  The loop is deliberately NOT unrolled: one copy of
// spectrum_eval's five-way switch per shade kernel.
template <int TYPE, bool UNROLL = false>
__device__ __forceinline__ void eval_vertex_spectra(const DScene &sc, const MskBsdf &b, bool need_bsdf, int id_le, int id_ln, float4 wl,
                                                    BsdfSpectra &sp, float4 &le, float4 &ln) {
    const int t = TYPE >= 0 ? TYPE : b.type;
    const bool dielectric = t == MSK_BSDF_ROUGHDIELECTRIC || t == MSK_BSDF_DIELECTRIC, conductor = t == MSK_BSDF_CONDUCTOR || t == MSK_BSDF_ROUGHCONDUCTOR;
    const int id_r = need_bsdf ? b.reflectance : -1, id_t = need_bsdf && dielectric ? b.transmittance : -1;
    const int id_e = need_bsdf && conductor ? b.eta : -1, id_k = need_bsdf && conductor ? b.k : -1;
    sp.R = sp.Tr = sp.eta = sp.k = le = ln = f4(0.f);
    if (UNROLL) { // k_shade_vol: its many live values spill around the loop (measured: 25.0 -> 27.0 ms on the vol workload)
        if (id_r >= 0) sp.R = spectrum_eval(sc, id_r, wl);
        if (id_t >= 0) sp.Tr = spectrum_eval(sc, id_t, wl);
        if (id_e >= 0) sp.eta = spectrum_eval(sc, id_e, wl);
        if (id_k >= 0) sp.k = spectrum_eval(sc, id_k, wl);
        if (id_le >= 0) le = spectrum_eval(sc, id_le, wl);
        if (id_ln >= 0) ln = spectrum_eval(sc, id_ln, wl);
        return;
    }
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        const int id = i == 0 ? id_r : (i == 1 ? id_t : (i == 2 ? id_e : (i == 3 ? id_k : (i == 4 ? id_le : id_ln))));
        if (id < 0) continue;
        const float4 v = spectrum_eval(sc, id, wl);
        if (i == 0) sp.R = v; else if (i == 1) sp.Tr = v; else if (i == 2) sp.eta = v; else if (i == 3) sp.k = v; else if (i == 4) le = v; else ln = v;
    }
}

} // namespace msk
