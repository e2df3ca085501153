// TEST INFRASTRUCTURE -- CPU oracle, not product code.  See oracle/README.md.
//
// Core math of the reference, restated without Eigen.  Each block cites the
// reference file:line it follows (paths relative to /root/reference).
//
// Parity status: the reference has no tests/golden vectors and cannot be built
// here (Eigen/pugixml/TBB/Embree/OIIO absent) => "parity unpinned" for
// everything except rgb2spec, which IS pinned against the compiled reference
// (oracle/_ref/librgb2spec_ref.so, tests/test_oracle_rgb2spec.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>

namespace orc {

// ---- constants: include/misaki/core/mathutils.h:10-20 ----
constexpr float Pi        = float(3.14159265358979323846);
constexpr float InvPi     = float(0.31830988618379067154);
constexpr float InvTwoPi  = float(0.15915494309189533577);
constexpr float InvFourPi = float(0.07957747154594766788);
constexpr float Infinity  = std::numeric_limits<float>::infinity();
constexpr float Epsilon   = std::numeric_limits<float>::epsilon() / 2;
constexpr float RayEpsilon    = Epsilon * 1500;
constexpr float ShadowEpsilon = RayEpsilon * 10;

inline float sqr(float a) { return a * a; }
inline float safe_sqrt(float a) { return std::sqrt(std::max(a, 0.f)); } // mathutils.h:52-54

struct V2 { float x, y; };
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator-(V3 a) { return { -a.x, -a.y, -a.z }; }
inline V3 operator*(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline V3 operator*(float s, V3 a) { return { a.x * s, a.y * s, a.z * s }; }
inline V3 operator/(V3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
// Eigen's dot() of a 3-vector is the plain left-to-right sum of products.
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
inline float squared_norm(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
// Eigen normalized(): v / sqrt(squaredNorm) when squaredNorm > 0
inline V3 normalized(V3 a) {
    float z = squared_norm(a);
    return z > 0.f ? a / std::sqrt(z) : a;
}
inline float max_abs_coeff(V3 a) { return std::max(std::abs(a.x), std::max(std::abs(a.y), std::abs(a.z))); }

// 4 hero wavelengths: include/misaki/core/fwd.h:40-41 (Spectrum = SpectrumArray<float,4>)
struct Spec {
    float v[4];
    Spec() : v{ 0, 0, 0, 0 } {}
    explicit Spec(float c) : v{ c, c, c, c } {}
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float max_coeff() const { return std::max(std::max(v[0], v[1]), std::max(v[2], v[3])); }
    bool is_zero() const { return v[0] == 0.f && v[1] == 0.f && v[2] == 0.f && v[3] == 0.f; }
};
#define ORC_SPEC_OP(op)                                                                             \
    inline Spec operator op(const Spec &a, const Spec &b) {                                        \
        Spec r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] op b.v[i]; return r; }                 \
    inline Spec operator op(const Spec &a, float b) {                                              \
        Spec r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] op b; return r; }                      \
    inline Spec operator op(float a, const Spec &b) {                                              \
        Spec r; for (int i = 0; i < 4; ++i) r.v[i] = a op b.v[i]; return r; }
ORC_SPEC_OP(+) ORC_SPEC_OP(-) ORC_SPEC_OP(*) ORC_SPEC_OP(/)
#undef ORC_SPEC_OP
inline Spec &operator+=(Spec &a, const Spec &b) { a = a + b; return a; }
inline Spec &operator*=(Spec &a, const Spec &b) { a = a * b; return a; }
inline Spec &operator*=(Spec &a, float b) { a = a * b; return a; }
inline Spec &operator/=(Spec &a, float b) { a = a / b; return a; }
inline Spec sqrt(const Spec &a) { Spec r; for (int i = 0; i < 4; ++i) r.v[i] = std::sqrt(a.v[i]); return r; }

// ---- PCG32: include/misaki/core/mathutils.h:85-143 ----
constexpr uint64_t PCG32_DEFAULT_STATE  = 0x853c49e6748fea9bULL;
constexpr uint64_t PCG32_DEFAULT_STREAM = 0xda3e39cb94b95bdbULL;
constexpr uint64_t PCG32_MULT           = 0x5851f42d4c957f2dULL;

struct PCG32 {
    uint64_t state = PCG32_DEFAULT_STATE, inc = PCG32_DEFAULT_STREAM;
    void seed(uint64_t initstate, uint64_t initseq = 1) { // :95-101
        state = 0U;
        inc   = (initseq << 1u) | 1u;
        next_uint32();
        state += initstate;
        next_uint32();
    }
    uint32_t next_uint32() { // :102-109
        uint64_t oldstate   = state;
        state               = oldstate * PCG32_MULT + inc;
        uint32_t xorshifted = (uint32_t) (((oldstate >> 18u) ^ oldstate) >> 27u);
        uint32_t rot        = (uint32_t) (oldstate >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    float next_float32() { // :111-120
        uint32_t u = (next_uint32() >> 9) | 0x3f800000u;
        float f;
        std::memcpy(&f, &u, 4);
        return f - 1.0f;
    }
};

// ---- IndependentSampler: src/librender/samplers/independent.cpp:20-35 ----
struct Sampler {
    PCG32 rng;
    uint64_t base_seed = 0;
    void seed(uint64_t seed_value) { rng.seed(seed_value + base_seed, PCG32_DEFAULT_STREAM); }
    float next1d() { return rng.next_float32(); }
    V2 next2d() { float a = next1d(); float b = next1d(); return { a, b }; } // braced init: left-to-right
};

// ---- coordinate_system: mathutils.h:196-203 ; Frame: frame.h:11-40 ----
inline void coordinate_system(V3 n, V3 &s, V3 &t) {
    float sign    = std::copysign(1.f, n.z);
    const float a = -1.f / (sign + n.z);
    const float b = n.x * n.y * a;
    s = { 1.f + sign * n.x * n.x * a, sign * b, -sign * n.x };
    t = { b, sign + n.y * n.y * a, -n.y };
}
struct Frame {
    V3 s, t, n;
    Frame() {}
    explicit Frame(V3 v) : n(v) { coordinate_system(v, s, t); }
    V3 to_local(V3 v) const { return { dot(v, s), dot(v, t), dot(v, n) }; }
    V3 to_world(V3 v) const { return s * v.x + t * v.y + n * v.z; }
    static float cos_theta(V3 v) { return v.z; }
    static float cos_theta_2(V3 v) { return sqr(v.z); }
};

// ---- warps: include/misaki/core/warp.h:11-63 ----
inline V2 square_to_uniform_triangle(V2 s) {
    float t = safe_sqrt(1.f - s.x);
    return { 1.f - t, t * s.y };
}
inline V2 square_to_uniform_disk_concentric(V2 s) {
    float x = 2.f * s.x - 1.f;
    float y = 2.f * s.y - 1.f;
    float phi, r;
    if (x == 0 && y == 0) {
        r = phi = 0;
    } else if (x * x > y * y) {
        r   = x;
        phi = (Pi / 4.f) * (y / x);
    } else {
        r   = y;
        phi = (Pi / 2.f) - (x / y) * (Pi / 4.f);
    }
    return { r * std::cos(phi), r * std::sin(phi) };
}
inline V3 square_to_cosine_hemisphere(V2 s) {
    V2 p    = square_to_uniform_disk_concentric(s);
    float z = safe_sqrt(1.f - (p.x * p.x + p.y * p.y));
    return { p.x, p.y, z };
}
inline float square_to_cosine_hemisphere_pdf(V3 v) { return InvPi * v.z; }
inline V3 square_to_uniform_sphere(V2 s) {
    float z = -2.f * s.y + 1.f, r = safe_sqrt(-z * z + 1.f);
    float t = 2.f * Pi * s.x;
    float sn = std::sin(t), c = std::cos(t);
    return { r * c, r * sn, z };
}

// ---- wavelength sampling: include/misaki/core/spectrum.h:152-181, mathutils.h:167-182 ----
inline void sample_wavelength(float sample, Spec &wavelengths, Spec &weight) {
    for (int i = 0; i < 4; ++i) {
        float value = sample + float(i) / 4.f;            // shift = Index / Size
        float u     = (value <= 1.f) ? value : value - 1.f; // (value <= one).select(value, value-one)
        float lam   = 538.f - std::atanh(0.8569106254698279f - 1.8275019724092267f * u) * 138.88888888888889f;
        float tmp   = std::cosh(0.0072f * (lam - 538.f));
        wavelengths[i] = lam;
        weight[i]      = 253.82f * tmp * tmp;
    }
}

// ---- Fresnel: include/misaki/render/fresnel.h ----
inline V3 reflect(V3 wi) { return { -wi.x, -wi.y, wi.z }; }                       // :11-14
inline V3 reflect(V3 wi, V3 n) { return n * 2.f * dot(wi, n) - wi; }              // :16-20
inline V3 refract(V3 wi, float cos_theta_t, float eta_ti) {                       // :22-27
    return { -eta_ti * wi.x, -eta_ti * wi.y, cos_theta_t };
}
inline V3 refract(V3 wi, V3 m, float cos_theta_t, float eta_ti) {                 // :29-34
    return m * (dot(wi, m) * eta_ti + cos_theta_t) - wi * eta_ti;
}
struct FresnelResult { float F, cos_theta_t, eta_it, eta_ti; };
inline FresnelResult fresnel(float cos_theta_i, float eta) {                      // :37-63
    float eta_it, eta_ti;
    if (cos_theta_i >= 0.f) { eta_it = eta; eta_ti = 1.f / eta; }
    else                    { eta_it = 1.f / eta; eta_ti = eta; }
    float cos_theta_t_sqr = 1.f - eta_ti * eta_ti * (1.f - cos_theta_i * cos_theta_i);
    float cos_theta_i_abs = std::abs(cos_theta_i);
    float cos_theta_t_abs = safe_sqrt(cos_theta_t_sqr);
    float a_s = (cos_theta_i_abs - eta_it * cos_theta_t_abs) / (cos_theta_i_abs + eta_it * cos_theta_t_abs);
    float a_p = (cos_theta_t_abs - eta_it * cos_theta_i_abs) / (cos_theta_t_abs + eta_it * cos_theta_i_abs);
    float r;
    if (eta == 1.f || cos_theta_i_abs == 0.f)
        r = eta == 1.f ? 0.f : 1.f;
    else
        r = 0.5f * (a_s * a_s + a_p * a_p);
    float cos_theta_t = cos_theta_t_abs * std::copysign(1.f, -cos_theta_i);
    return { r, cos_theta_t, eta_it, eta_ti };
}
// fresnel.h:65-88, with Color3 -> Spec (the reference's RGB form evaluated per wavelength)
inline Spec fresnel_conductor(float cos_theta_i, const Spec &eta, const Spec &k) {
    float cos_theta_i_2 = cos_theta_i * cos_theta_i, sin_theta_i_2 = 1.f - cos_theta_i_2,
          sin_theta_i_4 = sin_theta_i_2 * sin_theta_i_2;
    Spec eta_r = eta, eta_i = k;
    Spec temp_1   = eta_r * eta_r - eta_i * eta_i - sin_theta_i_2;
    Spec a_2_pb_2 = sqrt(temp_1 * temp_1 + 4.f * eta_i * eta_i * eta_r * eta_r);
    Spec a        = sqrt(.5f * (a_2_pb_2 + temp_1));
    Spec term_1 = a_2_pb_2 + cos_theta_i_2, term_2 = 2.f * cos_theta_i * a;
    Spec r_s    = (term_1 - term_2) / (term_1 + term_2);
    Spec term_3 = a_2_pb_2 * cos_theta_i_2 + sin_theta_i_4, term_4 = term_2 * sin_theta_i_2;
    Spec r_p    = r_s * (term_3 - term_4) / (term_3 + term_4);
    return .5f * (r_s + r_p);
}

// ---- GGX microfacet distribution: include/misaki/render/microfacet.h ----
// Only Type::GGX is defined by the reference (Beckmann eval()==0 / sample() falls
// off the end, :113-115,:134-136), so only GGX is restated.
struct Microfacet {
    float alpha_u, alpha_v;
    Microfacet(float au, float av) : alpha_u(std::max(au, 1e-4f)), alpha_v(std::max(av, 1e-4f)) {} // :190-193
    void scale_alpha(float v) { alpha_u *= v; alpha_v *= v; }                                      // :184-187
    static float eval_ggx(V3 m, float au, float av) {                                              // :11-18
        float cos_theta2  = Frame::cos_theta_2(m);
        float beckman_exp = ((m.x * m.x / (au * au)) + (m.y * m.y) / (av * av)) / cos_theta2;
        float root        = (1.f + beckman_exp) * cos_theta2;
        return 1.f / (Pi * au * av * root * root);
    }
    float eval(V3 m) const {                                                                       // :108-125
        if (Frame::cos_theta(m) <= 0) return 0.0f;
        float result = eval_ggx(m, alpha_u, alpha_v);
        return result * Frame::cos_theta(m) > 1e-20f ? result : 0.f;
    }
    float pdf(V3 /*wi*/, V3 m) const { return eval(m) * Frame::cos_theta(m); }                     // :127-129
    std::pair<V3, float> sample(V3 /*wi*/, V2 s) const {                                           // :20-40
        float phi_m = std::atan(alpha_u / alpha_v * std::tan(Pi + 2 * Pi * s.y)) + Pi * std::floor(2 * s.y + 0.5f);
        float sin_phi_m = std::sin(phi_m), cos_phi_m = std::cos(phi_m);
        float c = cos_phi_m / alpha_u, sn = sin_phi_m / alpha_v;
        float alpha_sqr       = 1.f / (c * c + sn * sn);
        float tan_theta_m_sqr = alpha_sqr * s.x / (1.f - s.x);
        float cos_theta_m     = 1.f / std::sqrt(1.f + tan_theta_m_sqr);
        float tmp             = 1 + tan_theta_m_sqr / alpha_sqr;
        float pdf = InvPi / (alpha_u * alpha_v * cos_theta_m * cos_theta_m * cos_theta_m * tmp * tmp);
        if (pdf < 1e-20f) pdf = 0;
        float sin_theta_m = safe_sqrt(1 - cos_theta_m * cos_theta_m);
        return { V3(sin_theta_m * cos_phi_m, sin_theta_m * sin_phi_m, cos_theta_m), pdf };
    }
    float smith_g1(V3 v, V3 m) const {                                                             // :150-175
        float xy_alpha_2 = sqr(alpha_u * v.x) + sqr(alpha_v * v.y), tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z);
        if (xy_alpha_2 == 0.f) return 1.f;
        if (dot(v, m) * Frame::cos_theta(v) <= 0.f) return 0.f;
        return 2.f / (1.f + std::sqrt(1.f + tan_theta_alpha_2));
    }
    float G(V3 wi, V3 wo, V3 m) const { return smith_g1(wi, m) * smith_g1(wo, m); }                // :145-148
};

} // namespace orc
