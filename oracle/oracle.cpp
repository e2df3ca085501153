// TEST INFRASTRUCTURE -- CPU oracle, not product code.
//
// A CPU restatement of the path-tracing hot path of jczh98/misaki-render
// (SURVEY.md section 8a), written without Eigen/Embree/TBB.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (misaki_render_b200) never does.
//
// PARITY: PARTLY PINNED.  The reference ships no tests, golden vectors or fixtures and
// cannot be built as a whole in this image (Eigen3, pugixml, TBB, Embree 3.12.2,
// OpenImageIO absent).  What IS checked against outputs of the reference itself, compiled
// here by oracle/Makefile.ref from the sources where they lie:
//   * rgb2spec_fetch / srgb.coeff            (ext/rgb2spec; tests/test_oracle_rgb2spec.py)
//   * the math layer, bit for bit            (include/misaki/core/{mathutils,warp,frame,spectrum,distribution}.h,
//     include/misaki/render/{fresnel,microfacet,srgb}.h, src/librender/spectrum.cpp against a minimal Eigen stand-in,
//     oracle/ref_shim: PCG32 + sampler floats, the warps, coordinate_system / Frame, Fresnel, reflect / refract, GGX
//     eval / pdf / sample / G / smith_g1, sample_wavelength, spectrum_to_xyz, xyz_to_srgb, srgb_model_eval,
//     Distribution1D; tests/golden/ref_math.json, tests/test_oracle_ref_math.py)
//   * plugin sources the reference's build compiles, over stand-ins for its object system: bsdfs/diffuse.cpp, rfilter.cpp +
//     filters/gaussian.cpp, sampler.cpp + samplers/independent.cpp, spectra/regular.cpp, spectra/uniform.cpp, and
//     mesh.cpp + shape.cpp + records.cpp + interaction.cpp (hit reconstruction, area distribution, position / direct
//     sampling) (same tests)
//   * PathTracer::sample itself (integrators/path.cpp) with scene.cpp's emitter sampling, emitters/area.cpp, emitter.cpp,
//     emitters/constant.cpp, bsdf.cpp: 800 paths through two scenes bit for bit (Embree's calls restated around the same
//     brute-force intersector); AOVIntegrator::sample (integrators/aov.cpp); imageblock.cpp (splat, block merge, spiral):
//     whole films equal by SHA-256; srgb.cpp + spectra/{srgb,srgb_d65,d65}.cpp (what an <rgb> tag becomes)
//   * SamplingIntegrator::render / render_block / render_sample themselves (src/librender/integrator.cpp, serial stand-in for
//     tbb::parallel_for, camera rays as inputs): three whole films -- one of them config C1's Cornell box with the reference's
//     own <rgb> textures -- bit for bit against render() below in its test-only
//     ORC_RENDER_REFERENCE_SEEDING mode
//   * Film / HDRFilm (film.cpp, films/hdrfilm.cpp): prepare / put inside that render loop, and HDRFilm::image (the develop
//     step) bit for bit against orc_develop and the product's host develop
// UNPINNED (restated from the cited lines, checked by known-answer tests only): volpath.cpp, the BSDF plugins other than
// diffuse (their sources are stale-API and compile with no Eigen), the camera -- and Embree's arithmetic.  Every function cites the reference file:line it follows; paths are relative to /root/reference.
//
// Third-party arithmetic outside the reference tree: Embree 3.12.2 (vcpkg port
// embree3, vcpkg/ports/embree3/vcpkg.json).  Its default triangle intersector
// (Moeller-Trumbore, kernels/geometry/triangle_intersector_moeller.h) is
// restated in tri_intersect() from its published algorithm; the ground truth for
// closest hits is a brute-force loop over all triangles with that test.
//
// Determinism contract added on top of the reference: the
// reference never seeds per pixel (integrator.cpp:57, independent.cpp:9-12), so
// its output depends on TBB scheduling.  Here pixel p = y*W + x, sample s uses
// Sampler::seed(p*spp + s) (independent.cpp:20-26) and draws in source order.
// XML max_depth / rr_depth / hide_emitters are honoured (the reference shadows
// them, path.cpp:135-136).
#include "oracle.h"
#include "oracle_math.h"
#include "../misaki_render_b200/csrc/spectral_tables.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

using namespace orc;

namespace {

thread_local std::string g_error;
int fail(const std::string &msg) { g_error = msg; return -1; }

// ------------------------------------------------------------------------------------------
// Spectra: src/librender/spectra/*.cpp, include/misaki/render/srgb.h
// ------------------------------------------------------------------------------------------
struct OSpectrum {
    int kind;
    float c[3];
    float value;
    std::vector<float> table;
    float lambda_min, lambda_max;
    float inv_interval_size; // regular.cpp:58 (double 1/interval stored to float)
    int child0 = -1, child1 = -1; // checkerboard.cpp:12-13 "color0" / "color1"
    float to_uv[6] = { 1, 0, 0, 0, 1, 0 }; // checkerboard.cpp:14: Transform4f::extract() = top-left 3x3, rows 0 and 1
};

// srgb.h:8-19
Spec srgb_model_eval(const float c[3], const Spec &wl) {
    if (std::isinf(c[2]))
        return Spec(std::copysign(1.f, c[2]) * .5f + .5f);
    Spec r;
    for (int i = 0; i < 4; ++i) {
        float v     = (c[0] * wl[i] + c[1]) * wl[i] + c[2];
        float rsqrt = 1.f / std::sqrt(v * v + 1.f);
        r[i]        = std::max(.5f * v * rsqrt + .5f, 0.f);
    }
    return r;
}

// regular.cpp:73-91 (SpectrumContinuousDistribution::eval_pdf)
Spec regular_eval(const OSpectrum &s, const Spec &wl) {
    Spec r;
    uint32_t last = (uint32_t) s.table.size() - 2;
    for (int i = 0; i < 4; ++i) {
        float x      = (wl[i] - s.lambda_min) * s.inv_interval_size;
        uint32_t idx = std::min((uint32_t) x, last);
        float y0 = s.table[idx], y1 = s.table[idx + 1];
        float w1 = x - (float) idx, w0 = 1.f - w1;
        r[i] = w0 * y0 + w1 * y1;
    }
    return r;
}

Spec spectrum_eval(const OSpectrum &s, const Spec &wl) {
    switch (s.kind) {
        case MSK_SPEC_UNIFORM: { // uniform.cpp:19-26: all four wavelengths must be in range
            bool in = true;
            for (int i = 0; i < 4; ++i) in &= (wl[i] >= 360.f) && (wl[i] <= 830.f);
            return in ? Spec(s.value) : Spec(0.f);
        }
        case MSK_SPEC_SRGB: return srgb_model_eval(s.c, wl);                              // srgb.cpp:18-20
        case MSK_SPEC_SRGB_D65: return regular_eval(s, wl) * srgb_model_eval(s.c, wl);    // srgb_d65.cpp:35-37
        case MSK_SPEC_REGULAR: return regular_eval(s, wl);
        case MSK_SPEC_SRGB_UNBOUNDED: return srgb_model_eval(s.c, wl) * s.value;          // builder decision (F4)
    }
    return Spec(0.f);
}

// textures/checkerboard.cpp:25-33 (Texture::eval(si)): Transform3f::transform_affine_point (transform.h:41-48) of
// si.uv, fractional parts, then color0 where both or neither exceed one half, else color1
Spec texture_eval(const std::vector<OSpectrum> &spectra, int id, float u, float v, const Spec &wl) {
    const OSpectrum &s = spectra[id];
    if (s.kind != MSK_SPEC_CHECKERBOARD) return spectrum_eval(s, wl);
    float tu = (s.to_uv[0] * u + s.to_uv[1] * v) + s.to_uv[2] * 1.f;
    float tv = (s.to_uv[3] * u + s.to_uv[4] * v) + s.to_uv[5] * 1.f;
    tu = tu - std::floor(tu);
    tv = tv - std::floor(tv);
    return texture_eval(spectra, ((tu > .5f) == (tv > .5f)) ? s.child0 : s.child1, u, v, wl);
}

// ------------------------------------------------------------------------------------------
// CIE 1931 / XYZ: include/misaki/core/spectrum.h:83-143
// ------------------------------------------------------------------------------------------
void spectrum_to_xyz(const Spec &value, const Spec &wl, float xyz[3]) {
    Spec X, Y, Z;
    for (int s = 0; s < 4; ++s) {
        float t     = (wl[s] - MSK_CIE_MIN) * ((MSK_CIE_SAMPLES - 1) / (MSK_CIE_MAX - MSK_CIE_MIN));
        uint32_t i0 = std::min(std::max((uint32_t) t, 0u), (uint32_t) (MSK_CIE_SAMPLES - 2)), i1 = i0 + 1;
        float w1 = t - float(i0), w0 = 1.f - w1;
        X[s] = (w0 * msk_cie_d65_rows[i0][0] + w1 * msk_cie_d65_rows[i1][0]) * value[s];
        Y[s] = (w0 * msk_cie_d65_rows[i0][1] + w1 * msk_cie_d65_rows[i1][1]) * value[s];
        Z[s] = (w0 * msk_cie_d65_rows[i0][2] + w1 * msk_cie_d65_rows[i1][2]) * value[s];
    }
    // Eigen's vectorised 4-wide reduction adds (v0+v2)+(v1+v3); mean() = sum / 4
    xyz[0] = ((X[0] + X[2]) + (X[1] + X[3])) / 4.f;
    xyz[1] = ((Y[0] + Y[2]) + (Y[1] + Y[3])) / 4.f;
    xyz[2] = ((Z[0] + Z[2]) + (Z[1] + Z[3])) / 4.f;
}

// ------------------------------------------------------------------------------------------
// Scene data
// ------------------------------------------------------------------------------------------
struct OMesh {
    std::vector<float> verts; // 8 floats per vertex (obj.cpp:139-142)
    std::vector<uint32_t> tris;
    uint32_t nverts = 0, ntris = 0;
    int bsdf = 0, emitter = -1;
    bool has_normals = false, has_uvs = false;
    int interior_medium = -1, exterior_medium = -1; // shape.cpp:28-39
    std::vector<float> cdf; // Distribution1D::m_cdf, distribution.h:84-93
    float surface_area = 0.f;
    V3 pos(uint32_t i) const { return { verts[i * 8], verts[i * 8 + 1], verts[i * 8 + 2] }; }
    V3 nrm(uint32_t i) const { return { verts[i * 8 + 3], verts[i * 8 + 4], verts[i * 8 + 5] }; }
    V2 uv(uint32_t i) const { return { verts[i * 8 + 6], verts[i * 8 + 7] }; }
};

struct BVHNode { // oracle-private SAH BVH2 (speed only; results equal the brute-force loop)
    float lo[3], hi[3];
    uint32_t left, right; // children, or [first, count] when leaf
    uint32_t first, count;
};

struct TriRef { uint32_t geom, prim; };

struct OScene {
    std::vector<OMesh> meshes;
    std::vector<MskBsdf> bsdfs;
    std::vector<MskEmitter> emitters;
    std::vector<OSpectrum> spectra;
    int environment = -1;
    std::vector<MskMedium> media; // media/homogeneous.cpp
    int sensor_medium = -1;       // sensor.cpp:12-18
    MskCamera cam;
    V3 bbox_min, bbox_max;
    float env_radius = 0.f; // constant.cpp:21-28
    std::vector<BVHNode> nodes;
    std::vector<TriRef> refs;
};

struct Ray { V3 o, d; float mint, maxt; Spec wavelengths; };

// interaction.h:8-108 (fields the compiled path actually consumes)
struct SceneInteraction {
    float t = Infinity;
    V3 p, n;
    Frame sh_frame;
    V2 uv;
    V3 wi;
    V3 dp_du, dp_dv;
    uint32_t prim_index = 0;
    int shape = -1;
    Spec wavelengths;
    bool is_valid() const { return t != Infinity; }
    V3 to_world(V3 v) const { return sh_frame.to_world(v); }
    V3 to_local(V3 v) const { return sh_frame.to_local(v); }
    Ray spawn_ray(V3 d) const { // interaction.h:40-44
        return { p, d, (1.f + max_abs_coeff(p)) * RayEpsilon, Infinity, wavelengths };
    }
};

// records.h:8-40
struct DirectIllumSample {
    V3 p, n;
    V2 uv;
    float pdf = 0.f;
    int object = -1; // emitter index
    V3 d;
    float dist = 0.f;
};

struct RayCounters { uint64_t closest = 0, shadow = 0; };

// ------------------------------------------------------------------------------------------
// Ray/triangle: Embree 3.12.2 Moeller-Trumbore (third-party, restated from the published
// algorithm: kernels/geometry/triangle_intersector_moeller.h, MoellerTrumboreIntersector1)
//   e1 = v0-v1, e2 = v2-v0, Ng = e2 x e1, C = v0-o, R = C x d, den = Ng.d
//   U = (R.e2)*sgn(den), V = (R.e1)*sgn(den), T = (Ng.C)*sgn(den)
//   hit iff den != 0, U >= 0, V >= 0, U+V <= |den|, |den|*tnear < T <= |den|*tfar
//   u = U/|den|, v = V/|den|, t = T/|den|   (hit point = (1-u-v) v0 + u v1 + v v2)
// ------------------------------------------------------------------------------------------
inline bool tri_intersect(V3 v0, V3 v1, V3 v2, V3 o, V3 d, float tnear, float tfar, float &t, float &u, float &v) {
    V3 e1 = v0 - v1, e2 = v2 - v0, Ng = cross(e2, e1);
    V3 C = v0 - o, R = cross(C, d);
    float den = dot(Ng, d);
    if (den == 0.f) return false;
    float absden = std::abs(den), sgn = den < 0.f ? -1.f : 1.f;
    float U = dot(R, e2) * sgn, V = dot(R, e1) * sgn;
    if (!(U >= 0.f) || !(V >= 0.f) || !(U + V <= absden)) return false;
    float T = dot(Ng, C) * sgn;
    if (!(absden * tnear < T) || !(T <= absden * tfar)) return false;
    float rcp = 1.f / absden;
    u = U * rcp; v = V * rcp; t = T * rcp;
    return true;
}

struct RawHit { float t = Infinity, u = 0, v = 0; uint32_t prim = 0xffffffffu, geom = 0xffffffffu; };

// closest hit; ties in t resolved towards the lower (geom, prim) so that the BVH and the
// brute-force loop return identical records
inline void consider(const OScene &sc, uint32_t geom, uint32_t prim, const Ray &ray, RawHit &best) {
    const OMesh &m = sc.meshes[geom];
    const uint32_t *f = &m.tris[prim * 3];
    float t, u, v;
    if (tri_intersect(m.pos(f[0]), m.pos(f[1]), m.pos(f[2]), ray.o, ray.d, ray.mint, ray.maxt, t, u, v)) {
        bool better = t < best.t || (t == best.t && (geom < best.geom || (geom == best.geom && prim < best.prim)));
        if (better) { best.t = t; best.u = u; best.v = v; best.prim = prim; best.geom = geom; }
    }
}

RawHit intersect_brute(const OScene &sc, const Ray &ray) {
    RawHit best;
    for (uint32_t g = 0; g < sc.meshes.size(); ++g)
        for (uint32_t p = 0; p < sc.meshes[g].ntris; ++p) consider(sc, g, p, ray, best);
    return best;
}

inline bool slab(const BVHNode &n, const Ray &ray, const float inv[3], float tmax) {
    float t0 = ray.mint, t1 = tmax;
    const float o[3] = { ray.o.x, ray.o.y, ray.o.z };
    for (int a = 0; a < 3; ++a) {
        float ta = (n.lo[a] - o[a]) * inv[a], tb = (n.hi[a] - o[a]) * inv[a];
        if (ta > tb) std::swap(ta, tb);
        // NaN-safe, conservative: widen by 2 ulp-ish so the BVH never culls a true MT hit
        t0 = std::max(t0, ta * (1.f - 4e-7f) - 1e-30f);
        t1 = std::min(t1, tb * (1.f + 4e-7f) + 1e-30f);
    }
    return t0 <= t1;
}

RawHit intersect_bvh(const OScene &sc, const Ray &ray, bool any_hit) {
    RawHit best;
    if (sc.nodes.empty()) return best;
    float inv[3] = { 1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z };
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const BVHNode &n = sc.nodes[stack[--sp]];
        // use <= best.t (not <) so equal-t candidates are still visited for the tie rule
        if (!slab(n, ray, inv, std::min(ray.maxt, best.t))) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; ++i) {
                const TriRef &r = sc.refs[n.first + i];
                consider(sc, r.geom, r.prim, ray, best);
                if (any_hit && best.t != Infinity) return best;
            }
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.right;
        }
    }
    return best;
}

// binned SAH BVH2 build (oracle-private)
void build_bvh(OScene &sc) {
    sc.refs.clear(); sc.nodes.clear();
    std::vector<float> clo, chi;
    for (uint32_t g = 0; g < sc.meshes.size(); ++g) {
        const OMesh &m = sc.meshes[g];
        for (uint32_t p = 0; p < m.ntris; ++p) sc.refs.push_back({ g, p });
    }
    size_t n = sc.refs.size();
    if (!n) return;
    std::vector<float> blo(n * 3), bhi(n * 3), cen(n * 3);
    for (size_t i = 0; i < n; ++i) {
        const OMesh &m = sc.meshes[sc.refs[i].geom];
        const uint32_t *f = &m.tris[sc.refs[i].prim * 3];
        V3 v[3] = { m.pos(f[0]), m.pos(f[1]), m.pos(f[2]) };
        for (int a = 0; a < 3; ++a) {
            float lo = std::min(v[0][a], std::min(v[1][a], v[2][a])), hi = std::max(v[0][a], std::max(v[1][a], v[2][a]));
            blo[i * 3 + a] = lo; bhi[i * 3 + a] = hi; cen[i * 3 + a] = 0.5f * (lo + hi);
        }
    }
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    struct Task { uint32_t node, first, count; };
    std::vector<Task> todo;
    sc.nodes.reserve(2 * n);
    sc.nodes.push_back({});
    todo.push_back({ 0, 0, (uint32_t) n });
    auto area = [](const float lo[3], const float hi[3]) {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.f * (dx * dy + dy * dz + dz * dx);
    };
    while (!todo.empty()) {
        Task t = todo.back(); todo.pop_back();
        float lo[3] = { Infinity, Infinity, Infinity }, hi[3] = { -Infinity, -Infinity, -Infinity };
        float clo3[3] = { Infinity, Infinity, Infinity }, chi3[3] = { -Infinity, -Infinity, -Infinity };
        for (uint32_t i = 0; i < t.count; ++i) {
            uint32_t r = idx[t.first + i];
            for (int a = 0; a < 3; ++a) {
                lo[a] = std::min(lo[a], blo[r * 3 + a]); hi[a] = std::max(hi[a], bhi[r * 3 + a]);
                clo3[a] = std::min(clo3[a], cen[r * 3 + a]); chi3[a] = std::max(chi3[a], cen[r * 3 + a]);
            }
        }
        BVHNode &nd = sc.nodes[t.node];
        for (int a = 0; a < 3; ++a) { nd.lo[a] = lo[a]; nd.hi[a] = hi[a]; }
        nd.left = nd.right = 0; nd.first = t.first; nd.count = t.count;
        if (t.count <= 2) continue;
        constexpr int NB = 16;
        float best_cost = Infinity; int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            float ext = chi3[a] - clo3[a];
            if (!(ext > 0.f)) continue;
            float bl[NB][3], bh[NB][3]; uint32_t bc[NB] = {};
            for (int b = 0; b < NB; ++b) for (int k = 0; k < 3; ++k) { bl[b][k] = Infinity; bh[b][k] = -Infinity; }
            float scale = NB / ext;
            for (uint32_t i = 0; i < t.count; ++i) {
                uint32_t r = idx[t.first + i];
                int b = std::min(NB - 1, (int) ((cen[r * 3 + a] - clo3[a]) * scale));
                bc[b]++;
                for (int k = 0; k < 3; ++k) { bl[b][k] = std::min(bl[b][k], blo[r * 3 + k]); bh[b][k] = std::max(bh[b][k], bhi[r * 3 + k]); }
            }
            float ra[NB]; uint32_t rc[NB];
            float al[3] = { Infinity, Infinity, Infinity }, ah[3] = { -Infinity, -Infinity, -Infinity }; uint32_t c = 0;
            for (int b = NB - 1; b > 0; --b) {
                for (int k = 0; k < 3; ++k) { al[k] = std::min(al[k], bl[b][k]); ah[k] = std::max(ah[k], bh[b][k]); }
                c += bc[b]; ra[b] = c ? area(al, ah) : 0.f; rc[b] = c;
            }
            float ll[3] = { Infinity, Infinity, Infinity }, lh[3] = { -Infinity, -Infinity, -Infinity }; c = 0;
            for (int b = 0; b < NB - 1; ++b) {
                for (int k = 0; k < 3; ++k) { ll[k] = std::min(ll[k], bl[b][k]); lh[k] = std::max(lh[k], bh[b][k]); }
                c += bc[b];
                if (!c || !rc[b + 1]) continue;
                float cost = area(ll, lh) * c + ra[b + 1] * rc[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
            }
        }
        uint32_t mid;
        if (best_axis < 0) {
            if (t.count <= 8) continue;
            mid = t.first + t.count / 2; // coincident centroids: split the list
        } else {
            float leaf_cost = area(lo, hi) * t.count;
            if (t.count <= 4 && best_cost >= leaf_cost) continue;
            float scale = NB / (chi3[best_axis] - clo3[best_axis]);
            auto it = std::partition(idx.begin() + t.first, idx.begin() + t.first + t.count, [&](uint32_t r) {
                int b = std::min(NB - 1, (int) ((cen[r * 3 + best_axis] - clo3[best_axis]) * scale));
                return b <= best_bin;
            });
            mid = (uint32_t) (it - idx.begin());
            if (mid == t.first || mid == t.first + t.count) mid = t.first + t.count / 2;
        }
        uint32_t l = (uint32_t) sc.nodes.size();
        sc.nodes.push_back({}); sc.nodes.push_back({});
        BVHNode &nd2 = sc.nodes[t.node];
        nd2.left = l; nd2.right = l + 1; nd2.count = 0;
        todo.push_back({ l, t.first, mid - t.first });
        todo.push_back({ l + 1, mid, t.first + t.count - mid });
    }
    std::vector<TriRef> sorted(n);
    for (size_t i = 0; i < n; ++i) sorted[i] = sc.refs[idx[i]];
    sc.refs.swap(sorted);
}

// ------------------------------------------------------------------------------------------
// Mesh: src/librender/mesh.cpp
// ------------------------------------------------------------------------------------------
// Distribution1D::init, distribution.h:88-96
std::vector<float> distribution_cdf(const float *table, size_t n) {
    std::vector<float> cdf(1, 0.f);
    float acc = 0.f; // std::partial_sum accumulates in float
    for (size_t i = 0; i < n; ++i) {
        acc = (i == 0) ? table[0] : acc + table[i];
        cdf.push_back(acc);
    }
    const float inv_sum = 1.f / cdf.back();
    for (auto &c : cdf) c *= inv_sum;
    return cdf;
}
// Distribution1D::sample_reuse, distribution.h:107-123
inline std::pair<uint32_t, float> distribution_sample_reuse(const std::vector<float> &cdf, float u) {
    auto it   = std::upper_bound(cdf.begin(), cdf.end(), u);
    int index = std::min(std::max(int(it - cdf.begin()) - 1, 0), int(cdf.size()) - 2);
    float pmf = cdf[index + 1] - cdf[index];
    return { (uint32_t) index, (u - cdf[index]) / pmf };
}

void area_distr_build(OMesh &m) { // mesh.cpp:39-48 + distribution.h:84-93
    std::vector<float> table(m.ntris);
    m.surface_area = 0.f; // the reference never initialises m_surface_area (mesh.h:93); restated as 0
    for (uint32_t i = 0; i < m.ntris; ++i) {
        const uint32_t *f = &m.tris[i * 3];
        V3 p0 = m.pos(f[0]), p1 = m.pos(f[1]), p2 = m.pos(f[2]);
        float a = 0.5f * norm(cross(p1 - p0, p2 - p0)); // mesh.h:51-57
        m.surface_area += a;
        table[i] = a;
    }
    m.cdf = distribution_cdf(table.data(), table.size());
}

// mesh.cpp:51-101 + interaction.cpp:24-37 + interaction.h:55-60
SceneInteraction compute_scene_interaction(const OScene &sc, const Ray &ray, const RawHit &h) {
    SceneInteraction si;
    si.wavelengths = ray.wavelengths;
    if (h.t == Infinity) {
        si.t  = Infinity;
        si.wi = -ray.d;
        return si;
    }
    const OMesh &m = sc.meshes[h.geom];
    float b1 = h.u, b2 = h.v, b0 = 1.f - b1 - b2;
    const uint32_t *fi = &m.tris[h.prim * 3];
    V3 p0 = m.pos(fi[0]), p1 = m.pos(fi[1]), p2 = m.pos(fi[2]);
    V3 dp0 = p1 - p0, dp1 = p2 - p0;
    si.t  = h.t;
    si.p  = p0 * b0 + p1 * b1 + p2 * b2;
    si.n  = normalized(cross(dp0, dp1));
    si.uv = { h.u, h.v };
    coordinate_system(si.n, si.dp_du, si.dp_dv);
    if (m.has_uvs) {
        V2 uv0 = m.uv(fi[0]), uv1 = m.uv(fi[1]), uv2 = m.uv(fi[2]);
        si.uv  = { uv0.x * b0 + uv1.x * b1 + uv2.x * b2, uv0.y * b0 + uv1.y * b1 + uv2.y * b2 };
        V2 duv0 = { uv1.x - uv0.x, uv1.y - uv0.y }, duv1 = { uv2.x - uv0.x, uv2.y - uv0.y };
        float det = duv0.x * duv1.y - duv0.y * duv1.x, inv_det = 1.f / det;
        if (det != 0.f) {
            si.dp_du = (duv1.y * dp0 - duv0.y * dp1) * inv_det;
            si.dp_dv = (-duv1.x * dp0 + duv0.x * dp1) * inv_det;
        }
    }
    if (m.has_normals) {
        V3 n0 = m.nrm(fi[0]), n1 = m.nrm(fi[1]), n2 = m.nrm(fi[2]);
        si.sh_frame.n = normalized(n0 * b0 + n1 * b1 + n2 * b2);
        // dn_du/dn_dv (mesh.cpp:86-96) are computed by the reference but consumed by nothing
    } else {
        si.sh_frame.n = si.n;
    }
    si.prim_index = h.prim;
    si.shape      = (int) h.geom;
    // initialize_sh_frame, interaction.h:55-60
    V3 face_forward = -si.sh_frame.n * dot(si.sh_frame.n, si.dp_du) + si.dp_du;
    si.sh_frame.s   = normalized(face_forward);
    si.sh_frame.t   = cross(si.sh_frame.n, si.sh_frame.s);
    si.wi           = si.to_local(-ray.d);
    return si;
}

// scene.cpp:216-253: hit <=> tfar != maxt
SceneInteraction ray_intersect(const OScene &sc, const Ray &ray, RayCounters &rc, bool brute = false) {
    rc.closest++;
    RawHit h = brute ? intersect_brute(sc, ray) : intersect_bvh(sc, ray, false);
    if (h.t == ray.maxt) h = RawHit();
    return compute_scene_interaction(sc, ray, h);
}
// scene.cpp:255-273
bool ray_test(const OScene &sc, const Ray &ray, RayCounters &rc) {
    rc.shadow++;
    RawHit h = intersect_bvh(sc, ray, true);
    return h.t != Infinity && h.t != ray.maxt;
}

struct PositionSample { V3 p, n; V2 uv; float pdf; };

// mesh.cpp:103-133 + distribution.h:114-123 + warp.h:11-15
PositionSample mesh_sample_position(const OMesh &m, V2 sample) {
    uint32_t index;
    std::tie(index, sample.y) = distribution_sample_reuse(m.cdf, sample.y);
    const uint32_t *fi = &m.tris[(uint32_t) index * 3];
    V3 p0 = m.pos(fi[0]), p1 = m.pos(fi[1]), p2 = m.pos(fi[2]);
    V3 e0 = p1 - p0, e1 = p2 - p0;
    V2 b = square_to_uniform_triangle(sample);
    PositionSample ps;
    ps.p  = p0 + e0 * b.x + e1 * b.y;
    ps.uv = b;
    if (m.has_uvs) {
        V2 uv0 = m.uv(fi[0]), uv1 = m.uv(fi[1]), uv2 = m.uv(fi[2]);
        float w = 1.f - b.x - b.y;
        ps.uv   = { uv0.x * w + uv1.x * b.x + uv2.x * b.y, uv0.y * w + uv1.y * b.x + uv2.y * b.y };
    }
    V3 ns = normalized(cross(e0, e1));
    if (m.has_normals) {
        V3 n0 = m.nrm(fi[0]), n1 = m.nrm(fi[1]), n2 = m.nrm(fi[2]);
        ns = normalized(n0 * (1.f - b.x - b.y) + n1 * b.x + n2 * b.y);
    }
    ps.n   = ns;
    ps.pdf = 1.f / m.surface_area;
    return ps;
}

// shape.cpp:64-78
DirectIllumSample shape_sample_direct(const OMesh &m, const SceneInteraction &si, V2 sample) {
    PositionSample ps = mesh_sample_position(m, sample);
    DirectIllumSample ds;
    ds.p = ps.p; ds.n = ps.n; ds.uv = ps.uv; ds.pdf = ps.pdf;
    ds.d = ds.p - si.p;
    float dist_squared = squared_norm(ds.d);
    ds.dist            = std::sqrt(dist_squared);
    ds.d               = ds.d / ds.dist;
    float dp           = std::abs(dot(ds.d, ds.n));
    ds.pdf *= (dp != 0.f) ? dist_squared / dp : 0.f;
    return ds;
}
// shape.cpp:80-86 + mesh.cpp:135-137
float shape_pdf_direct(const OMesh &m, const DirectIllumSample &ds) {
    float pdf = 1.f / m.surface_area, dp = std::abs(dot(ds.d, ds.n));
    pdf *= (dp != 0.f) ? (ds.dist * ds.dist) / dp : 0.f;
    return pdf;
}

// ------------------------------------------------------------------------------------------
// Emitters: src/librender/emitters/{area,constant}.cpp
// ------------------------------------------------------------------------------------------
std::pair<DirectIllumSample, Spec> emitter_sample_direct(const OScene &sc, int e, const SceneInteraction &ref, V2 sample) {
    const MskEmitter &em = sc.emitters[e];
    if (em.type == MSK_EMITTER_AREA) { // area.cpp:33-45
        DirectIllumSample ds = shape_sample_direct(sc.meshes[em.shape], ref, sample);
        ds.object = e;
        if (dot(ds.d, ds.n) < 0.f && ds.pdf != 0.f)
            return { ds, texture_eval(sc.spectra, em.radiance, ds.uv.x, ds.uv.y, ref.wavelengths) / ds.pdf }; // SceneInteraction si(ds)
        ds.pdf = 0;
        return { ds, Spec(0.f) };
    } else { // constant.cpp:55-73
        V3 d       = square_to_uniform_sphere(sample);
        float dist = 2.f * sc.env_radius;
        DirectIllumSample ds;
        ds.p = ref.p + d * dist; ds.n = -d; ds.uv = { 0.f, 0.f };
        ds.pdf = InvFourPi; ds.object = e; ds.d = d; ds.dist = dist;
        // the reference evaluates radiance with a default-constructed interaction whose
        // wavelengths are uninitialised (constant.cpp:70-72); restated with the query's wavelengths
        return { ds, texture_eval(sc.spectra, em.radiance, 0.f, 0.f, ref.wavelengths) / ds.pdf };
    }
}
float emitter_pdf_direct(const OScene &sc, int e, const DirectIllumSample &ds) {
    const MskEmitter &em = sc.emitters[e];
    if (em.type == MSK_EMITTER_AREA) return shape_pdf_direct(sc.meshes[em.shape], ds); // area.cpp:47-49
    return InvFourPi;                                                                  // constant.cpp:75-77
}
Spec emitter_eval(const OScene &sc, int e, const SceneInteraction &si) {
    const MskEmitter &em = sc.emitters[e];
    if (em.type == MSK_EMITTER_AREA) // area.cpp:51-54
        return Frame::cos_theta(si.wi) > 0.f ? texture_eval(sc.spectra, em.radiance, si.uv.x, si.uv.y, si.wavelengths) : Spec(0.f);
    return texture_eval(sc.spectra, em.radiance, 0.f, 0.f, si.wavelengths); // constant.cpp:79-81 (a miss carries no uv)
}

// scene.cpp:69-103
std::pair<DirectIllumSample, Spec> sample_emitter_direct(const OScene &sc, const SceneInteraction &ref, V2 sample, RayCounters &rc) {
    DirectIllumSample ds;
    Spec spec;
    size_t ne = sc.emitters.size();
    if (ne) {
        if (ne == 1) {
            std::tie(ds, spec) = emitter_sample_direct(sc, 0, ref, sample);
        } else {
            float light_sel_pdf = 1.f / ne;
            uint32_t index      = std::min(uint32_t(sample.x * (float) ne), (uint32_t) ne - 1);
            sample.x            = (sample.x - index * light_sel_pdf) * ne;
            std::tie(ds, spec)  = emitter_sample_direct(sc, (int) index, ref, sample);
            ds.pdf *= light_sel_pdf;
            spec *= (float) ne;
        }
        if (ds.pdf != 0.f) {
            Ray ray{ ref.p, ds.d, RayEpsilon * (1.f + max_abs_coeff(ref.p)), ds.dist * (1.f - ShadowEpsilon), ref.wavelengths };
            if (ray_test(sc, ray, rc)) spec = Spec(0.f);
        }
    } else {
        spec = Spec(0.f);
    }
    return { ds, spec };
}
// scene.cpp:105-112
float pdf_emitter_direct(const OScene &sc, const DirectIllumSample &ds) {
    if (sc.emitters.size() == 1) return emitter_pdf_direct(sc, 0, ds);
    return emitter_pdf_direct(sc, ds.object, ds) * (1.f / sc.emitters.size());
}

// ------------------------------------------------------------------------------------------
// BSDFs: src/librender/bsdfs/*.cpp  (BSDFContext is always the default: Radiance, all lobes)
// ------------------------------------------------------------------------------------------
enum : uint32_t { // bsdf.h:14-31
    F_Null = 0x1, F_DiffuseReflection = 0x2, F_DiffuseTransmission = 0x4, F_GlossyReflection = 0x8,
    F_GlossyTransmission = 0x10, F_DeltaReflection = 0x20, F_DeltaTransmission = 0x40,
    F_Delta = F_Null | F_DeltaReflection | F_DeltaTransmission,
    F_Smooth = F_DiffuseReflection | F_DiffuseTransmission | F_GlossyReflection | F_GlossyTransmission
};
struct BSDFSample { V3 wo; float pdf = 0.f, eta = 1.f; uint32_t sampled_type = 0; };

uint32_t bsdf_flags(const MskBsdf &b) {
    switch (b.type) {
        case MSK_BSDF_DIFFUSE: return F_DiffuseReflection;                           // diffuse.cpp:14
        case MSK_BSDF_CONDUCTOR: return F_DeltaReflection;                           // conductor.cpp:18
        case MSK_BSDF_ROUGHCONDUCTOR: return F_GlossyReflection;                     // roughconductor.cpp:47
        case MSK_BSDF_ROUGHDIELECTRIC: return F_GlossyReflection | F_GlossyTransmission; // roughdielectric.cpp:52-54
        case MSK_BSDF_DIELECTRIC: return F_DeltaReflection | F_DeltaTransmission;    // dielectric.cpp:21-23
    }
    return 0;
}

Spec tex(const OScene &sc, int id, const SceneInteraction &si) { return texture_eval(sc.spectra, id, si.uv.x, si.uv.y, si.wavelengths); }

std::pair<BSDFSample, Spec> bsdf_sample_1(const OScene &sc, const MskBsdf &b, const SceneInteraction &si, float sample1, V2 sample) {
    BSDFSample bs;
    float cos_theta_i = Frame::cos_theta(si.wi);
    switch (b.type) {
        case MSK_BSDF_DIFFUSE: { // diffuse.cpp:19-32
            if (cos_theta_i <= 0.f) return { bs, Spec(0.f) };
            bs.wo = square_to_cosine_hemisphere(sample);
            bs.pdf = square_to_cosine_hemisphere_pdf(bs.wo);
            bs.eta = 1.f; bs.sampled_type = F_DiffuseReflection;
            return { bs, bs.pdf > 0.f ? tex(sc, b.reflectance, si) : Spec(0.f) };
        }
        case MSK_BSDF_CONDUCTOR: { // conductor.cpp:22-39 (stale RGB API restated spectrally)
            if (cos_theta_i <= 0.f) return { bs, Spec(0.f) };
            bs.wo = reflect(si.wi); bs.pdf = 1.f; bs.eta = 1.f; bs.sampled_type = F_DeltaReflection;
            Spec value = tex(sc, b.reflectance, si) * fresnel_conductor(cos_theta_i, tex(sc, b.eta, si), tex(sc, b.k, si));
            return { bs, value };
        }
        case MSK_BSDF_ROUGHCONDUCTOR: { // roughconductor.cpp:53-80
            if (cos_theta_i <= 0.f) return { bs, Spec(0.f) };
            Microfacet distr(b.alpha_u, b.alpha_v);
            V3 m;
            std::tie(m, bs.pdf) = distr.sample(si.wi, sample);
            bs.wo = reflect(si.wi, m); bs.eta = 1.f; bs.sampled_type = F_GlossyReflection;
            if (!(bs.pdf != 0.f && Frame::cos_theta(bs.wo) > 0.f)) return { bs, Spec(0.f) };
            float weight;
            if (b.sample_visible) weight = distr.smith_g1(bs.wo, m);
            else weight = distr.G(si.wi, bs.wo, m) * dot(si.wi, m) / (cos_theta_i * Frame::cos_theta(m));
            bs.pdf /= 4.f * dot(bs.wo, m);
            Spec F = fresnel_conductor(dot(si.wi, m), tex(sc, b.eta, si), tex(sc, b.k, si));
            return { bs, F * weight }; // NB: the reference omits specular_reflectance here (:79) but not in eval (:99)
        }
        case MSK_BSDF_ROUGHDIELECTRIC: { // roughdielectric.cpp:58-114
            float m_eta = b.int_ior / b.ext_ior;
            Microfacet distr(b.alpha_u, b.alpha_v);
            Microfacet sample_distr(distr);
            if (!b.sample_visible) sample_distr.scale_alpha(1.2f - .2f * std::sqrt(std::abs(cos_theta_i)));
            V3 m;
            std::tie(m, bs.pdf) = sample_distr.sample(si.wi * std::copysign(1.f, cos_theta_i), sample);
            if (bs.pdf == 0) return { bs, Spec(0.f) };
            FresnelResult fr = fresnel(dot(si.wi, m), m_eta);
            bool selected_r  = sample1 <= fr.F;
            Spec weight(1.f);
            bs.pdf *= selected_r ? fr.F : (1.f - fr.F);
            bs.eta          = selected_r ? 1.f : fr.eta_it;
            bs.sampled_type = selected_r ? F_GlossyReflection : F_GlossyTransmission;
            float dwh_dwo = 0.f;
            if (selected_r) {
                bs.wo = reflect(si.wi, m);
                weight *= tex(sc, b.reflectance, si);
                dwh_dwo = 1.f / (4.f * dot(bs.wo, m));
            } else {
                bs.wo = refract(si.wi, m, fr.cos_theta_t, fr.eta_ti);
                weight *= sqr(fr.eta_ti); // TransportMode::Radiance
                dwh_dwo = sqr(bs.eta) * dot(bs.wo, m) / sqr(dot(si.wi, m) + bs.eta * dot(bs.wo, m));
            }
            if (b.sample_visible) weight *= distr.smith_g1(bs.wo, m);
            else weight *= distr.G(si.wi, bs.wo, m) * dot(si.wi, m) / (cos_theta_i * Frame::cos_theta(m));
            bs.pdf *= std::abs(dwh_dwo);
            return { bs, weight };
        }
        case MSK_BSDF_DIELECTRIC: { // dielectric.cpp:26-72
            float m_eta      = b.int_ior / b.ext_ior;
            FresnelResult fr = fresnel(cos_theta_i, m_eta);
            float r_i = fr.F, t_i = 1.f - r_i;
            bool selected_r = sample.x <= r_i;
            bs.pdf          = selected_r ? r_i : t_i;
            bs.sampled_type = selected_r ? F_DeltaReflection : F_DeltaTransmission;
            bs.wo           = selected_r ? reflect(si.wi) : refract(si.wi, fr.cos_theta_t, fr.eta_ti);
            bs.eta          = selected_r ? 1.f : fr.eta_it;
            Spec weight(1.f);
            if (selected_r) weight *= tex(sc, b.reflectance, si);
            else weight *= tex(sc, b.transmittance, si) * fr.eta_ti * fr.eta_ti;
            return { bs, weight };
        }
    }
    return { bs, Spec(0.f) };
}

Spec bsdf_eval_1(const OScene &sc, const MskBsdf &b, const SceneInteraction &si, V3 wo) {
    float cos_theta_i = Frame::cos_theta(si.wi), cos_theta_o = Frame::cos_theta(wo);
    switch (b.type) {
        case MSK_BSDF_DIFFUSE: // diffuse.cpp:34-46
            if (cos_theta_i > 0.f && cos_theta_o > 0.f) return tex(sc, b.reflectance, si) * InvPi * cos_theta_o;
            return Spec(0.f);
        case MSK_BSDF_ROUGHCONDUCTOR: { // roughconductor.cpp:82-100
            if (!(cos_theta_i > 0.f && cos_theta_o > 0.f)) return Spec(0.f);
            V3 H = normalized(wo + si.wi);
            Microfacet distr(b.alpha_u, b.alpha_v);
            float D = distr.eval(H);
            if (D == 0) return Spec(0.f);
            float G      = distr.G(si.wi, wo, H);
            float result = D * G / (4.f * Frame::cos_theta(si.wi));
            Spec F       = fresnel_conductor(dot(si.wi, H), tex(sc, b.eta, si), tex(sc, b.k, si));
            return F * tex(sc, b.reflectance, si) * result;
        }
        case MSK_BSDF_ROUGHDIELECTRIC: { // roughdielectric.cpp:116-153
            if (cos_theta_i == 0.f) return Spec(0.f);
            float m_eta = b.int_ior / b.ext_ior, m_inv_eta = b.ext_ior / b.int_ior;
            bool reflect_ = cos_theta_i * cos_theta_o > 0.f;
            float eta = cos_theta_i > 0.f ? m_eta : m_inv_eta, inv_eta = cos_theta_i > 0.f ? m_inv_eta : m_eta;
            V3 m = normalized(si.wi + wo * (reflect_ ? 1.f : eta));
            m    = m * std::copysign(1.f, Frame::cos_theta(m));
            Microfacet distr(b.alpha_u, b.alpha_v);
            float D = distr.eval(m);
            float F = fresnel(dot(si.wi, m), m_eta).F;
            float G = distr.G(si.wi, wo, m);
            if (reflect_)
                return F * D * G * tex(sc, b.reflectance, si) / (4.f * std::abs(cos_theta_i));
            float scale = sqr(inv_eta);
            return tex(sc, b.transmittance, si) *
                   std::abs((scale * (1.f - F) * D * G * eta * eta * dot(si.wi, m) * dot(wo, m)) /
                            (cos_theta_i * sqr(dot(si.wi, m) + eta * dot(wo, m))));
        }
        default: return Spec(0.f); // conductor.cpp:41-44, dielectric.cpp:74-77
    }
}

float bsdf_pdf_1(const MskBsdf &b, const SceneInteraction &si, V3 wo) {
    float cos_theta_i = Frame::cos_theta(si.wi), cos_theta_o = Frame::cos_theta(wo);
    switch (b.type) {
        case MSK_BSDF_DIFFUSE: // diffuse.cpp:48-57
            return (cos_theta_i > 0.f && cos_theta_o > 0.f) ? square_to_cosine_hemisphere_pdf(wo) : 0.f;
        case MSK_BSDF_ROUGHCONDUCTOR: { // roughconductor.cpp:102-120
            V3 m = normalized(wo + si.wi);
            if (!(cos_theta_i > 0.f && cos_theta_o > 0.f && dot(si.wi, m) > 0.f && dot(wo, m) > 0.f)) return 0.f;
            Microfacet distr(b.alpha_u, b.alpha_v);
            if (b.sample_visible) return distr.eval(m) * distr.smith_g1(si.wi, m) / (4.f * cos_theta_i);
            return distr.pdf(si.wi, m) / (4.f * dot(wo, m));
        }
        case MSK_BSDF_ROUGHDIELECTRIC: { // roughdielectric.cpp:155-190
            if (cos_theta_i == 0.f) return 0.f;
            float m_eta = b.int_ior / b.ext_ior, m_inv_eta = b.ext_ior / b.int_ior;
            bool reflect_ = cos_theta_i * cos_theta_o > 0.f;
            float eta     = cos_theta_i > 0.f ? m_eta : m_inv_eta;
            V3 m = normalized(si.wi + wo * (reflect_ ? 1.f : eta));
            m    = m * std::copysign(1.f, Frame::cos_theta(m));
            if (dot(si.wi, m) * Frame::cos_theta(si.wi) <= 0.f || dot(wo, m) * Frame::cos_theta(wo) <= 0.f) return 0.f;
            float dwh_dwo = reflect_ ? 1.f / (4.f * dot(wo, m)) : (eta * eta * dot(wo, m)) / sqr(dot(si.wi, m) + eta * dot(wo, m));
            Microfacet sample_distr(b.alpha_u, b.alpha_v);
            if (!b.sample_visible) sample_distr.scale_alpha(1.2f - .2f * std::sqrt(std::abs(Frame::cos_theta(si.wi))));
            float prob = sample_distr.pdf(si.wi * std::copysign(1.f, Frame::cos_theta(si.wi)), m);
            float F    = fresnel(dot(si.wi, m), m_eta).F;
            prob *= (reflect_ ? F : 1.f - F);
            return prob * std::abs(dwh_dwo);
        }
        default: return 0.f;
    }
}

// twosided.cpp:38-101 (same BRDF on both sides)
std::pair<BSDFSample, Spec> bsdf_sample(const OScene &sc, const MskBsdf &b, const SceneInteraction &si_, float s1, V2 s2) {
    if (!b.twosided) return bsdf_sample_1(sc, b, si_, s1, s2);
    SceneInteraction si(si_);
    BSDFSample bs;
    std::pair<BSDFSample, Spec> ret{ bs, Spec(0.f) };
    if (Frame::cos_theta(si.wi) > 0.f) ret = bsdf_sample_1(sc, b, si, s1, s2);
    if (Frame::cos_theta(si.wi) < 0.f) {
        si.wi.z *= -1.f;
        ret = bsdf_sample_1(sc, b, si, s1, s2);
        ret.first.wo.z *= -1.f;
    }
    return ret;
}
Spec bsdf_eval(const OScene &sc, const MskBsdf &b, const SceneInteraction &si_, V3 wo) {
    if (!b.twosided) return bsdf_eval_1(sc, b, si_, wo);
    SceneInteraction si(si_);
    Spec result(0.f);
    if (Frame::cos_theta(si.wi) > 0.f) result = bsdf_eval_1(sc, b, si, wo);
    if (Frame::cos_theta(si.wi) < 0.f) { si.wi.z *= -1.f; wo.z *= -1.f; result = bsdf_eval_1(sc, b, si, wo); }
    return result;
}
float bsdf_pdf(const MskBsdf &b, const SceneInteraction &si_, V3 wo) {
    if (!b.twosided) return bsdf_pdf_1(b, si_, wo);
    SceneInteraction si(si_);
    float result = 0.f;
    if (Frame::cos_theta(si.wi) > 0.f) result = bsdf_pdf_1(b, si, wo);
    if (Frame::cos_theta(si.wi) < 0.f) { si.wi.z *= -1.f; wo.z *= -1.f; result = bsdf_pdf_1(b, si, wo); }
    return result;
}

// ------------------------------------------------------------------------------------------
// Camera: src/librender/sensors/perspective.cpp:22-41, transform.h:129-136
// ------------------------------------------------------------------------------------------
V3 apply_point(const float M[16], V3 p) {
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = M[i * 4 + 0] * p.x + M[i * 4 + 1] * p.y + M[i * 4 + 2] * p.z + M[i * 4 + 3] * 1.f;
    return V3(r[0], r[1], r[2]) / r[3];
}
V3 apply_vector(const float M[16], V3 v) {
    return { M[0] * v.x + M[1] * v.y + M[2] * v.z, M[4] * v.x + M[5] * v.y + M[6] * v.z, M[8] * v.x + M[9] * v.y + M[10] * v.z };
}
std::pair<Ray, Spec> camera_sample_ray(const MskCamera &cam, float wavelength_sample, V2 pos_sample) {
    Ray ray;
    Spec wav_weight;
    sample_wavelength(wavelength_sample, ray.wavelengths, wav_weight);
    V3 near_p   = apply_point(cam.sample_to_camera, { pos_sample.x, pos_sample.y, 0.f });
    V3 d        = normalized(near_p);
    float inv_z = 1.f / d.z;
    ray.mint    = cam.near_clip * inv_z;
    ray.maxt    = cam.far_clip * inv_z;
    ray.o       = apply_point(cam.to_world, { 0.f, 0.f, 0.f });
    ray.d       = apply_vector(cam.to_world, d);
    return { ray, wav_weight };
}

// ------------------------------------------------------------------------------------------
// PathTracer::sample: src/librender/integrators/path.cpp:23-131
// ------------------------------------------------------------------------------------------
inline float mis_weight(float pdf_a, float pdf_b) { // :127-131
    pdf_a *= pdf_a;
    pdf_b *= pdf_b;
    return pdf_a > 0.f ? pdf_a / (pdf_a + pdf_b) : 0.f;
}

struct PathParams {
    int max_depth, rr_depth; bool hide_emitter; int integrator = MSK_INTEGRATOR_PATH;
    // path.cpp:71-72 draws `bsdf->sample(ctx, si, sampler->next1d(), sampler->next2d())`: the order of the two draws is
    // unspecified in C++.  The determinism contract fixes it left to right (next1d first); GCC evaluates arguments right to
    // left, so the golden vectors of the reference compiled here need the other order to be replayed (test-only switch).
    bool draw_bsdf_samples_right_to_left = false;
};

Spec path_sample(const OScene &sc, Sampler &sampler, const Ray &ray_, const PathParams &pp, RayCounters &rc) {
    Ray ray = ray_;
    Spec throughput(1.f), result(0.f);
    float eta      = 1.f;
    bool scattered = false;
    SceneInteraction si = ray_intersect(sc, ray, rc);
    for (int depth = 1; depth <= pp.max_depth || pp.max_depth < 0; depth++) {
        if (!si.is_valid()) {
            if (depth == 1 && (!pp.hide_emitter || scattered))
                if (sc.environment >= 0) result += throughput * emitter_eval(sc, sc.environment, si);
            break;
        }
        int emitter = sc.meshes[si.shape].emitter;
        if (emitter >= 0 && depth == 1 && (!pp.hide_emitter || scattered))
            result += throughput * emitter_eval(sc, emitter, si);
        if (depth >= pp.max_depth && pp.max_depth > 0) break;

        DirectIllumSample ds;
        const MskBsdf &bsdf = sc.bsdfs[sc.meshes[si.shape].bsdf];
        if (bsdf_flags(bsdf) & F_Smooth) {
            Spec emitter_val;
            std::tie(ds, emitter_val) = sample_emitter_direct(sc, si, sampler.next2d(), rc);
            if (ds.pdf != 0.f) {
                V3 wo         = si.to_local(ds.d);
                Spec bsdf_val = bsdf_eval(sc, bsdf, si, wo);
                float bpdf    = bsdf_pdf(bsdf, si, wo);
                float weight  = mis_weight(ds.pdf, bpdf);
                result += throughput * emitter_val * bsdf_val * weight;
            }
        }
        // draw order fixed left-to-right: next1d() then next2d() (unspecified in the reference, :72)
        float s1;
        V2 s2;
        if (!pp.draw_bsdf_samples_right_to_left) { s1 = sampler.next1d(); s2 = sampler.next2d(); }
        else { s2 = sampler.next2d(); s1 = sampler.next1d(); }
        auto [bs, bsdf_val] = bsdf_sample(sc, bsdf, si, s1, s2);
        scattered |= bs.sampled_type != (uint32_t) F_Null;
        // Output-equivalent shortcut for a failed sample (waives q5): with bsdf_val == 0 the
        // throughput is 0 from here on and nothing further can change `result`.
        if (bsdf_val.is_zero()) break;

        V3 wo            = si.to_world(bs.wo);
        bool hit_emitter = false;
        Spec value(0.f);
        ray                      = si.spawn_ray(wo);
        SceneInteraction si_bsdf = ray_intersect(sc, ray, rc);
        if (si_bsdf.is_valid()) {
            emitter = sc.meshes[si_bsdf.shape].emitter;
            if (emitter >= 0) {
                value = emitter_eval(sc, emitter, si_bsdf);
                // DirectIllumSample::set_query, records.cpp:7-14
                ds.p = si_bsdf.p; ds.n = si_bsdf.sh_frame.n; ds.uv = si_bsdf.uv; ds.object = emitter;
                ds.d = ray.d; ds.dist = si_bsdf.t;
                hit_emitter = true;
            }
        } else {
            if (sc.environment >= 0) {
                if (pp.hide_emitter && !scattered) break;
                value       = emitter_eval(sc, sc.environment, si); // passes the OLD si (q4): wavelengths only
                hit_emitter = true;
                // NB (q8): ds is NOT updated here, so pdf_emitter_direct below sees the NEE record
            } else
                break;
        }
        throughput *= bsdf_val;
        eta *= bs.eta;
        if (hit_emitter) {
            float emitter_pdf = 0.f;
            if (!(bs.sampled_type & F_Delta)) {
                // ds.object == -1 can only happen for a non-smooth BSDF, which is Delta
                emitter_pdf = pdf_emitter_direct(sc, ds);
            }
            result += throughput * value * mis_weight(bs.pdf, emitter_pdf);
        }
        si = si_bsdf;
        if (depth + 1 >= pp.rr_depth) {
            float q = std::min(throughput.max_coeff() * eta * eta, 0.95f);
            if (sampler.next1d() >= q) break;
            throughput /= q;
        }
    }
    return result;
}

// ------------------------------------------------------------------------------------------
// Volumetric path tracer: integrators/volpath.cpp:26-167 with media/homogeneous.cpp, phase/isotropic.cpp and
// Scene::sample_attenuated_emitter_direct / eval_transmittance (scene.cpp:114-184).  SURVEY 8f rank 4.
//
// The plugin is written against a stale API (RGB Spectrum, ray::spawn, MediumSample-based phase functions) and is
// commented out of the reference build (src/librender/CMakeLists.txt:106-112), so it is restated on the current
// spectral types with these decisions, shared with the GPU kernel:
//  * sigma_a / sigma_s are spectra evaluated at the path's four wavelengths; `channel` (volpath.cpp:39,
//    min(next1d * 3, 3 - 1) over RGB) becomes min(next1d * 4, 3) over the four wavelengths (Spectrum::Size)
//  * `.mean()` is Eigen's packet reduction ((v0 + v2) + (v1 + v3)) / 4, like spectrum_to_xyz
//  * ray::spawn(ray, 0, si.t) = the same ray over [0, si.t]; ray::spawn<false>(p, d) = Ray(p, d, RayEpsilon, inf)
//  * medium NEE (volpath.cpp:50-51) passes the SURFACE interaction `si` as the reference point -- whose p is
//    uninitialised when the ray escaped; the evident intent (a medium interaction at ms.p) is used instead
//  * no compiled BSDF carries the Null flag (bsdfs/mask.cpp is not built), so eval_transmittance's loop
//    (scene.cpp:152-182) reduces to: any surface within [mint, dist (1 - ShadowEpsilon)] -> 0, else
//    exp(-sigma_t dist) when the reference point lies in a medium
//  * eval_transmittance starts its visibility ray at the UNSCALED RayEpsilon (scene.cpp:146-149, marked "TODO: Need to
//    fix in volpath" there): in a scene of Cornell-box scale that is below the float spacing of the hit point, and
//    whether a shadow ray re-hits its own surface is decided by rounding.  The scaled offset of
//    Scene::sample_emitter_direct (scene.cpp:91-93), RayEpsilon (1 + max|p|), is used instead
//  * exp(sigma_t * -inf) with sigma_t == 0 is NaN in the reference (homogeneous.cpp:56-59 on an escaped ray);
//    a channel without extinction transmits 1 here
// Kept as written: `scale` is read and never applied (homogeneous.cpp:18); NEE is added WITHOUT the MIS weight
// that is computed next to it (volpath.cpp:105-109) and emitter hits are added whenever `emitted_radiance` is set
// (initially, and after a delta bounce; a medium scattering event leaves the flag untouched, :44-74); Russian
// roulette uses depth + 1 >= rr_depth (:158).
struct MediumSample { float t = Infinity; V3 p; Spec sigma_s, transmittance; float pdf = 0.f; };

inline float spec_mean(const Spec &v) { return ((v[0] + v[2]) + (v[1] + v[3])) / 4.f; }
// exp(sigma_t * -d), homogeneous.cpp:48,56-59 (with the zero-extinction guard above)
inline Spec medium_tr(const Spec &sigma_t, float d) {
    Spec r;
    for (int i = 0; i < 4; ++i) r[i] = sigma_t[i] == 0.f ? 1.f : std::exp(sigma_t[i] * (-d));
    return r;
}
inline Spec medium_sigma_t(const OScene &sc, int medium, const Spec &wl, Spec *sigma_s = nullptr) {
    const MskMedium &m = sc.media[medium];
    Spec sa = spectrum_eval(sc.spectra[m.sigma_a], wl), ss = spectrum_eval(sc.spectra[m.sigma_s], wl);
    if (sigma_s) *sigma_s = ss;
    return ss + sa; // homogeneous.cpp:17
}

// homogeneous.cpp:21-53
std::pair<bool, MediumSample> medium_sample_distance(const OScene &sc, int medium, const Ray &ray, float sample, uint32_t channel) {
    MediumSample ms;
    Spec sigma_s, sigma_t = medium_sigma_t(sc, medium, ray.wavelengths, &sigma_s);
    float sampled_distance = -std::log(1 - sample) / sigma_t[channel];
    bool success = true;
    if (sampled_distance < ray.maxt - ray.mint) {
        ms.t       = sampled_distance + ray.mint;
        ms.p       = ray.o + ray.d * ms.t;
        if (ms.p.x == ray.o.x && ms.p.y == ray.o.y && ms.p.z == ray.o.z) {
            ms.t    = Infinity;
            ms.pdf  = spec_mean(medium_tr(sigma_t, sampled_distance));
            success = false;
        } else
            ms.pdf = spec_mean(medium_tr(sigma_t, sampled_distance) * sigma_t);
    } else {
        ms.t             = Infinity;
        sampled_distance = ray.maxt - ray.mint;
        ms.pdf           = spec_mean(medium_tr(sigma_t, sampled_distance));
        success          = false;
    }
    ms.sigma_s       = sigma_s;
    ms.transmittance = medium_tr(sigma_t, sampled_distance);
    if (ms.transmittance.max_coeff() < 1e-20f) ms.transmittance = Spec(0.f);
    return { success, ms };
}

// scene.cpp:114-184 with the reductions stated above; `ref_p` is the surface point or ms.p
std::pair<DirectIllumSample, Spec> sample_attenuated_emitter_direct(const OScene &sc, const SceneInteraction &ref, int medium, V2 sample,
                                                                    RayCounters &rc) {
    DirectIllumSample ds;
    Spec spec(0.f);
    size_t ne = sc.emitters.size();
    if (ne == 0) return { ds, spec };
    if (ne == 1) {
        std::tie(ds, spec) = emitter_sample_direct(sc, 0, ref, sample);
    } else {
        float light_sel_pdf = 1.f / ne;
        uint32_t index = std::min(uint32_t(sample.x * (float) ne), (uint32_t) ne - 1);
        sample.x = (sample.x - index * light_sel_pdf) * ne;
        std::tie(ds, spec) = emitter_sample_direct(sc, (int) index, ref, sample);
        ds.pdf *= light_sel_pdf;
        spec *= (float) ne;
    }
    if (ds.pdf != 0.f) { // eval_transmittance(ref.p, ds.p, medium), scene.cpp:141-184
        V3 d = ds.p - ref.p;
        float remaining = norm(d);
        d = d / remaining;
        Ray ray{ ref.p, d, RayEpsilon * (1.f + max_abs_coeff(ref.p)), remaining * (1 - ShadowEpsilon), ref.wavelengths };
        if (ray_test(sc, ray, rc)) spec = Spec(0.f); // a surface without the Null flag blocks the segment (:156-158)
        else if (medium >= 0) spec = spec * medium_tr(medium_sigma_t(sc, medium, ref.wavelengths), remaining); // :159-164
    }
    return { ds, spec };
}

Spec volpath_sample(const OScene &sc, Sampler &sampler, const Ray &ray_, const PathParams &pp, RayCounters &rc) {
    Ray ray = ray_;
    Spec throughput(1.f), result(0.f);
    float eta  = 1.f;
    int medium = sc.sensor_medium; // integrator.cpp:116: sample(scene, sampler, ray, sensor->medium(), ...)
    bool ms_flag = false, scattered = false, emitted_radiance = true;
    SceneInteraction si = ray_intersect(sc, ray, rc);
    MediumSample ms;
    uint32_t channel = std::min<uint32_t>((uint32_t) (sampler.next1d() * 4), 4 - 1);
    for (int depth = 1; depth <= pp.max_depth || pp.max_depth < 0; depth++) {
        if (medium >= 0) {
            Ray seg{ ray.o, ray.d, 0.f, si.t, ray.wavelengths };
            std::tie(ms_flag, ms) = medium_sample_distance(sc, medium, seg, sampler.next1d(), channel);
        }
        if (medium >= 0 && ms_flag) {
            throughput *= ms.sigma_s * ms.transmittance / ms.pdf;
            SceneInteraction mi; // medium interaction at ms.p (see the header comment)
            mi.p = ms.p; mi.wavelengths = ray.wavelengths;
            auto [ds, spec] = sample_attenuated_emitter_direct(sc, mi, medium, sampler.next2d(), rc);
            if (!spec.is_zero()) result += throughput * spec * InvFourPi; // isotropic.cpp:24-27
            if (depth + 1 >= pp.max_depth && pp.max_depth > 0) break;
            V2 s2       = sampler.next2d();
            V3 phase_wo = square_to_uniform_sphere(s2); // isotropic.cpp:16-22, phase_val = 1
            ray         = Ray{ ms.p, phase_wo, RayEpsilon, Infinity, ray.wavelengths };
            si          = ray_intersect(sc, ray, rc);
            scattered   = true;
        } else {
            if (medium >= 0) throughput *= ms.transmittance / ms.pdf;
            if (!si.is_valid()) {
                if (emitted_radiance && (!pp.hide_emitter || scattered)) {
                    Spec value = throughput * (sc.environment >= 0 ? emitter_eval(sc, sc.environment, si) : Spec(0.f));
                    if (medium >= 0) value = value * medium_tr(medium_sigma_t(sc, medium, ray.wavelengths), ray.maxt - ray.mint);
                    result += value;
                }
                break;
            }
            const OMesh &mesh = sc.meshes[si.shape];
            if (mesh.emitter >= 0 && emitted_radiance && (!pp.hide_emitter || scattered))
                result += throughput * emitter_eval(sc, mesh.emitter, si);
            const MskBsdf &bsdf = sc.bsdfs[mesh.bsdf];
            if (bsdf_flags(bsdf) & F_Smooth) {
                auto [ds, emitter_val] = sample_attenuated_emitter_direct(sc, si, medium, sampler.next2d(), rc);
                if (ds.pdf != 0.f) {
                    V3 wo         = si.to_local(ds.d);
                    Spec bsdf_val = bsdf_eval(sc, bsdf, si, wo);
                    result += throughput * emitter_val * bsdf_val; // the MIS weight of :108 is computed and unused
                }
            }
            float s1 = sampler.next1d();
            V2 s2    = sampler.next2d();
            auto [bs, bsdf_val] = bsdf_sample(sc, bsdf, si, s1, s2);
            if (bsdf_val.is_zero()) break;
            emitted_radiance = false;
            bool recursive   = depth + 1 < pp.max_depth || pp.max_depth < 0;
            // :129-138 with Null never set: a delta bounce re-enables directly seen emitters
            if ((depth < pp.max_depth || pp.max_depth < 0) && (bs.sampled_type & F_Delta)) {
                emitted_radiance = true;
                recursive        = true;
            }
            if (!recursive) break;
            V3 wo = si.to_world(bs.wo);
            throughput *= bsdf_val;
            eta *= bs.eta;
            if (mesh.interior_medium >= 0 || mesh.exterior_medium >= 0) // interaction.cpp:10-13
                medium = dot(wo, si.n) > 0 ? mesh.exterior_medium : mesh.interior_medium;
            ray       = si.spawn_ray(wo);
            si        = ray_intersect(sc, ray, rc);
            scattered = true; // |= !Null
        }
        if (depth + 1 >= pp.rr_depth) {
            float q = std::min(throughput.max_coeff() * eta * eta, 0.95f);
            if (sampler.next1d() >= q) break;
            throughput /= q;
        }
    }
    return result;
}

Spec integrator_sample(const OScene &sc, Sampler &sampler, const Ray &ray, const PathParams &pp, RayCounters &rc) {
    return pp.integrator == MSK_INTEGRATOR_VOLPATH ? volpath_sample(sc, sampler, ray, pp, rc) : path_sample(sc, sampler, ray, pp, rc);
}

// ------------------------------------------------------------------------------------------
// Film: rfilter.{h,cpp}, filters/gaussian.cpp, imageblock.cpp, films/hdrfilm.cpp
// ------------------------------------------------------------------------------------------
constexpr int FILTER_RES = 32; // rfilter.h:6

inline float eval_discretized(const MskCamera &cam, float x) { // rfilter.h:13-16
    float scale_factor = float(FILTER_RES) / cam.filter_radius;
    return cam.filter_table[std::min((int) std::abs(x * scale_factor), FILTER_RES)];
}

struct Block {
    int ox, oy, sx, sy, border;
    int nch = 5;
    std::vector<float> data; // (sy+2b) x (sx+2b) x nch
};

// imageblock.cpp:55-114
void block_put(Block &b, const MskCamera &cam, V2 pos_, const float *value) {
    float r = cam.filter_radius;
    int w = b.sx + 2 * b.border, h = b.sy + 2 * b.border;
    float px = pos_.x - 0.5f - (b.ox - b.border), py = pos_.y - 0.5f - (b.oy - b.border);
    int lox = std::max((int) std::ceil(px - r), 0), loy = std::max((int) std::ceil(py - r), 0);
    int hix = std::min((int) std::floor(px + r), w - 1), hiy = std::min((int) std::floor(py + r), h - 1);
    float wx[8], wy[8];
    for (int x = lox, i = 0; x <= hix; ++x) wx[i++] = eval_discretized(cam, x - px);
    for (int y = loy, i = 0; y <= hiy; ++y) wy[i++] = eval_discretized(cam, y - py);
    for (int y = loy, yr = 0; y <= hiy; ++y, ++yr) {
        float *dest = b.data.data() + (y * (size_t) w + lox) * b.nch;
        for (int x = lox, xr = 0; x <= hix; ++x, ++xr) {
            float weight = wx[xr] * wy[yr];
            for (int k = 0; k < b.nch; ++k) *dest++ += weight * value[k];
        }
    }
}

// imageblock.cpp:36-53,133-173 specialised to "padded block into border-less film"
void film_put(float *film, int W, int H, const Block &b) {
    int w = b.sx + 2 * b.border, h = b.sy + 2 * b.border;
    for (int y = 0; y < h; ++y) {
        int fy = b.oy - b.border + y;
        if (fy < 0 || fy >= H) continue;
        for (int x = 0; x < w; ++x) {
            int fx = b.ox - b.border + x;
            if (fx < 0 || fx >= W) continue;
            for (int k = 0; k < b.nch; ++k) film[((size_t) fy * W + fx) * b.nch + k] += b.data[((size_t) y * w + x) * b.nch + k];
        }
    }
}

// BlockGenerator: imageblock.cpp:176-246 (spiral order)
struct BlockDesc { int ox, oy, sx, sy; };
std::vector<BlockDesc> spiral_blocks(int W, int H, int bs) {
    int bx = (int) std::ceil(W / (float) bs), by = (int) std::ceil(H / (float) bs);
    int count = bx * by;
    std::vector<BlockDesc> out;
    int px = bx / 2, py = by / 2, dir = 0 /*Right,Down,Left,Up*/, steps_left = 1, steps = 1;
    for (int c = 0; c < count; ++c) {
        int ox = px * bs, oy = py * bs;
        out.push_back({ ox, oy, std::min(W - ox, bs), std::min(H - oy, bs) });
        if (c + 1 != count) {
            do {
                switch (dir) { case 0: ++px; break; case 1: ++py; break; case 2: --px; break; case 3: --py; break; }
                if (--steps_left == 0) {
                    dir = (dir + 1) % 4;
                    if (dir == 2 || dir == 0) ++steps;
                    steps_left = steps;
                }
            } while (px < 0 || py < 0 || px >= bx || py >= by);
        }
    }
    return out;
}

} // namespace

// ==========================================================================================
// C API (see oracle.h)
// ==========================================================================================
struct OrcScene { OScene sc; };

// Batch queries are independent per ray: chunks of 256 rays handed out to all host threads.
template <typename F> static void parallel_rays(size_t n, F &&body) {
    const unsigned nthreads = (unsigned) std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), (n + 255) / 256);
    if (nthreads <= 1) { for (size_t i = 0; i < n; ++i) body(i); return; }
    std::atomic<size_t> next{ 0 };
    auto worker = [&]() {
        for (;;) {
            const size_t b = next.fetch_add(256);
            if (b >= n) break;
            for (size_t i = b, e = std::min(n, b + 256); i < e; ++i) body(i);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

extern "C" {

const char *orc_last_error(void) { return g_error.c_str(); }

int orc_scene_create(const MskSceneDesc *d, OrcScene **out) {
    if (!d || !out) return fail("null argument");
    auto S = std::make_unique<OrcScene>();
    OScene &sc = S->sc;
    sc.cam = d->camera;
    sc.environment = d->environment;
    sc.bsdfs.assign(d->bsdfs, d->bsdfs + d->nbsdfs);
    sc.emitters.assign(d->emitters, d->emitters + d->nemitters);
    for (uint32_t i = 0; i < d->nspectra; ++i) {
        const MskSpectrum &s = d->spectra[i];
        OSpectrum o{};
        o.kind = s.kind; o.c[0] = s.c[0]; o.c[1] = s.c[1]; o.c[2] = s.c[2]; o.value = s.value;
        o.lambda_min = s.lambda_min; o.lambda_max = s.lambda_max; o.inv_interval_size = 0.f;
        if (s.kind == MSK_SPEC_REGULAR || s.kind == MSK_SPEC_SRGB_D65) {
            if (s.table_size < 2 || s.table_offset + s.table_size > d->ntable_floats) return fail("bad spectrum table");
            o.table.assign(d->spectrum_tables + s.table_offset, d->spectrum_tables + s.table_offset + s.table_size);
            // regular.cpp:38-39,58: interval size in double, reciprocal stored as float
            double range = double(s.lambda_max) - double(s.lambda_min), interval = range / (s.table_size - 1);
            o.inv_interval_size = float(1. / interval);
        }
        if (s.kind == MSK_SPEC_CHECKERBOARD) {
            if (s.child0 < 0 || s.child1 < 0 || (uint32_t) s.child0 >= i || (uint32_t) s.child1 >= i) return fail("bad checkerboard children");
            o.child0 = s.child0; o.child1 = s.child1;
            for (int k = 0; k < 6; ++k) o.to_uv[k] = s.to_uv[k];
        }
        sc.spectra.push_back(std::move(o));
    }
    for (uint32_t i = 0; i < d->nmedia; ++i) {
        const MskMedium &m = d->media[i];
        if (m.sigma_a < 0 || (uint32_t) m.sigma_a >= d->nspectra || m.sigma_s < 0 || (uint32_t) m.sigma_s >= d->nspectra) return fail("bad medium spectrum");
        if (m.phase != MSK_PHASE_ISOTROPIC) return fail("unknown phase function");
        sc.media.push_back(m);
    }
    sc.sensor_medium = d->nmedia && d->sensor_medium >= 0 && d->sensor_medium < (int) d->nmedia ? d->sensor_medium : -1;
    sc.bbox_min = V3(Infinity, Infinity, Infinity);
    sc.bbox_max = V3(-Infinity, -Infinity, -Infinity);
    for (uint32_t i = 0; i < d->nmeshes; ++i) {
        const MskMesh &m = d->meshes[i];
        OMesh o;
        o.nverts = m.nverts; o.ntris = m.ntris; o.bsdf = m.bsdf; o.emitter = m.emitter;
        o.has_normals = m.has_normals; o.has_uvs = m.has_uvs;
        if (d->nmedia) { // medium ids are only meaningful when the description carries media
            if (m.interior_medium < -1 || m.interior_medium >= (int) d->nmedia || m.exterior_medium < -1 || m.exterior_medium >= (int) d->nmedia)
                return fail("bad medium id");
            o.interior_medium = m.interior_medium; o.exterior_medium = m.exterior_medium;
        }
        o.verts.assign(m.verts, m.verts + (size_t) m.nverts * 8);
        o.tris.assign(m.tris, m.tris + (size_t) m.ntris * 3);
        for (uint32_t t = 0; t < m.ntris * 3; ++t)
            if (o.tris[t] >= m.nverts) return fail("triangle index out of range");
        if (m.bsdf < 0 || (uint32_t) m.bsdf >= d->nbsdfs) return fail("bad bsdf id");
        for (uint32_t v = 0; v < m.nverts; ++v) { // Mesh::recompute_bbox, mesh.cpp:22-26
            V3 p = o.pos(v);
            sc.bbox_min = V3(std::min(sc.bbox_min.x, p.x), std::min(sc.bbox_min.y, p.y), std::min(sc.bbox_min.z, p.z));
            sc.bbox_max = V3(std::max(sc.bbox_max.x, p.x), std::max(sc.bbox_max.y, p.y), std::max(sc.bbox_max.z, p.z));
        }
        area_distr_build(o);
        sc.meshes.push_back(std::move(o));
    }
    // constant.cpp:21-28 + bbox.h:109-112
    if (d->nmeshes) {
        V3 c = (sc.bbox_max + sc.bbox_min) * 0.5f;
        float radius  = norm(c - sc.bbox_max);
        sc.env_radius = std::max(RayEpsilon, radius * (1.f + RayEpsilon));
    }
    build_bvh(sc);
    *out = S.release();
    return 0;
}

void orc_scene_destroy(OrcScene *s) { delete s; }

int orc_intersect(OrcScene *s, const MskRay *rays, MskHit *hits, size_t n, int brute_force) {
    if (!s) return fail("null scene");
    parallel_rays(n, [&](size_t i) {
        Ray r{ V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmin, rays[i].tmax, Spec() };
        RawHit h = brute_force ? intersect_brute(s->sc, r) : intersect_bvh(s->sc, r, false);
        if (h.t == r.maxt) h = RawHit();
        hits[i] = { h.t, h.u, h.v, h.prim, h.geom };
    });
    return 0;
}

int orc_occluded(OrcScene *s, const MskRay *rays, uint8_t *occ, size_t n) {
    if (!s) return fail("null scene");
    parallel_rays(n, [&](size_t i) {
        RayCounters rc;
        Ray r{ V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmin, rays[i].tmax, Spec() };
        occ[i] = ray_test(s->sc, r, rc) ? 1 : 0;
    });
    return 0;
}

// "second closest" distance, used by the tests to define non-degenerate rays (hard part iv)
int orc_intersect_margin(OrcScene *s, const MskRay *rays, float *second_t, float *min_bary, size_t n) {
    if (!s) return fail("null scene");
    parallel_rays(n, [&](size_t i) {
        Ray r{ V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmin, rays[i].tmax, Spec() };
        float t1 = Infinity, t2 = Infinity, mb = 0.f;
        for (uint32_t g = 0; g < s->sc.meshes.size(); ++g) {
            const OMesh &m = s->sc.meshes[g];
            for (uint32_t p = 0; p < m.ntris; ++p) {
                const uint32_t *f = &m.tris[p * 3];
                float t, u, v;
                if (tri_intersect(m.pos(f[0]), m.pos(f[1]), m.pos(f[2]), r.o, r.d, r.mint, r.maxt, t, u, v)) {
                    if (t < t1) { t2 = t1; t1 = t; mb = std::min(u, std::min(v, 1.f - u - v)); }
                    else if (t < t2) t2 = t;
                }
            }
        }
        second_t[i] = t2; min_bary[i] = mb;
    });
    return 0;
}

int orc_camera_rays(OrcScene *s, const float *samples /* n x 3: px, py, wavelength sample */, MskRay *rays, size_t n) {
    if (!s) return fail("null scene");
    for (size_t i = 0; i < n; ++i) {
        auto [ray, w] = camera_sample_ray(s->sc.cam, samples[i * 3 + 2], { samples[i * 3], samples[i * 3 + 1] });
        rays[i] = { { ray.o.x, ray.o.y, ray.o.z }, ray.mint, { ray.d.x, ray.d.y, ray.d.z }, ray.maxt };
    }
    return 0;
}

// The camera as a plain C callback for oracle/ref_render_wrap.cpp (whose Sensor stand-in asks a callback for the ray of
// each sample; perspective.cpp is not part of the pinned build): orc_ref_camera_bind fixes the camera, the function returned
// by orc_ref_camera_callback then fills out16 = o[3] d[3] mint maxt | wavelengths[4] | ray weight[4].  Read-only after bind.
static MskCamera g_callback_camera;
static void ref_camera_callback(float wavelength_sample, float px, float py, float *out16) {
    auto [ray, w] = camera_sample_ray(g_callback_camera, wavelength_sample, { px, py });
    const float v[16] = { ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, ray.mint, ray.maxt,
                          ray.wavelengths[0], ray.wavelengths[1], ray.wavelengths[2], ray.wavelengths[3], w[0], w[1], w[2], w[3] };
    memcpy(out16, v, sizeof(v));
}
int orc_ref_camera_bind(OrcScene *s) {
    if (!s) return fail("null scene");
    g_callback_camera = s->sc.cam;
    return 0;
}
void *orc_ref_camera_callback(void) { return (void *) &ref_camera_callback; }

// AOVIntegrator::sample, integrators/aov.cpp:87-144.  `types` lists the AOVs in channel order; an
// MSK_AOV_INTEGRATOR_RGBA entry is the nested PathTracer (aov.cpp:124-140).  Deviations (oracle.h): fields of a
// missed ray read 0 (uninitialised in the reference, scene.cpp:247-251), and without a nested integrator the
// returned spectrum is 0 (uninitialised `result`, aov.cpp:91,142).
static int aov_channel_count(const int32_t *types, uint32_t ntypes) {
    static const int per[6] = { 1, 3, 2, 3, 3, 4 };
    int n = 0;
    for (uint32_t i = 0; i < ntypes; ++i) {
        if (types[i] < 0 || types[i] > 5) return -1;
        n += per[types[i]];
    }
    return n;
}

static Spec aov_sample(const OScene &sc, Sampler &sampler, const Ray &ray, const PathParams &pp, RayCounters &rc,
                       const int32_t *types, uint32_t ntypes, float *aovs) {
    SceneInteraction si = ray_intersect(sc, ray, rc); // aov.cpp:90
    if (!si.is_valid()) { si.p = V3{ 0, 0, 0 }; si.n = V3{ 0, 0, 0 }; si.uv = V2{ 0, 0 }; si.sh_frame.n = V3{ 0, 0, 0 }; }
    Spec result(0.f);
    size_t ctr = 0;
    for (uint32_t i = 0; i < ntypes; ++i) {
        switch (types[i]) {
            case MSK_AOV_DEPTH: *aovs++ = si.t == Infinity ? 0.f : si.t; break;
            case MSK_AOV_POSITION: *aovs++ = si.p.x; *aovs++ = si.p.y; *aovs++ = si.p.z; break;
            case MSK_AOV_UV: *aovs++ = si.uv.x; *aovs++ = si.uv.y; break;
            case MSK_AOV_GEO_NORMAL: *aovs++ = si.n.x; *aovs++ = si.n.y; *aovs++ = si.n.z; break;
            case MSK_AOV_SH_NORMAL: *aovs++ = si.sh_frame.n.x; *aovs++ = si.sh_frame.n.y; *aovs++ = si.sh_frame.n.z; break;
            case MSK_AOV_INTEGRATOR_RGBA: {
                Spec spec = integrator_sample(sc, sampler, ray, pp, rc);
                float xyz[3];
                spectrum_to_xyz(spec, ray.wavelengths, xyz);
                // xyz_to_srgb, spectrum.h:138-143
                *aovs++ = 3.240479f * xyz[0] + -1.537150f * xyz[1] + -0.498535f * xyz[2];
                *aovs++ = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
                *aovs++ = 0.055648f * xyz[0] + -0.204043f * xyz[1] + 1.057311f * xyz[2];
                *aovs++ = 1.f;
                if (ctr == 0) result = spec;
                ctr++;
            } break;
        }
    }
    return result;
}

// SamplingIntegrator::render, integrator.cpp:31-126.  ntypes == 0 and types == nullptr: the plain PathTracer.
static int render_impl(OrcScene *s, const MskRenderDesc *rd, const int32_t *types, uint32_t ntypes, bool aov, float *film, int nthreads,
                       OrcStats *stats) {
    if (!s || !rd || !film) return fail("null argument");
    const OScene &sc = s->sc;
    const int W = (int) sc.cam.width, H = (int) sc.cam.height;
    int extra = aov ? aov_channel_count(types, ntypes) : 0;
    if (extra < 0) return fail("invalid AOV type");
    const int nch = 5 + extra;
    if (rd->clear_film) std::fill(film, film + (size_t) W * H * nch, 0.f);
    PathParams pp{ rd->max_depth, rd->rr_depth, rd->hide_emitters != 0, (int) rd->integrator };
    // Test-only: seed the sampler the way the reference does when its tile loop runs as one task -- ONE clone seeded
    // with PCG32_DEFAULT_STATE (IndependentSampler(), independent.cpp:11-14) whose sequence runs through every block in
    // spiral order and is never re-seeded -- and draw the BSDF samples in GCC's argument order.  This is how
    // the film of the compiled reference's own SamplingIntegrator::render is replayed (tests/test_oracle_ref_math.py).
    const bool ref_seeding = (rd->flags & ORC_RENDER_REFERENCE_SEEDING) != 0;
    if (ref_seeding) { pp.draw_bsdf_samples_right_to_left = true; nthreads = 1; }
    Sampler ref_sampler;
    ref_sampler.seed(PCG32_DEFAULT_STATE);
    const int border = (int) std::ceil(sc.cam.filter_radius - .5f); // rfilter.cpp:22
    auto blocks = spiral_blocks(W, H, 32);                            // imageblock.h:8, integrator.cpp:48
    std::vector<Block> done(blocks.size());
    std::atomic<size_t> next{ 0 };
    std::atomic<uint64_t> rays_c{ 0 }, rays_s{ 0 };
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        RayCounters rc;
        Sampler own_sampler;
        own_sampler.base_seed = rd->base_seed;
        std::vector<float> aovs((size_t) nch);
        for (;;) {
            size_t bi = next.fetch_add(1);
            if (bi >= blocks.size()) break;
            const BlockDesc &bd = blocks[bi];
            Block &b = done[bi];
            b.ox = bd.ox; b.oy = bd.oy; b.sx = bd.sx; b.sy = bd.sy; b.border = border; b.nch = nch;
            b.data.assign((size_t) (bd.sx + 2 * border) * (bd.sy + 2 * border) * nch, 0.f);
            // render_block, integrator.cpp:82-101
            for (int y = 0; y < bd.sy; ++y)
                for (int x = 0; x < bd.sx; ++x) {
                    int gx = x + bd.ox, gy = y + bd.oy;
                    uint64_t pixel = (uint64_t) gy * W + gx;
                    for (uint32_t smp = rd->sample_begin; smp < rd->sample_end; ++smp) {
                        Sampler &sampler = ref_seeding ? ref_sampler : own_sampler;
                        if (!ref_seeding) sampler.seed(pixel * rd->spp + smp); // determinism contract
                        // render_sample, integrator.cpp:103-126
                        V2 j = sampler.next2d();
                        V2 position_sample{ float(gx) + j.x, float(gy) + j.y };
                        float wavelength_sample = sampler.next1d();
                        sampler.next2d(); // aperture sample: consumed, unused
                        auto [ray, ray_weight] = camera_sample_ray(sc.cam, wavelength_sample, position_sample);
                        Spec result = (aov ? aov_sample(sc, sampler, ray, pp, rc, types, ntypes, aovs.data() + 5)
                                           : integrator_sample(sc, sampler, ray, pp, rc)) * ray_weight;
                        spectrum_to_xyz(result, ray.wavelengths, aovs.data());
                        aovs[3] = 1.f; aovs[4] = 1.f;
                        block_put(b, sc.cam, position_sample, aovs.data());
                    }
                }
        }
        rays_c += rc.closest; rays_s += rc.shadow;
    };
    if (nthreads <= 0) nthreads = (int) std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    // Film::put under the mutex, hdrfilm.cpp:43-46 -- here in spiral order for determinism
    for (const Block &b : done) film_put(film, W, H, b);
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->paths        = (uint64_t) W * H * (rd->sample_end - rd->sample_begin);
        stats->rays_closest = rays_c; stats->rays_shadow = rays_s;
        stats->seconds      = std::chrono::duration<double>(t1 - t0).count();
        stats->threads      = nthreads;
    }
    return 0;
}

int orc_render(OrcScene *s, const MskRenderDesc *rd, float *film, int nthreads, OrcStats *stats) {
    return render_impl(s, rd, nullptr, 0, false, film, nthreads, stats);
}

int orc_aov_channels(const int32_t *types, uint32_t ntypes) { return aov_channel_count(types, ntypes); }

// film: H x W x (5 + orc_aov_channels) floats
int orc_render_aov(OrcScene *s, const MskRenderDesc *rd, const int32_t *types, uint32_t ntypes, float *film, int nthreads, OrcStats *stats) {
    if (ntypes && !types) return fail("null argument");
    return render_impl(s, rd, types, ntypes, true, film, nthreads, stats);
}

// Per-sample radiance for a list of (pixel, sample) pairs: lets the tests compare individual
// paths with the GPU instead of only the filtered film.  out: n x 9 floats
// [pos.x, pos.y, X, Y, Z, L0..L3 (result * ray_weight)].
int orc_trace_samples(OrcScene *s, const MskRenderDesc *rd, const uint32_t *pixel_sample /* n x 2 */, float *out, size_t n) {
    if (!s) return fail("null scene");
    const OScene &sc = s->sc;
    const int W = (int) sc.cam.width;
    PathParams pp{ rd->max_depth, rd->rr_depth, rd->hide_emitters != 0, (int) rd->integrator };
    RayCounters rc;
    Sampler sampler;
    sampler.base_seed = rd->base_seed;
    for (size_t i = 0; i < n; ++i) {
        uint32_t pixel = pixel_sample[i * 2], smp = pixel_sample[i * 2 + 1];
        int gx = pixel % W, gy = pixel / W;
        sampler.seed((uint64_t) pixel * rd->spp + smp);
        V2 j = sampler.next2d();
        V2 position_sample{ float(gx) + j.x, float(gy) + j.y };
        float wavelength_sample = sampler.next1d();
        sampler.next2d();
        auto [ray, ray_weight] = camera_sample_ray(sc.cam, wavelength_sample, position_sample);
        Spec result = integrator_sample(sc, sampler, ray, pp, rc) * ray_weight;
        float xyz[3];
        spectrum_to_xyz(result, ray.wavelengths, xyz);
        float *o = out + i * 9;
        o[0] = position_sample.x; o[1] = position_sample.y; o[2] = xyz[0]; o[3] = xyz[1]; o[4] = xyz[2];
        for (int k = 0; k < 4; ++k) o[5 + k] = result[k];
    }
    return 0;
}

// HDRFilm::image, hdrfilm.cpp:48-90 + xyz_to_srgb, spectrum.h:138-143
void orc_develop(const float *film, float *rgba, size_t npixels) {
    for (size_t i = 0; i < npixels; ++i) {
        const float *p = film + i * 5;
        float r = 3.240479f * p[0] + -1.537150f * p[1] + -0.498535f * p[2];
        float g = -0.969256f * p[0] + 1.875991f * p[1] + 0.041556f * p[2];
        float b = 0.055648f * p[0] + -0.204043f * p[1] + 1.057311f * p[2];
        float weight = p[4], inv_weight = weight != 0 ? 1.f / weight : 0.f;
        rgba[i * 4 + 0] = r * inv_weight; rgba[i * 4 + 1] = g * inv_weight; rgba[i * 4 + 2] = b * inv_weight;
        rgba[i * 4 + 3] = p[3] * inv_weight;
    }
}

// GaussianFilter + ReconstructionFilter::init_discretization: gaussian.cpp:9-20, rfilter.cpp:12-27
void orc_gaussian_filter(float stddev, float *radius, float table[33]) {
    float m_radius = 4 * stddev;
    float alpha    = -1.f / (2.f * stddev * stddev);
    float bias     = std::exp(alpha * m_radius * m_radius);
    float sum      = 0.f;
    for (size_t i = 0; i < FILTER_RES; ++i) {
        float x  = float(m_radius * i) / FILTER_RES;
        table[i] = std::max(0.f, std::exp(alpha * x * x) - bias);
        sum += table[i];
    }
    table[FILTER_RES] = 0;
    sum *= 2 * m_radius / FILTER_RES;
    float normalization = 1.0f / sum;
    for (size_t i = 0; i < FILTER_RES; ++i) table[i] *= normalization;
    *radius = m_radius;
}

// ---- unit-level entry points for the known-answer tests ----
void orc_pcg32_floats(uint64_t seed, uint64_t base_seed, float *out, size_t n) {
    Sampler s; s.base_seed = base_seed; s.seed(seed);
    for (size_t i = 0; i < n; ++i) out[i] = s.next1d();
}
void orc_pcg32_uints(uint64_t initstate, uint64_t initseq, uint32_t *out, size_t n) {
    PCG32 r; r.seed(initstate, initseq);
    for (size_t i = 0; i < n; ++i) out[i] = r.next_uint32();
}
void orc_sample_wavelength(float u, float wl[4], float w[4]) {
    Spec a, b; sample_wavelength(u, a, b);
    for (int i = 0; i < 4; ++i) { wl[i] = a[i]; w[i] = b[i]; }
}
void orc_warp(int which, float u, float v, float out[3]) {
    V2 s{ u, v };
    if (which == 0) { V2 r = square_to_uniform_triangle(s); out[0] = r.x; out[1] = r.y; out[2] = 0; }
    else if (which == 1) { V2 r = square_to_uniform_disk_concentric(s); out[0] = r.x; out[1] = r.y; out[2] = 0; }
    else if (which == 2) { V3 r = square_to_cosine_hemisphere(s); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
    else { V3 r = square_to_uniform_sphere(s); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
}
void orc_fresnel(float cos_theta_i, float eta, float out[4]) {
    FresnelResult f = fresnel(cos_theta_i, eta);
    out[0] = f.F; out[1] = f.cos_theta_t; out[2] = f.eta_it; out[3] = f.eta_ti;
}
void orc_fresnel_conductor(float cos_theta_i, const float eta[4], const float k[4], float out[4]) {
    Spec e, kk; for (int i = 0; i < 4; ++i) { e[i] = eta[i]; kk[i] = k[i]; }
    Spec r = fresnel_conductor(cos_theta_i, e, kk);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
}
void orc_srgb_model_eval(const float c[3], const float wl[4], float out[4]) {
    Spec w; for (int i = 0; i < 4; ++i) w[i] = wl[i];
    Spec r = srgb_model_eval(c, w);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
}
int orc_texture_eval(OrcScene *s, int id, float u, float v, const float wl[4], float out[4]) {
    if (!s || id < 0 || (size_t) id >= s->sc.spectra.size()) return fail("bad spectrum id");
    Spec w; for (int i = 0; i < 4; ++i) w[i] = wl[i];
    Spec r = texture_eval(s->sc.spectra, id, u, v, w);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
    return 0;
}
int orc_spectrum_eval(OrcScene *s, int id, const float wl[4], float out[4]) {
    if (!s || id < 0 || (size_t) id >= s->sc.spectra.size()) return fail("bad spectrum id");
    Spec w; for (int i = 0; i < 4; ++i) w[i] = wl[i];
    Spec r = spectrum_eval(s->sc.spectra[id], w);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
    return 0;
}
void orc_spectrum_to_xyz(const float value[4], const float wl[4], float xyz[3]) {
    Spec v, w; for (int i = 0; i < 4; ++i) { v[i] = value[i]; w[i] = wl[i]; }
    spectrum_to_xyz(v, w, xyz);
}
void orc_ggx(int which, float au, float av, const float a[3], const float b[3], float out[4]) {
    Microfacet d(au, av);
    V3 va(a[0], a[1], a[2]), vb(b[0], b[1], b[2]);
    if (which == 0) out[0] = d.eval(va);
    else if (which == 1) { auto [m, pdf] = d.sample(va, { b[0], b[1] }); out[0] = m.x; out[1] = m.y; out[2] = m.z; out[3] = pdf; }
    else if (which == 2) out[0] = d.smith_g1(va, vb);
}
// Remaining math helpers in isolation, for the golden vectors produced by the compiled reference headers
// (tests/test_oracle_ref_math.py).  which: 0 coordinate_system(n = in[0..2]) -> s, t            (mathutils.h:196-203)
//   1 Frame(n = in[0..2]).to_local(v = in[3..5]) | to_world(v)                                  (frame.h:16-26)
//   2 reflect(wi = in[0..2], m = in[3..5]) | refract(wi, m, cos_theta_t = in[6], eta_ti = in[7]) (fresnel.h:16-34)
//   3 ggx(au = in[0], av = in[1]): pdf(wi = in[2..4], m = in[5..7]), G(wi = in[2..4], wo = in[5..7], m = in[8..10])
//   4 xyz_to_srgb(in[0..2])                                                                      (spectrum.h:138-143)
void orc_math(int which, const float *in, float *out) {
    auto v3 = [&](int o) { return V3(in[o], in[o + 1], in[o + 2]); };
    auto put = [&](int o, V3 v) { out[o] = v.x; out[o + 1] = v.y; out[o + 2] = v.z; };
    if (which == 0) { V3 s, t; coordinate_system(v3(0), s, t); put(0, s); put(3, t); }
    else if (which == 1) { Frame f(v3(0)); put(0, f.to_local(v3(3))); put(3, f.to_world(v3(3))); }
    else if (which == 2) { put(0, reflect(v3(0), v3(3))); put(3, refract(v3(0), v3(3), in[6], in[7])); }
    else if (which == 3) { Microfacet d(in[0], in[1]); out[0] = d.pdf(v3(2), v3(5)); out[1] = d.G(v3(2), v3(5), v3(8)); }
    else if (which == 4) {
        float film[5] = { in[0], in[1], in[2], 1.f, 1.f }, rgba[4];
        orc_develop(film, rgba, 1);
        out[0] = rgba[0]; out[1] = rgba[1]; out[2] = rgba[2];
    }
}
// PathTracer::sample / VolumetricPathTracer::sample for ONE given camera ray with the sampler seeded as
// IndependentSampler::seed(seed): the entry point the golden vectors of the compiled reference path tracer are replayed
// through (tests/test_oracle_ref_math.py)
int orc_sample_ray(OrcScene *s, const MskRenderDesc *rd, uint64_t seed, const float o[3], const float d[3], float mint, float maxt, const float wl[4],
                   int bsdf_draws_right_to_left, float out[4]) {
    if (!s || !rd) return fail("null argument");
    PathParams pp{ rd->max_depth, rd->rr_depth, rd->hide_emitters != 0, (int) rd->integrator, bsdf_draws_right_to_left != 0 };
    RayCounters rc;
    Sampler sampler;
    sampler.base_seed = rd->base_seed;
    sampler.seed(seed);
    Spec w; for (int i = 0; i < 4; ++i) w[i] = wl[i];
    Ray ray{ V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), mint, maxt, w };
    Spec r = integrator_sample(s->sc, sampler, ray, pp, rc);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
    return 0;
}
// The film accumulation in isolation, driven as SamplingIntegrator::render drives it (layouts as oracle/ref_film_wrap.cpp):
// spiral block order, per block the filtered splat of the samples whose pixel lies in it, then the merge into the film
int orc_film_accumulate(float stddev, int W, int H, int nch, int block_size, const float *samples, size_t n, float *film_out, int *block_order) {
    MskCamera cam{};
    cam.width = (uint32_t) W; cam.height = (uint32_t) H;
    orc_gaussian_filter(stddev, &cam.filter_radius, cam.filter_table);
    const int border = (int) std::ceil(cam.filter_radius - .5f);
    std::fill(film_out, film_out + (size_t) W * H * nch, 0.f);
    std::vector<BlockDesc> blocks = spiral_blocks(W, H, block_size);
    const size_t stride = 2 + (size_t) nch;
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
        const BlockDesc &bd = blocks[bi];
        if (block_order) { block_order[4 * bi] = bd.ox; block_order[4 * bi + 1] = bd.oy; block_order[4 * bi + 2] = bd.sx; block_order[4 * bi + 3] = bd.sy; }
        Block b;
        b.ox = bd.ox; b.oy = bd.oy; b.sx = bd.sx; b.sy = bd.sy; b.border = border; b.nch = nch;
        b.data.assign((size_t) (bd.sx + 2 * border) * (bd.sy + 2 * border) * nch, 0.f);
        for (size_t i = 0; i < n; ++i) {
            const float *sm = samples + i * stride;
            int px = (int) std::floor(sm[0]), py = (int) std::floor(sm[1]);
            if (px < bd.ox || py < bd.oy || px >= bd.ox + bd.sx || py >= bd.oy + bd.sy) continue;
            block_put(b, cam, V2{ sm[0], sm[1] }, sm + 2);
        }
        film_put(film_out, W, H, b);
    }
    return 0;
}
// AOVIntegrator::sample for one camera ray (same conventions as orc_sample_ray); out_aovs: the channels of `types` in order
int orc_aov_sample_ray(OrcScene *s, const MskRenderDesc *rd, uint64_t seed, const float o[3], const float d[3], float mint, float maxt,
                       const float wl[4], int bsdf_draws_right_to_left, const int32_t *types, uint32_t ntypes, float *out_aovs, float out[4]) {
    if (!s || !rd || (ntypes && !types)) return fail("null argument");
    PathParams pp{ rd->max_depth, rd->rr_depth, rd->hide_emitters != 0, (int) rd->integrator, bsdf_draws_right_to_left != 0 };
    RayCounters rc;
    Sampler sampler;
    sampler.base_seed = rd->base_seed;
    sampler.seed(seed);
    Spec w; for (int i = 0; i < 4; ++i) w[i] = wl[i];
    Ray ray{ V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), mint, maxt, w };
    Spec r = aov_sample(s->sc, sampler, ray, pp, rc, types, ntypes, out_aovs);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
    return 0;
}
// Hit reconstruction and mesh sampling on a single mesh given as raw arrays (layouts as oracle/ref_mesh_wrap.cpp):
// interaction out[27] = t | p | n | uv | sh_frame.s | sh_frame.t | sh_frame.n | wi | dp_du | dp_dv
static OMesh make_single_mesh(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, int normals, int uvs) {
    OMesh o;
    o.nverts = nverts; o.ntris = ntris; o.has_normals = normals != 0; o.has_uvs = uvs != 0;
    o.verts.assign(verts, verts + (size_t) nverts * 8);
    o.tris.assign(tris, tris + (size_t) ntris * 3);
    area_distr_build(o);
    return o;
}
void orc_mesh_interaction(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, int normals, int uvs, uint32_t prim, float u,
                          float v, float t, const float o[3], const float d[3], float out[27]) {
    OScene sc;
    sc.meshes.push_back(make_single_mesh(verts, nverts, tris, ntris, normals, uvs));
    Ray ray{ V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), 0.f, Infinity, Spec(500.f) };
    RawHit h; h.t = t; h.u = u; h.v = v; h.prim = prim; h.geom = 0;
    SceneInteraction si = compute_scene_interaction(sc, ray, h);
    float *w = out;
    auto put3 = [&](V3 x) { *w++ = x.x; *w++ = x.y; *w++ = x.z; };
    *w++ = si.t; put3(si.p); put3(si.n); *w++ = si.uv.x; *w++ = si.uv.y;
    put3(si.sh_frame.s); put3(si.sh_frame.t); put3(si.sh_frame.n); put3(si.wi); put3(si.dp_du); put3(si.dp_dv);
}
// sampling out[22] = ps.p | ps.n | ps.uv | ps.pdf | ds.p | ds.n | ds.d | ds.dist | ds.pdf | pdf_direct(ds) | surface_area
void orc_mesh_sampling(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, int normals, int uvs, const float sample[2],
                       const float ref_p[3], float out[22], float *cdf_out) {
    OMesh m = make_single_mesh(verts, nverts, tris, ntris, normals, uvs);
    PositionSample ps = mesh_sample_position(m, { sample[0], sample[1] });
    SceneInteraction si;
    si.p = V3(ref_p[0], ref_p[1], ref_p[2]);
    DirectIllumSample ds = shape_sample_direct(m, si, { sample[0], sample[1] });
    float *w = out;
    auto put3 = [&](V3 x) { *w++ = x.x; *w++ = x.y; *w++ = x.z; };
    put3(ps.p); put3(ps.n); *w++ = ps.uv.x; *w++ = ps.uv.y; *w++ = ps.pdf;
    put3(ds.p); put3(ds.n); put3(ds.d); *w++ = ds.dist; *w++ = ds.pdf; *w++ = shape_pdf_direct(m, ds); *w++ = m.surface_area;
    if (cdf_out) std::copy(m.cdf.begin(), m.cdf.end(), cdf_out);
}
// Distribution1D::init + sample_reuse as mesh_sample_position uses them (distribution.h:88-123): cdf has n + 1 entries
void orc_distribution_sample_reuse(const float *pdf, size_t n, const float *u, size_t nu, uint32_t *index, float *reused, float *cdf_out) {
    std::vector<float> cdf = distribution_cdf(pdf, n);
    for (size_t i = 0; i < nu; ++i) std::tie(index[i], reused[i]) = distribution_sample_reuse(cdf, u[i]);
    if (cdf_out) std::copy(cdf.begin(), cdf.end(), cdf_out);
}
// BSDF in isolation: wi (local), samples -> [wo.xyz, pdf, eta, sampled_type, weight0..3]; eval/pdf for a given wo
int orc_bsdf(OrcScene *s, int bsdf_id, const float wi[3], const float wl[4], const float smp[3], const float wo_in[3], float out_sample[10], float out_eval[4], float *out_pdf) {
    if (!s || bsdf_id < 0 || (size_t) bsdf_id >= s->sc.bsdfs.size()) return fail("bad bsdf id");
    SceneInteraction si;
    si.t = 1.f; si.wi = V3(wi[0], wi[1], wi[2]);
    for (int i = 0; i < 4; ++i) si.wavelengths[i] = wl[i];
    const MskBsdf &b = s->sc.bsdfs[bsdf_id];
    auto [bs, w] = bsdf_sample(s->sc, b, si, smp[0], { smp[1], smp[2] });
    out_sample[0] = bs.wo.x; out_sample[1] = bs.wo.y; out_sample[2] = bs.wo.z; out_sample[3] = bs.pdf; out_sample[4] = bs.eta;
    out_sample[5] = (float) bs.sampled_type;
    for (int i = 0; i < 4; ++i) out_sample[6 + i] = w[i];
    V3 wo(wo_in[0], wo_in[1], wo_in[2]);
    Spec e = bsdf_eval(s->sc, b, si, wo);
    for (int i = 0; i < 4; ++i) out_eval[i] = e[i];
    *out_pdf = bsdf_pdf(b, si, wo);
    return 0;
}

// ---- rgb2spec_fetch restated: ext/rgb2spec/rgb2spec.c:12-47 (load), :59-75 (interval), :77-119 (fetch) ----
struct OrcRgb2Spec { uint32_t res; std::vector<float> scale, data; };

OrcRgb2Spec *orc_rgb2spec_load(const char *filename) {
    FILE *f = std::fopen(filename, "rb");
    if (!f) { fail("cannot open coefficient file"); return nullptr; }
    char header[4];
    auto m = std::make_unique<OrcRgb2Spec>();
    bool ok = std::fread(header, 4, 1, f) == 1 && std::memcmp(header, "SPEC", 4) == 0 && std::fread(&m->res, 4, 1, f) == 1;
    if (ok) {
        m->scale.resize(m->res);
        m->data.resize((size_t) m->res * m->res * m->res * 9);
        ok = std::fread(m->scale.data(), 4, m->scale.size(), f) == m->scale.size() &&
             std::fread(m->data.data(), 4, m->data.size(), f) == m->data.size();
    }
    std::fclose(f);
    if (!ok) { fail("malformed coefficient file"); return nullptr; }
    return m.release();
}
void orc_rgb2spec_free(OrcRgb2Spec *m) { delete m; }
void orc_rgb2spec_fetch(const OrcRgb2Spec *model, const float rgb_[3], float out[3]) {
    int i = 0, res = (int) model->res;
    float rgb[3];
    for (int j = 0; j < 3; ++j) rgb[j] = std::max(std::min(rgb_[j], 1.f), 0.f);
    for (int j = 1; j < 3; ++j) if (rgb[j] >= rgb[i]) i = j;
    float z = rgb[i], scale = (res - 1) / z, x = rgb[(i + 1) % 3] * scale, y = rgb[(i + 2) % 3] * scale;
    // rgb2spec_find_interval
    int left = 0, last_interval = res - 2, size = last_interval;
    while (size > 0) {
        int half = size >> 1, middle = left + half + 1;
        if (model->scale[middle] <= z) { left = middle; size -= half + 1; } else size = half;
    }
    uint32_t zi = (uint32_t) std::min(left, last_interval);
    uint32_t xi = std::min((uint32_t) x, (uint32_t) (res - 2)), yi = std::min((uint32_t) y, (uint32_t) (res - 2));
    uint32_t offset = (((i * res + zi) * res + yi) * res + xi) * 3, dx = 3, dy = 3 * res, dz = 3 * res * res;
    float x1 = x - xi, x0 = 1.f - x1, y1 = y - yi, y0 = 1.f - y1;
    float z1 = (z - model->scale[zi]) / (model->scale[zi + 1] - model->scale[zi]), z0 = 1.f - z1;
    const float *D = model->data.data();
    for (int j = 0; j < 3; ++j) {
        out[j] = ((D[offset] * x0 + D[offset + dx] * x1) * y0 + (D[offset + dy] * x0 + D[offset + dy + dx] * x1) * y1) * z0 +
                 ((D[offset + dz] * x0 + D[offset + dz + dx] * x1) * y0 + (D[offset + dz + dy] * x0 + D[offset + dz + dy + dx] * x1) * y1) * z1;
        offset++;
    }
}

} // extern "C"
