// C entry point over the reference's OWN path tracer: PathTracer::sample (src/librender/integrators/path.cpp) running on
// the reference's own Scene::sample_emitter_direct / pdf_emitter_direct (scene.cpp:69-112), AreaLight (emitters/area.cpp),
// Emitter (emitter.cpp), SmoothDiffuse (bsdfs/diffuse.cpp), Mesh / Shape / interaction code and IndependentSampler, all
// #included from where they lie over the stand-ins under oracle/ref_shim/ (oracle/Makefile.ref).
// TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
//
// What is NOT the reference here, and why:
//   * Scene::ray_intersect / ray_test.  In the reference they are Embree calls fenced by MSK_ENABLE_EMBREE
//     (scene.cpp:197-275); Embree is outside the reference tree and unavailable.  They are restated below around a
//     brute-force Moeller-Trumbore loop -- the SAME routine the oracle uses as its intersection ground truth -- keeping
//     everything scene.cpp does around the Embree call (hit <=> tfar != maxt, the PreliminaryIntersection hand-over).
//     The comparison therefore isolates the integrator / emitter / BSDF / interaction code, not the intersector.
//   * Texture::D65 (texture.cpp needs the plugin manager; never reached: every emitter gets an explicit radiance)
#include "ref_wrap_common.h"
#include <misaki/render/bsdf.h>
#include <misaki/render/emitter.h>
#include <misaki/render/integrator.h>
#include <misaki/render/sampler.h>
#include <misaki/render/scene.h>
#include <bsdfs/diffuse.cpp>        // class SmoothDiffuse (in-class members only; also compiled in ref_bsdf_wrap.cpp)
#include <samplers/independent.cpp> // class IndependentSampler (ditto, ref_plugins_wrap.cpp)
#include <emitter.cpp>
#include <emitters/area.cpp>
#include <emitters/constant.cpp>
#include <scene.cpp>
#include <integrators/path.cpp>
#include <integrators/aov.cpp>

namespace misaki {

// (SamplingIntegrator / MonteCarloIntegrator members come from integrator.cpp itself, compiled in ref_render_wrap.cpp)
ref<Texture> Texture::D65(float) { return make_const(1.f); }

// ---- Embree stand-in (see the header comment)
static bool tri_intersect(const Eigen::Vector3f &v0, const Eigen::Vector3f &v1, const Eigen::Vector3f &v2, const Eigen::Vector3f &o,
                          const Eigen::Vector3f &d, float tnear, float tfar, float &t, float &u, float &v) {
    Eigen::Vector3f e1 = v0 - v1, e2 = v2 - v0, Ng = e2.cross(e1);
    Eigen::Vector3f C = v0 - o, R = C.cross(d);
    float den = Ng.dot(d);
    if (den == 0.f) return false;
    float absden = std::abs(den), sgn = den < 0.f ? -1.f : 1.f;
    float U = R.dot(e2) * sgn, V = R.dot(e1) * sgn;
    if (!(U >= 0.f) || !(V >= 0.f) || !(U + V <= absden)) return false;
    float T = Ng.dot(C) * sgn;
    if (!(absden * tnear < T) || !(T <= absden * tfar)) return false;
    float rcp = 1.f / absden;
    u = U * rcp; v = V * rcp; t = T * rcp;
    return true;
}
struct BruteHit { float t, u, v; uint32_t prim, geom; bool found; };
static BruteHit brute_force(const std::vector<ref<Shape>> &shapes, const Ray &ray, bool any) {
    BruteHit best{ math::Infinity<float>, 0.f, 0.f, 0xffffffffu, 0xffffffffu, false };
    for (uint32_t g = 0; g < shapes.size(); ++g) {
        const Mesh *m = static_cast<const Mesh *>(shapes[g].get());
        for (uint32_t p = 0; p < m->face_count(); ++p) {
            Eigen::Vector3f fi = m->face_indices(p);
            float t, u, v;
            if (tri_intersect(m->vertex_position((uint32_t) fi[0]), m->vertex_position((uint32_t) fi[1]), m->vertex_position((uint32_t) fi[2]), ray.o,
                              ray.d, ray.mint, ray.maxt, t, u, v)) {
                if (any) { best.found = true; best.t = t; return best; }
                if (t < best.t) best = BruteHit{ t, u, v, p, g, true }; // ties -> lowest (geom, prim): first in loop order
            }
        }
    }
    return best;
}
void Scene::accel_init(const Properties &) {}
void Scene::accel_release() {}
SceneInteraction Scene::ray_intersect(const Ray &ray) const { // scene.cpp:216-253 around the intersector call
    BruteHit h = brute_force(m_shapes, ray, false);
    float tfar = h.found ? h.t : ray.maxt;
    SceneInteraction si;
    if (tfar != ray.maxt) {
        PreliminaryIntersection pi;
        pi.shape_index = h.geom;
        pi.shape       = m_shapes[h.geom];
        pi.t           = tfar;
        pi.prim_index  = h.prim;
        pi.prim_uv     = Eigen::Vector2f(h.u, h.v);
        si = pi.compute_scene_interaction(ray);
    } else {
        si.wavelengths = ray.wavelengths;
        si.wi          = -ray.d;
        si.t           = math::Infinity<float>;
    }
    return si;
}
bool Scene::ray_test(const Ray &ray) const { // scene.cpp:255-273
    BruteHit h = brute_force(m_shapes, ray, true);
    float tfar = h.found ? h.t : ray.maxt;
    return tfar != ray.maxt;
}

} // namespace misaki

using namespace misaki;

#include "ref_path_scene.h"
#include <functional>
misaki::Texture *msk_ref_make_rgb_texture(const float rgb[3], bool within_emitter); // ref_spectra_wrap.cpp
misaki::SamplingIntegrator *msk_ref_path_tracer(RefPathScene *s) { return s->tracer; }
misaki::Scene *msk_ref_scene(RefPathScene *s) { return s->scene; }

extern "C" {

// Meshes as in MskSceneDesc (verts nverts x 8, tris ntris x 3), each with a diffuse reflectance (constant spectrum) and an
// optional area-light radiance (constant spectrum, < 0: none); Scene::m_shapes order == the order given.  env_radiance >= 0:
// a "constant" environment emitter after the shapes (emitters/constant.cpp).
static void *create_scene(uint32_t nmeshes, const float *const *verts, const uint32_t *nverts, const uint32_t *const *tris, const uint32_t *ntris,
                          const int *normals, const int *uvs, const std::function<ref<Texture>(uint32_t)> &reflectance,
                          const std::function<ref<Texture>(uint32_t)> &radiance /* null reference: no area light */, float env_radiance) {
    try {
        RefPathScene *s = new RefPathScene;
        Properties sp;
        for (uint32_t i = 0; i < nmeshes; ++i) {
            Properties bp;
            bp.make_default = make_const;
            bp.textures["reflectance"] = reflectance(i);
            Properties mp;
            mp.children.push_back({ "_arg_0", ref<Object>(new SmoothDiffuse(bp)) });
            if (ref<Texture> rad = radiance(i)) {
                Properties ep;
                ep.textures["radiance"] = rad;
                mp.children.push_back({ "_arg_1", ref<Object>(new AreaLight(ep)) });
            }
            RefMesh *m = new RefMesh(verts[i], nverts[i], tris[i], ntris[i], normals[i] != 0, uvs[i] != 0, mp);
            s->meshes.push_back(m);
            sp.children.push_back({ "_arg_" + std::to_string(i), ref<Object>(m) });
        }
        if (env_radiance >= 0.f) {
            Properties ep;
            ep.textures["radiance"] = make_const(env_radiance);
            sp.children.push_back({ "_arg_env", ref<Object>(new ConstantBackgroundEmitter(ep)) });
        }
        s->scene = new RefScene(sp);
        s->tracer = new PathTracer(Properties());
        return s;
    } catch (...) { return nullptr; }
}
void *ref_path_scene_create(uint32_t nmeshes, const float *const *verts, const uint32_t *nverts, const uint32_t *const *tris, const uint32_t *ntris,
                            const int *normals, const int *uvs, const float *reflectance, const float *radiance, float env_radiance) {
    return create_scene(nmeshes, verts, nverts, tris, ntris, normals, uvs, [&](uint32_t i) { return make_const(reflectance[i]); },
                        [&](uint32_t i) { return radiance[i] >= 0.f ? make_const(radiance[i]) : ref<Texture>(); }, env_radiance);
}
// The same with the textures an <rgb> tag becomes (xml.cpp:269-277): reflectance_rgb / radiance_rgb hold 3 floats per mesh,
// a negative radiance_rgb[3 i] means no area light -- the scenes of the reference's own XML files (assets/cbox/scene.xml).
void *ref_path_scene_create_rgb(uint32_t nmeshes, const float *const *verts, const uint32_t *nverts, const uint32_t *const *tris, const uint32_t *ntris,
                                const int *normals, const int *uvs, const float *reflectance_rgb, const float *radiance_rgb) {
    return create_scene(nmeshes, verts, nverts, tris, ntris, normals, uvs,
                        [&](uint32_t i) { return ref<Texture>(msk_ref_make_rgb_texture(reflectance_rgb + 3 * i, false)); },
                        [&](uint32_t i) { return radiance_rgb[3 * i] >= 0.f ? ref<Texture>(msk_ref_make_rgb_texture(radiance_rgb + 3 * i, true)) : ref<Texture>(); },
                        -1.f);
}
// PathTracer::sample for one camera ray with the sampler seeded as IndependentSampler::seed(seed) (base_seed 0);
// max_depth -1 / rr_depth 5 / hide_emitter false are hard-wired in the reference (path.cpp:135-136)
int ref_path_sample(void *handle, uint64_t seed, const float o[3], const float d[3], float mint, float maxt, const float wl[4], float out[4]) {
    try {
        RefPathScene *s = (RefPathScene *) handle;
        IndependentSampler sampler;
        sampler.seed(seed);
        Ray ray(Eigen::Vector3f(o[0], o[1], o[2]), Eigen::Vector3f(d[0], d[1], d[2]), mint, maxt, 0.f, Wavelength(wl[0], wl[1], wl[2], wl[3]));
        Spectrum r = s->tracer->sample(s->scene, &sampler, RayDifferential(ray), nullptr, nullptr);
        for (int i = 0; i < 4; ++i) out[i] = r.coeff(i);
        return 0;
    } catch (...) { return -2; }
}

// AOVIntegrator::sample (integrators/aov.cpp:87-144) with aovs = "d:depth p:position u:uv g:geo_normal s:sh_normal" and
// the path tracer nested: out = 12 geometry channels + R, G, B, A of the nested integrator + the 4 returned radiance values
int ref_aov_sample(void *handle, uint64_t seed, const float o[3], const float d[3], float mint, float maxt, const float wl[4], float out[20]) {
    try {
        RefPathScene *s = (RefPathScene *) handle;
        Properties p;
        p.strings["aovs"] = "d:depth p:position u:uv g:geo_normal s:sh_normal";
        p.children.push_back({ "img", ref<Object>(s->tracer) });
        AOVIntegrator aov(p);
        IndependentSampler sampler;
        sampler.seed(seed);
        Ray ray(Eigen::Vector3f(o[0], o[1], o[2]), Eigen::Vector3f(d[0], d[1], d[2]), mint, maxt, 0.f, Wavelength(wl[0], wl[1], wl[2], wl[3]));
        for (int i = 0; i < 20; ++i) out[i] = 0.f;
        Spectrum r = aov.sample(s->scene, &sampler, RayDifferential(ray), nullptr, out);
        for (int i = 0; i < 4; ++i) out[16 + i] = r.coeff(i);
        return 0;
    } catch (...) { return -2; }
}

} // extern "C"

// The AOV integrator of ref_aov_sample as an object, for the whole-film comparison of ref_render_wrap.cpp
misaki::SamplingIntegrator *msk_ref_make_aov(RefPathScene *s) {
    misaki::Properties p;
    p.strings["aovs"] = "d:depth p:position u:uv g:geo_normal s:sh_normal";
    p.children.push_back({ "img", misaki::ref<misaki::Object>(s->tracer) });
    return new misaki::AOVIntegrator(p);
}
