// C entry points over the reference's OWN src/librender/{mesh,shape,records,interaction}.cpp, #included from where they
// lie over the stand-ins under oracle/ref_shim/ (oracle/Makefile.ref): hit reconstruction
// (Mesh::compute_scene_interaction + PreliminaryIntersection::compute_scene_interaction + initialize_sh_frame), the area
// distribution, Mesh::sample_position, Shape::sample_direct / pdf_direct.  TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "msk_ref_prelude.h"
#include <misaki/render/bsdf.h>
#include <misaki/render/interaction.h>
#include <misaki/render/records.h>
#include <misaki/render/mesh.h>
#include "ref_wrap_common.h"
#include <shape.cpp>
#include <mesh.cpp>
#include <records.cpp>
#include <interaction.cpp>

using namespace misaki;

extern "C" {

// out (27 floats + dn_du, dn_dv): t | p | n | uv | sh_frame.s | sh_frame.t | sh_frame.n | wi | dp_du | dp_dv | dn_du | dn_dv  (the last two: zeros
// unless the mesh has vertex normals)
int ref_mesh_interaction(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, int normals, int uvs, uint32_t prim, float u,
                         float v, float t, const float o[3], const float d[3], float out[36]) {
    try {
        RefMesh m(verts, nverts, tris, ntris, normals != 0, uvs != 0);
        PreliminaryIntersection pi;
        pi.t = t; pi.prim_uv = Eigen::Vector2f(u, v); pi.prim_index = prim; pi.shape_index = 0; pi.shape = &m;
        Ray ray(Eigen::Vector3f(o[0], o[1], o[2]), Eigen::Vector3f(d[0], d[1], d[2]), 0.f, Wavelength(400.f, 500.f, 600.f, 700.f));
        SceneInteraction si = pi.compute_scene_interaction(ray);
        float *w = out;
        *w++ = si.t;
        auto put3 = [&](const Eigen::Vector3f &x) { *w++ = x.x(); *w++ = x.y(); *w++ = x.z(); };
        put3(si.p); put3(si.n); *w++ = si.uv.x(); *w++ = si.uv.y();
        put3(si.sh_frame.s); put3(si.sh_frame.t); put3(si.sh_frame.n); put3(si.wi); put3(si.dp_du); put3(si.dp_dv);
        if (normals) { put3(si.dn_du); put3(si.dn_dv); } else { for (int i = 0; i < 6; ++i) *w++ = 0.f; }
        return 0;
    } catch (...) { return -2; }
}
// Mesh::sample_position(sample) and Shape::sample_direct(ref_p, sample) + pdf_direct; cdf_out: ntris + 1 entries.
// out (22 floats): ps.p | ps.n | ps.uv | ps.pdf | ds.p | ds.n | ds.d | ds.dist | ds.pdf | pdf_direct(ds) | surface_area
int ref_mesh_sampling(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, int normals, int uvs, const float sample[2],
                      const float ref_p[3], float out[22], float *cdf_out) {
    try {
        RefMesh m(verts, nverts, tris, ntris, normals != 0, uvs != 0);
        PositionSample ps = m.sample_position(Eigen::Vector2f(sample[0], sample[1]));
        SceneInteraction si;
        si.p = Eigen::Vector3f(ref_p[0], ref_p[1], ref_p[2]);
        DirectIllumSample ds = m.sample_direct(si, Eigen::Vector2f(sample[0], sample[1]));
        float *w = out;
        auto put3 = [&](const Eigen::Vector3f &x) { *w++ = x.x(); *w++ = x.y(); *w++ = x.z(); };
        put3(ps.p); put3(ps.n); *w++ = ps.uv.x(); *w++ = ps.uv.y(); *w++ = ps.pdf;
        put3(ds.p); put3(ds.n); put3(ds.d); *w++ = ds.dist; *w++ = ds.pdf; *w++ = m.pdf_direct(ds); *w++ = m.surface_area();
        if (cdf_out) for (size_t i = 0; i < m.cdf().size(); ++i) cdf_out[i] = m.cdf()[i];
        return 0;
    } catch (...) { return -2; }
}

} // extern "C"
