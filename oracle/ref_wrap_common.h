// Helper classes shared by the oracle/ref_*_wrap.cpp translation units.  TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#pragma once
#include "msk_ref_prelude.h"
#include <misaki/render/interaction.h>
#include <misaki/render/mesh.h>
#include <misaki/render/texture.h>

namespace misaki {
// a constant spectrum stands for every spectral parameter: the arithmetic of the code under test is what is compared
class ConstTexture final : public Texture {
public:
    explicit ConstTexture(float v) : Texture(Properties()), m_value(v) {}
    float eval_1(const SceneInteraction &) const override { return m_value; }
    Spectrum eval(const SceneInteraction &) const override { return Spectrum::Constant(m_value); }
    Color3 eval_3(const SceneInteraction &) const override { return Color3::Constant(m_value); }
    float mean() const override { return m_value; }
    std::string to_string() const override { return "ConstTexture"; }
private:
    float m_value;
};
inline ref<Texture> make_const(float v) { return ref<Texture>(new ConstTexture(v)); }

// Mesh's constructor and buffers are protected: the loader plugins (shapes/obj.cpp:137-177) fill them like this
class RefMesh final : public Mesh {
public:
    RefMesh(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, bool normals, bool uvs, const Properties &props = Properties())
        : Mesh(props) {
        m_vertex_size = 8; m_face_size = 3; // [px py pz nx ny nz u v], obj.cpp:139-142
        m_normal_offset = normals ? 3 : 0; m_texcoord_offset = uvs ? 6 : 0;
        m_vertex_count = nverts; m_face_count = ntris;
        m_vertices = std::unique_ptr<float[]>(new float[(size_t) nverts * 8 + 1]);
        m_faces = std::unique_ptr<uint32_t[]>(new uint32_t[(size_t) ntris * 3 + 1]);
        memcpy(m_vertices.get(), verts, sizeof(float) * nverts * 8);
        memcpy(m_faces.get(), tris, sizeof(uint32_t) * ntris * 3);
        m_surface_area = 0.f; // never initialised by the reference (mesh.h:93); restated as 0 like the oracle
        area_distr_build();
        recompute_bbox(); // Mesh::Mesh already ran set_children() (mesh.cpp:14); the bbox must follow the vertex upload
    }
    std::string to_string() const override { return "RefMesh"; }
    const std::vector<float> &cdf() const { return m_area_distr.cdf(); }
};
} // namespace misaki
