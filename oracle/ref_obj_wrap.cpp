// C entry points over the reference's OWN src/librender/shapes/obj.cpp (the OBJ loader: vertex de-duplication, quad
// split, texcoord flip, the 8-float vertex layout), #included from where it lies over the stand-ins under
// oracle/ref_shim/ (identity to_world).  TEST INFRASTRUCTURE, see ref_math_wrap.cpp.
#include "msk_ref_prelude.h"
#include <misaki/render/mesh.h>
#include <shapes/obj.cpp>

using namespace misaki;

extern "C" {

void *ref_obj_load(const char *filename, int flip_tex_coords) {
    try {
        Properties p;
        p.strings["filename"] = filename;
        p.bools["filp_tex_coords"] = flip_tex_coords != 0; // (sic) obj.cpp:59
        return new OBJMesh(p);
    } catch (...) { return nullptr; }
}
// counts[4] = vertex_count, face_count, has_vertex_normals, has_vertex_texcoords; verts: vertex_count x 8; faces: face_count x 3
void ref_obj_get(void *handle, uint32_t counts[4], float *verts, uint32_t *faces) {
    const Mesh *m = (const OBJMesh *) handle;
    counts[0] = m->vertex_count(); counts[1] = m->face_count(); counts[2] = m->has_vertex_normals(); counts[3] = m->has_vertex_texcoords();
    if (verts) memcpy(verts, m->vertices(), sizeof(float) * 8 * m->vertex_count());
    if (faces) memcpy(faces, m->faces(), sizeof(uint32_t) * 3 * m->face_count());
}

} // extern "C"
