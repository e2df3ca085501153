/* TEST INFRASTRUCTURE -- C API of the CPU oracle (oracle/oracle.cpp).  Not product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load liboracle.so.  Scene descriptions use the same POD structs as the
 * product's C ABI (include/misaki_b200.h) so one description feeds both sides. */
#ifndef MSK_ORACLE_H
#define MSK_ORACLE_H
#include "../include/misaki_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct OrcScene OrcScene;
typedef struct OrcRgb2Spec OrcRgb2Spec;
typedef struct { uint64_t paths, rays_closest, rays_shadow; double seconds; int32_t threads; int32_t pad_; } OrcStats;

const char *orc_last_error(void);
int  orc_scene_create(const MskSceneDesc *d, OrcScene **out);
void orc_scene_destroy(OrcScene *s);
int  orc_intersect(OrcScene *s, const MskRay *rays, MskHit *hits, size_t n, int brute_force);
int  orc_occluded(OrcScene *s, const MskRay *rays, uint8_t *occ, size_t n);
int  orc_intersect_margin(OrcScene *s, const MskRay *rays, float *second_t, float *min_bary, size_t n);
int  orc_camera_rays(OrcScene *s, const float *samples, MskRay *rays, size_t n);
/* the same camera as a C callback (ws, px, py, out16) for oracle/ref_render_wrap.cpp's Sensor stand-in */
int  orc_ref_camera_bind(OrcScene *s);
void *orc_ref_camera_callback(void);
int  orc_render(OrcScene *s, const MskRenderDesc *rd, float *film, int nthreads, OrcStats *stats);
int  orc_aov_channels(const int32_t *types, uint32_t ntypes);
int  orc_render_aov(OrcScene *s, const MskRenderDesc *rd, const int32_t *types, uint32_t ntypes, float *film, int nthreads, OrcStats *stats);
int  orc_trace_samples(OrcScene *s, const MskRenderDesc *rd, const uint32_t *pixel_sample, float *out, size_t n);
void orc_develop(const float *film, float *rgba, size_t npixels);
void orc_gaussian_filter(float stddev, float *radius, float table[33]);
void orc_pcg32_floats(uint64_t seed, uint64_t base_seed, float *out, size_t n);
void orc_pcg32_uints(uint64_t initstate, uint64_t initseq, uint32_t *out, size_t n);
void orc_sample_wavelength(float u, float wl[4], float w[4]);
void orc_warp(int which, float u, float v, float out[3]);
void orc_fresnel(float cos_theta_i, float eta, float out[4]);
void orc_fresnel_conductor(float cos_theta_i, const float eta[4], const float k[4], float out[4]);
void orc_srgb_model_eval(const float c[3], const float wl[4], float out[4]);
int  orc_sample_ray(OrcScene *s, const MskRenderDesc *rd, uint64_t seed, const float o[3], const float d[3], float mint, float maxt,
                    const float wl[4], int bsdf_draws_right_to_left, float out[4]);
/* MskRenderDesc::flags bit understood by the oracle only (test infrastructure): reproduce the reference's sampler seeding
 * (one sampler for the whole image, spiral block order) and GCC's argument evaluation order, see oracle.cpp render_impl */
#define ORC_RENDER_REFERENCE_SEEDING 0x80000000u
int  orc_film_accumulate(float stddev, int W, int H, int nch, int block_size, const float *samples, size_t n, float *film_out, int *block_order);
int  orc_aov_sample_ray(OrcScene *s, const MskRenderDesc *rd, uint64_t seed, const float o[3], const float d[3], float mint, float maxt,
                        const float wl[4], int bsdf_draws_right_to_left, const int32_t *types, uint32_t ntypes, float *out_aovs, float out[4]);
void orc_mesh_interaction(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, int normals, int uvs, uint32_t prim,
                          float u, float v, float t, const float o[3], const float d[3], float out[27]);
void orc_mesh_sampling(const float *verts, uint32_t nverts, const uint32_t *tris, uint32_t ntris, int normals, int uvs, const float sample[2],
                       const float ref_p[3], float out[22], float *cdf_out);
void orc_math(int which, const float *in, float *out); /* coordinate_system, Frame, reflect/refract, ggx pdf/G, xyz_to_srgb */
void orc_distribution_sample_reuse(const float *pdf, size_t n, const float *u, size_t nu, uint32_t *index, float *reused, float *cdf_out);
int  orc_spectrum_eval(OrcScene *s, int id, const float wl[4], float out[4]);
int  orc_texture_eval(OrcScene *s, int id, float u, float v, const float wl[4], float out[4]); /* Texture::eval(si) at si.uv */
void orc_spectrum_to_xyz(const float value[4], const float wl[4], float xyz[3]);
void orc_ggx(int which, float au, float av, const float a[3], const float b[3], float out[4]);
int  orc_bsdf(OrcScene *s, int bsdf_id, const float wi[3], const float wl[4], const float smp[3], const float wo_in[3],
              float out_sample[10], float out_eval[4], float *out_pdf);
OrcRgb2Spec *orc_rgb2spec_load(const char *filename);
void orc_rgb2spec_free(OrcRgb2Spec *m);
void orc_rgb2spec_fetch(const OrcRgb2Spec *model, const float rgb[3], float out[3]);
#ifdef __cplusplus
}
#endif
#endif
