/* C API of the host front-end library (libmisaki_host.so): load a misaki XML scene, look at the flattened
 * device description, render it through the "path" integrator plugin (-> include/misaki_b200.h). */
#ifndef MSK_HOST_CAPI_H
#define MSK_HOST_CAPI_H
#include "../../include/misaki_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct MskhScene MskhScene;
const char *mskh_last_error(void);
void mskh_set_log_level(int level); /* 0 trace .. 4 error */
void mskh_add_search_path(const char *dir); /* FileResolver::append (data/srgb.coeff, meshes) */
int  mskh_load_file(const char *filename, MskhScene **out);
/* with $name substitutions (xml::load_file's ParameterList, reference xml.cpp:352-362) */
int  mskh_load_file_params(const char *filename, const char *const *names, const char *const *values, size_t nparams, MskhScene **out);
int  mskh_load_string(const char *xml_text, const char *base_dir, MskhScene **out);
void mskh_free(MskhScene *scene);
/* the POD description the integrator hands to msk_gpu_scene_create; owned by the handle */
const MskSceneDesc *mskh_scene_desc(MskhScene *scene);
int  mskh_render_desc(MskhScene *scene, MskRenderDesc *out);
/* Integrator::render + Film::develop (main.cpp:37-41); output_filename may be NULL (no file written) */
int  mskh_render(MskhScene *scene, const char *output_filename, MskStats *stats);
int  mskh_registered_plugins(char *buffer, size_t size);
/* HDRFilm::image (XYZAW -> RGBA) and the image writers behind Film::develop */
int  mskh_develop(const float *film_xyzaw, size_t npixels, float *rgba);
/* the same with AOV channels after W (hdrfilm.cpp:84-87): film npixels x nchannels -> out npixels x (nchannels - 1) */
int  mskh_develop_channels(const float *film, size_t npixels, size_t nchannels, float *out);
int  mskh_write_exr(const char *filename, const float *rgba, uint32_t width, uint32_t height);
int  mskh_write_pfm(const char *filename, const float *rgba, uint32_t width, uint32_t height);
/* srgb_model_fetch (rgb2spec) as the spectrum plugins use it */
int  mskh_srgb_model_fetch(const float rgb[3], float out[3]);
#ifdef __cplusplus
}
#endif
#endif
