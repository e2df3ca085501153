/*
 * misaki_b200.h -- C ABI of the B200-native backend for misaki-render's
 * path-tracing hot path.
 *
 * The reference (jczh98/misaki-render) has no FFI: its "plugins" are C++ classes
 * registered in-process (include/misaki/core/manager.h:39-45) and its only
 * accelerator seam is Scene::accel_init / ray_intersect / ray_test behind
 * `void *m_accel` (include/misaki/render/scene.h:17-20,58), implemented with
 * Embree (src/librender/scene.cpp:197-275).  This header is the boundary a
 * maintainer binds instead of Embree + the TBB tile loop:
 *
 *   entry point              replaces (reference file:line)
 *   ------------------------ ----------------------------------------------------
 *   msk_gpu_scene_create     Scene::accel_init           scene.cpp:201-212
 *                            Mesh::embree_geometry       mesh.cpp:141-151
 *   msk_gpu_scene_destroy    Scene::accel_release        scene.cpp:214
 *   msk_gpu_intersect[_dev]  Scene::ray_intersect        scene.cpp:216-253 (rtcIntersect1)
 *   msk_gpu_occluded[_dev]   Scene::ray_test             scene.cpp:255-273 (rtcOccluded1)
 *   msk_gpu_render[_dev]     SamplingIntegrator::render  integrator.cpp:31-126
 *                            PathTracer::sample          integrators/path.cpp:23-131
 *                            ImageBlock::put/Film::put   imageblock.cpp:36-114, hdrfilm.cpp:43-46
 *   msk_gpu_develop[_dev]    HDRFilm::image              hdrfilm.cpp:48-90
 *   msk_gpu_render_aov[_dev] AOVIntegrator::sample       integrators/aov.cpp:87-144 (+ the nested "path" child)
 *
 * Conventions: plain C, POD structs, no exceptions across the boundary.  Every
 * function returns 0 on success or a negative MskStatus; msk_gpu_last_error()
 * returns a thread-local message.  Host input arrays are copied during
 * msk_gpu_scene_create and may be freed afterwards.  One MskCtx per GPU;
 * calls on one ctx must be serialised by the caller.
 *
 * All floating-point data are IEEE float32; indices are uint32.
 */
#ifndef MISAKI_B200_H
#define MISAKI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSK_ABI_VERSION 5

typedef enum {
    MSK_OK            = 0,
    MSK_ERR_ARG       = -1, /* invalid argument / malformed scene description */
    MSK_ERR_CUDA      = -2, /* a CUDA runtime call failed (message has details) */
    MSK_ERR_NO_DEVICE = -3, /* no usable sm_100 device */
    MSK_ERR_OOM       = -4,
    MSK_ERR_UNSUPPORTED = -5
} MskStatus;

/* ---- spectra (include/misaki/render/texture.h, src/librender/spectra/ *.cpp) ---- */
typedef enum {
    MSK_SPEC_UNIFORM  = 0, /* spectra/uniform.cpp:19-26      value                      */
    MSK_SPEC_SRGB     = 1, /* spectra/srgb.cpp:14-23         rgb2spec coeffs c[3]       */
    MSK_SPEC_SRGB_D65 = 2, /* spectra/srgb_d65.cpp:14-36     coeffs c[3] x table        */
    MSK_SPEC_REGULAR  = 3, /* spectra/regular.cpp:73-91      table (also "d65")         */
    MSK_SPEC_SRGB_UNBOUNDED = 4, /* builder decision for conductor eta/k: value * srgb_model_eval(c) */
    MSK_SPEC_CHECKERBOARD = 5 /* textures/checkerboard.cpp:11-31: selects child0 / child1 by the surface uv */
} MskSpectrumKind;

typedef struct {
    int32_t  kind;         /* MskSpectrumKind */
    float    c[3];         /* rgb2spec polynomial coefficients (srgb_model_eval, srgb.h:8-19) */
    float    value;        /* UNIFORM: the value; SRGB_UNBOUNDED: the scale */
    uint32_t table_offset; /* REGULAR / SRGB_D65: first entry in MskSceneDesc::spectrum_tables */
    uint32_t table_size;   /*   number of entries (>= 2) */
    float    lambda_min;   /*   wavelength of entry 0 */
    float    lambda_max;   /*   wavelength of the last entry */
    int32_t  child0, child1; /* CHECKERBOARD: spectrum ids of "color0" / "color1" (both < this spectrum's own id) */
    float    to_uv[6];     /* CHECKERBOARD: rows 0,1 x columns 0..2 of the "to_uv" 4x4 (Transform4f::extract,
                            * transform.h:142-148): uv' = (m00 u + m01 v + m02, m10 u + m11 v + m12) */
} MskSpectrum;

/* ---- BSDFs (include/misaki/render/bsdf.h:82-126, src/librender/bsdfs/ *.cpp) ---- */
typedef enum {
    MSK_BSDF_DIFFUSE         = 0, /* bsdfs/diffuse.cpp         */
    MSK_BSDF_CONDUCTOR       = 1, /* bsdfs/conductor.cpp       */
    MSK_BSDF_ROUGHCONDUCTOR  = 2, /* bsdfs/roughconductor.cpp  */
    MSK_BSDF_ROUGHDIELECTRIC = 3, /* bsdfs/roughdielectric.cpp */
    MSK_BSDF_DIELECTRIC      = 4, /* bsdfs/dielectric.cpp      */
    MSK_BSDF_TYPE_COUNT      = 5
} MskBsdfType;

typedef struct {
    int32_t type;            /* MskBsdfType */
    int32_t reflectance;     /* spectrum id: diffuse "reflectance" / "specular_reflectance" */
    int32_t transmittance;   /* spectrum id: "specular_transmittance" (dielectrics), else -1 */
    int32_t eta;             /* spectrum id: conductor "eta", else -1 */
    int32_t k;               /* spectrum id: conductor "k",   else -1 */
    float   alpha_u, alpha_v;/* microfacet roughness ("alpha" sets both) */
    float   int_ior, ext_ior;/* dielectrics */
    int32_t distribution;    /* 0 = beckmann (a stub in the reference -> rejected), 1 = ggx */
    int32_t sample_visible;  /* microfacet "sample_visible" (reference samples D, not VNDF) */
    int32_t twosided;        /* 1: wrapped in bsdfs/twosided.cpp with the same BRDF on both sides */
} MskBsdf;

/* ---- emitters (src/librender/emitters/{area,constant}.cpp) ---- */
typedef enum { MSK_EMITTER_AREA = 0, MSK_EMITTER_CONSTANT = 1 } MskEmitterType;

typedef struct {
    int32_t type;     /* MskEmitterType */
    int32_t radiance; /* spectrum id */
    int32_t shape;    /* AREA: index into meshes; CONSTANT: -1 */
} MskEmitter;

/* ---- participating media (src/librender/medium.cpp, media/homogeneous.cpp, phase/isotropic.cpp) ---- */
typedef enum { MSK_PHASE_ISOTROPIC = 0 } MskPhaseType; /* phase/isotropic.cpp: the only phase function of the reference */

typedef struct {
    int32_t sigma_a;  /* spectrum id, homogeneous.cpp:15 "sigma_a" (absorption coefficient per scene unit) */
    int32_t sigma_s;  /* spectrum id, homogeneous.cpp:16 "sigma_s" (scattering coefficient) */
    int32_t phase;    /* MskPhaseType */
    float   scale;    /* homogeneous.cpp:18 "scale": read and stored by the reference, applied nowhere -- kept so */
} MskMedium;

/* ---- shapes (src/librender/mesh.cpp, shapes/obj.cpp:137-177) ---- */
typedef struct {
    const float    *verts;   /* nverts x 8 floats [px py pz nx ny nz u v], stride 32 B, world space */
    const uint32_t *tris;    /* ntris x 3 vertex indices */
    uint32_t nverts, ntris;
    int32_t  bsdf;           /* index into bsdfs */
    int32_t  emitter;        /* index into emitters, or -1 */
    uint8_t  has_normals;    /* Mesh::has_vertex_normals   */
    uint8_t  has_uvs;        /* Mesh::has_vertex_texcoords */
    uint8_t  pad_[2];
    int32_t  interior_medium; /* shape.cpp:28-39: index into media of the child medium named "interior", or -1 */
    int32_t  exterior_medium; /*   ... of any other child medium, or -1 (at most 254 media per scene) */
} MskMesh;

/* ---- sensor + film (sensors/perspective.cpp:9-41, film.cpp, filters/gaussian.cpp) ---- */
typedef struct {
    float    sample_to_camera[16]; /* row-major; pixel-unit sample -> camera space */
    float    to_world[16];         /* row-major camera-to-world */
    float    near_clip, far_clip;
    uint32_t width, height;
    float    filter_radius;        /* ReconstructionFilter::m_radius */
    float    filter_table[33];     /* normalised m_values, rfilter.cpp:12-27 */
} MskCamera;

typedef struct {
    const MskMesh     *meshes;   uint32_t nmeshes;   /* in Scene::m_shapes order == geomID */
    const MskBsdf     *bsdfs;    uint32_t nbsdfs;
    const MskEmitter  *emitters; uint32_t nemitters; /* in Scene::m_emitters order */
    const MskSpectrum *spectra;  uint32_t nspectra;
    const float       *spectrum_tables; uint32_t ntable_floats;
    int32_t            environment;                  /* emitter index of the environment, or -1 */
    MskCamera          camera;
    const MskMedium   *media;    uint32_t nmedia;    /* consumed by the "volpath" integrator only */
    int32_t            sensor_medium;                /* sensor.cpp:12-18: medium the camera sits in, or -1 */
} MskSceneDesc;

typedef struct {
    uint32_t spp;           /* Sampler::sample_count: samples per pixel of the WHOLE job (seeding) */
    uint32_t sample_begin;  /* this call renders samples [sample_begin, sample_end) of every pixel */
    uint32_t sample_end;
    int32_t  max_depth;     /* -1 = unbounded (integrator.cpp:134-136) */
    int32_t  rr_depth;      /* integrator.cpp:130 */
    int32_t  hide_emitters; /* integrator.cpp:23 */
    uint64_t base_seed;     /* Sampler "base_seed" (sampler.cpp:9) */
    uint32_t clear_film;    /* 1: zero the film first; 0: accumulate on top */
    uint32_t paths_per_batch; /* 0 = default pool size */
    uint32_t flags;         /* MSK_RENDER_* */
    uint32_t integrator;    /* MskIntegrator: which MonteCarloIntegrator::sample the wavefront executes */
} MskRenderDesc;

typedef enum {
    MSK_INTEGRATOR_PATH    = 0, /* integrators/path.cpp:23-131    (registered as "path")    */
    MSK_INTEGRATOR_VOLPATH = 1  /* integrators/volpath.cpp:26-167 (registered as "volpath") */
} MskIntegrator;

#define MSK_RENDER_STAGE_TIMERS 1u /* fill MskStats::ms_{raygen,intersect,shade,shadow,film} (adds event records) */
#define MSK_RENDER_TRAVERSAL_STATS 2u /* run the instrumented traversal kernels and fill MskStats::nodes_*, tris_* */

typedef struct {
    uint64_t paths;          /* camera samples completed */
    uint64_t rays_closest;   /* ray_intersect-equivalent queries */
    uint64_t rays_shadow;    /* ray_test-equivalent queries */
    uint64_t kernel_launches;
    float    ms_render;      /* device time, first stage launch -> film final */
    float    ms_intersect;   /* device time spent in closest-hit launches */
    float    ms_shadow, ms_shade, ms_raygen, ms_film;
    uint32_t bounces;        /* wavefront iterations executed (max over batches) */
    uint32_t batches;
    uint32_t n_intersect_launches, n_shade_launches, n_shadow_launches; /* with MSK_RENDER_STAGE_TIMERS */
    uint32_t pad_;
    uint64_t shaded_vertices; /* path vertices processed by the shade stage (hits + misses) */
    uint64_t nodes_closest, tris_closest; /* with MSK_RENDER_TRAVERSAL_STATS: wide nodes visited / triangles */
    uint64_t nodes_shadow, tris_shadow;   /*   tested, summed over all closest-hit / any-hit queries        */
    float    ms_sort;        /* device time of the material sort launches (with MSK_RENDER_STAGE_TIMERS) */
    uint32_t pad2_;
    /* unbounded-depth jobs finish with one per-path launch once the queue is short (k_tail); its rays are included in
     * rays_closest / rays_shadow and listed here so that per-kernel rooflines can separate them */
    uint64_t tail_rays_closest, tail_rays_shadow;
    float    ms_tail;        /* with MSK_RENDER_STAGE_TIMERS */
    uint32_t n_tail_launches;
} MskStats;

/* ---- AOV integrator (src/librender/integrators/aov.cpp:22-29,87-144) ---- */
typedef enum {
    MSK_AOV_DEPTH           = 0, /* 1 channel : si.t, 0 on a miss                (aov.cpp:97-99)   */
    MSK_AOV_POSITION        = 1, /* 3 channels: si.p                             (aov.cpp:101-105) */
    MSK_AOV_UV              = 2, /* 2 channels: si.uv                            (aov.cpp:107-110) */
    MSK_AOV_GEO_NORMAL      = 3, /* 3 channels: si.n                             (aov.cpp:112-116) */
    MSK_AOV_SH_NORMAL       = 4, /* 3 channels: si.sh_frame.n                    (aov.cpp:118-122) */
    MSK_AOV_INTEGRATOR_RGBA = 5  /* 4 channels: nested "path" integrator, linear sRGB + 1 (aov.cpp:124-140) */
} MskAovType;
#define MSK_AOV_MAX_CHANNELS 32

typedef struct {
    const int32_t *types;   /* MskAovType, in the order of the "aovs" string followed by the nested integrators */
    uint32_t       ntypes;
    uint32_t       pad_;
} MskAovDesc;

typedef struct { float o[3]; float tmin; float d[3]; float tmax; } MskRay;     /* 32 B */
typedef struct { float t, u, v; uint32_t prim; uint32_t geom; } MskHit;          /* 20 B, t=+inf: miss */

typedef struct {
    uint64_t ntris, nnodes;       /* wide nodes in the traversal structure */
    uint64_t node_bytes, tri_bytes;
    float    ms_build;            /* device time of the BVH build */
    float    sah_cost;
    uint32_t max_depth;
    uint32_t pad_;
} MskAccelInfo;

typedef struct MskCtx   MskCtx;
typedef struct MskScene MskScene;

int         msk_gpu_abi_version(void);
const char *msk_gpu_last_error(void);

/* number of CUDA devices visible to the process (0 without a driver); the host plugin's `devices` property and
 * MSK_DEVICES are resolved against it */
int  msk_gpu_device_count(void);
int  msk_gpu_init(int device, MskCtx **out);
void msk_gpu_shutdown(MskCtx *ctx);
/* the CUDA stream all work of this ctx is enqueued on (a cudaStream_t) */
void *msk_gpu_stream(MskCtx *ctx);

int  msk_gpu_scene_create(MskCtx *ctx, const MskSceneDesc *desc, MskScene **out);
void msk_gpu_scene_destroy(MskScene *scene);
int  msk_gpu_accel_info(MskScene *scene, MskAccelInfo *out);

/* host buffers in, host buffers out (copies inside) */
int  msk_gpu_intersect(MskScene *scene, const MskRay *rays, MskHit *hits, size_t n);
int  msk_gpu_occluded(MskScene *scene, const MskRay *rays, uint8_t *occluded, size_t n);
/* device buffers; asynchronous on msk_gpu_stream(ctx) */
int  msk_gpu_intersect_dev(MskScene *scene, const MskRay *d_rays, MskHit *d_hits, size_t n);
int  msk_gpu_occluded_dev(MskScene *scene, const MskRay *d_rays, uint8_t *d_occluded, size_t n);
/* per-ray traversal statistics (instrumented variant of the closest-hit kernel; host buffers) */
int  msk_gpu_intersect_stats(MskScene *scene, const MskRay *rays, size_t n,
                             uint32_t *nodes_visited, uint32_t *tris_tested);

/* film layout: height x width x 5 float32, channels X,Y,Z,A,W (integrator.cpp:39-40) */
int  msk_gpu_render(MskScene *scene, const MskRenderDesc *rd, float *film_host, MskStats *stats);
/* Allocates now the path pools a render of `rd` on this scene will use (one per batch in flight), so that the render
 * itself allocates nothing; optional -- msk_gpu_render* allocate on first use and keep the pools.  The reference has no
 * counterpart (its per-thread state is the sampler clone and the ImageBlock of integrator.cpp:57-61). */
int  msk_gpu_render_reserve(MskScene *scene, const MskRenderDesc *rd);
int  msk_gpu_render_dev(MskScene *scene, const MskRenderDesc *rd, float *d_film, MskStats *stats);
/* AOV integrator: film layout height x width x (5 + msk_gpu_aov_channels(aov)) float32 = X,Y,Z,A,W followed by
 * the AOV channels, every channel splatted with the reconstruction filter like XYZAW (integrator.cpp:103-126).
 * At most one MSK_AOV_INTEGRATOR_RGBA entry (the nested path tracer, configured by `rd`); without it X,Y,Z are 0
 * (the reference returns an uninitialised Spectrum there, aov.cpp:91,142) and fields of a missed ray read 0
 * (uninitialised in the reference, scene.cpp:247-251). */
int  msk_gpu_aov_channels(const MskAovDesc *aov);
int  msk_gpu_render_aov(MskScene *scene, const MskRenderDesc *rd, const MskAovDesc *aov, float *film_host, MskStats *stats);
int  msk_gpu_render_aov_dev(MskScene *scene, const MskRenderDesc *rd, const MskAovDesc *aov, float *d_film, MskStats *stats);
/* XYZAW film -> RGBA (linear sRGB / W, A / W), both host, n = width*height pixels */
int  msk_gpu_develop(MskScene *scene, const float *film_host, float *rgba_host);
/* the same on device buffers (XYZAW film as written by msk_gpu_render_dev; rgba: 4 floats per pixel, 16-byte aligned);
 * asynchronous on msk_gpu_stream(ctx); bit-identical to msk_gpu_develop */
int  msk_gpu_develop_dev(MskScene *scene, const float *d_film, float *d_rgba);

/* ---- multi-GPU film reduction over NVLink peer memory (replaces Film::put under the mutex,
 * films/hdrfilm.cpp:43-46, across GPUs).  One process per GPU.  Every rank creates a share (a cudaMalloc'ed film with
 * a control word in front), exports its CUDA IPC handle, receives the other ranks' handles out of band (bench.py and
 * misaki_render_b200/distributed.py use torch.distributed's object all-gather) and opens them.  The root passes the
 * handles of ALL other ranks in rank order, the other ranks need not open anything.  msk_gpu_reduce_film is
 * asynchronous on msk_gpu_stream(ctx): the root adds the peers' films of this epoch to its own in ONE kernel that
 * waits for the peers on the device and pulls their films through NVLink; a non-root rank publishes its film and
 * holds its stream until the root has read it.  `epoch`: non-zero, the same on every rank, different from the previous
 * reduction's.  Device-side waits are bounded by MSK_PEER_TIMEOUT_S (default 30 s); msk_gpu_film_share_check
 * synchronises and reports a time-out. */
typedef struct { unsigned char reserved[64]; } MskIpcMemHandle;
typedef struct MskFilmShare MskFilmShare;
int    msk_gpu_film_share_create(MskCtx *ctx, size_t nfloats, MskFilmShare **out);
float *msk_gpu_film_share_ptr(MskFilmShare *share); /* device pointer of the local film (render into it with *_dev) */
int    msk_gpu_film_share_export(MskFilmShare *share, MskIpcMemHandle *out);
int    msk_gpu_film_share_open(MskFilmShare *share, const MskIpcMemHandle *peers, uint32_t npeers);
/* one process driving several GPUs: the root maps its peers' shares (same process, one MskCtx per device) with
 * cudaDeviceEnablePeerAccess instead of IPC handles; peers in device order, as for msk_gpu_film_share_open */
int    msk_gpu_film_share_attach(MskFilmShare *share, MskFilmShare *const *peers, uint32_t npeers);
int    msk_gpu_reduce_film(MskFilmShare *share, int is_root, uint32_t epoch);
int    msk_gpu_film_share_check(MskFilmShare *share);
void   msk_gpu_film_share_destroy(MskFilmShare *share);

/* ---- one call, several GPUs of one process: what SamplingIntegrator::render does with every core of the machine
 * (integrator.cpp:54-75) done with every listed GPU.  scenes[i] was created from the SAME description on its own
 * context (one MskCtx per device); GPU i renders the i-th balanced sub-range of rd's [sample_begin, sample_end) of
 * every pixel on its own host thread, scenes[0]'s GPU sums the films in list order with the peer-memory kernel above
 * and the result is copied to film_host (XYZAW, as msk_gpu_render).  Identical to the one-GPU film up to float
 * summation order (per-(pixel, sample) seeding).  stats: counters summed, ms_render = the slowest GPU. */
int    msk_gpu_render_multi(MskScene *const *scenes, uint32_t nscenes, const MskRenderDesc *rd, float *film_host, MskStats *stats);

#ifdef __cplusplus
}
#endif
#endif /* MISAKI_B200_H */
