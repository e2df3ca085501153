// C ABI of the backend (include/misaki_b200.h): context, scene upload + BVH build,
// batch ray queries, render.  No torch types, no exceptions across the boundary.
#include "msk_bvh.h"
#include "msk_device.cuh"
#include "msk_render.h"
#include "spectral_tables.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace msk {

static thread_local std::string g_error;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

int cuda_fail(cudaError_t err, const char *expr, const char *file, int line) {
    const char *base = strrchr(file, '/');
    return fail(err == cudaErrorMemoryAllocation ? MSK_ERR_OOM : MSK_ERR_CUDA, "CUDA error %s (%s) at %s:%d: %s",
                cudaGetErrorName(err), cudaGetErrorString(err), base ? base + 1 : file, line, expr);
}

} // namespace msk

using namespace msk;

struct MskCtx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    Renderer renderer;
    float *film_cache = nullptr; // device staging film of msk_gpu_render (host-buffer entry point)
    size_t film_cache_bytes = 0;
};

struct MskScene {
    MskCtx *ctx = nullptr;
    DScene d{};
    BvhResult bvh;
    std::vector<void *> allocs;
    uint64_t ntris = 0;
    ~MskScene() {
        for (void *p : allocs) cudaFree(p);
        bvh_free(&bvh);
    }
};

// accessors for msk_peer.cu
int msk_ctx_device(MskCtx *ctx) { return ctx->device; }
int msk_ctx_sm_count(MskCtx *ctx) { return ctx->sm_count; }

namespace {

template <typename T> int upload(MskScene *s, const T *host, size_t n, const T **dev) {
    T *p = nullptr;
    MSK_CUDA_CHECK(cudaMalloc((void **) &p, std::max<size_t>(n, 1) * sizeof(T)));
    s->allocs.push_back(p);
    if (n) MSK_CUDA_CHECK(cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, s->ctx->stream));
    *dev = p;
    return MSK_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { int cur = -1; cudaGetDevice(&cur); if (prev >= 0 && cur != prev) cudaSetDevice(prev); }
};

int check_spectrum(const MskSceneDesc *d, int id, const char *what) {
    if (id < 0 || (uint32_t) id >= d->nspectra) return fail(MSK_ERR_ARG, "%s: spectrum id %d out of range", what, id);
    return MSK_OK;
}

} // namespace

static void warm_up(MskCtx *ctx);

extern "C" {

int msk_gpu_abi_version(void) { return MSK_ABI_VERSION; }
const char *msk_gpu_last_error(void) { return g_error.c_str(); }

int msk_gpu_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int msk_gpu_init(int device, MskCtx **out) {
    if (!out) return fail(MSK_ERR_ARG, "msk_gpu_init: null output");
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return fail(MSK_ERR_NO_DEVICE, "no CUDA device visible (%s); this backend has no CPU fallback",
                    err == cudaSuccess ? "count = 0" : cudaGetErrorString(err));
    if (device < 0 || device >= count) return fail(MSK_ERR_ARG, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    MSK_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MSK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    MskCtx *ctx = new (std::nothrow) MskCtx;
    if (!ctx) return fail(MSK_ERR_OOM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    DeviceGuard guard(device);
    cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e2 != cudaSuccess) { delete ctx; return cuda_fail(e2, "cudaStreamCreate", __FILE__, __LINE__); }
    int rc = ctx->renderer.init(ctx->sm_count);
    if (rc) { cudaStreamDestroy(ctx->stream); delete ctx; return rc; }
    // CUDA loads a kernel's code at its first launch.  Spread over the ~40 kernels of a first scene + render that was
    // 140 ms inside the first BVH build and ~0.5 s of the first msk_gpu_scene_create + msk_gpu_render (round-1 bench).
    // Pay it here, once per context, by pushing a two-triangle scene through build + render + queries (MSK_WARMUP=0 skips).
    const char *wu = getenv("MSK_WARMUP");
    if (!(wu && *wu && !atoi(wu))) warm_up(ctx);
    *out = ctx;
    return MSK_OK;
}

void msk_gpu_shutdown(MskCtx *ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->renderer.release();
    cudaFree(ctx->film_cache);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void *msk_gpu_stream(MskCtx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }

int msk_gpu_scene_create(MskCtx *ctx, const MskSceneDesc *d, MskScene **out) {
    if (!ctx || !d || !out) return fail(MSK_ERR_ARG, "msk_gpu_scene_create: null argument");
    *out = nullptr;
    DeviceGuard guard(ctx->device);
    // MSK_DEBUG_SETUP=1: host wall-clock breakdown of this call on stderr (validation | upload | BVH build | total)
    const bool debug_setup = getenv("MSK_DEBUG_SETUP") && atoi(getenv("MSK_DEBUG_SETUP"));
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    // ---- validation (malformed descriptions must fail loudly, not read out of bounds on the device)
    if (d->camera.width == 0 || d->camera.height == 0) return fail(MSK_ERR_ARG, "film size must be positive");
    if (!(d->camera.filter_radius > 0.f) || d->camera.filter_radius > 4.f)
        return fail(MSK_ERR_UNSUPPORTED, "filter radius %g outside (0, 4]", d->camera.filter_radius);
    if ((d->nmeshes && !d->meshes) || (d->nbsdfs && !d->bsdfs) || (d->nspectra && !d->spectra) || (d->nemitters && !d->emitters) ||
        (d->ntable_floats && !d->spectrum_tables))
        return fail(MSK_ERR_ARG, "scene description: null array with a non-zero count");
    for (uint32_t i = 0; i < d->nspectra; ++i) {
        const MskSpectrum &s = d->spectra[i];
        if (s.kind < 0 || s.kind > MSK_SPEC_CHECKERBOARD) return fail(MSK_ERR_ARG, "spectrum %u: unknown kind %d", i, s.kind);
        if (s.kind == MSK_SPEC_CHECKERBOARD && (s.child0 < 0 || s.child1 < 0 || (uint32_t) s.child0 >= i || (uint32_t) s.child1 >= i))
            return fail(MSK_ERR_ARG, "spectrum %u: checkerboard children (%d, %d) must be spectra declared before it", i, s.child0, s.child1);
        if (s.kind == MSK_SPEC_REGULAR || s.kind == MSK_SPEC_SRGB_D65) {
            if (s.table_size < 2 || (uint64_t) s.table_offset + s.table_size > d->ntable_floats)
                return fail(MSK_ERR_ARG, "spectrum %u: table out of range", i);
            if (!(s.lambda_min < s.lambda_max)) return fail(MSK_ERR_ARG, "ContinuousDistribution: invalid range!");
        }
    }
    for (uint32_t i = 0; i < d->nbsdfs; ++i) {
        const MskBsdf &b = d->bsdfs[i];
        if (b.type < 0 || b.type >= MSK_BSDF_TYPE_COUNT) return fail(MSK_ERR_ARG, "bsdf %u: unknown type %d", i, b.type);
        int rc = check_spectrum(d, b.reflectance, "bsdf reflectance");
        if (rc) return rc;
        // ids a BSDF type does not use must still be -1 or valid: the texture pass resolves every id >= 0 (bsdf_resolve_textures)
        for (int32_t id : { b.transmittance, b.eta, b.k })
            if (id < -1 || id >= (int32_t) d->nspectra) return fail(MSK_ERR_ARG, "bsdf %u: spectrum id %d out of range (unused ids must be -1)", i, id);
        if (b.type == MSK_BSDF_CONDUCTOR || b.type == MSK_BSDF_ROUGHCONDUCTOR) {
            if ((rc = check_spectrum(d, b.eta, "conductor eta")) || (rc = check_spectrum(d, b.k, "conductor k"))) return rc;
        }
        if (b.type == MSK_BSDF_ROUGHDIELECTRIC || b.type == MSK_BSDF_DIELECTRIC) {
            if ((rc = check_spectrum(d, b.transmittance, "specular_transmittance"))) return rc;
            if (b.int_ior < 0.f || b.ext_ior < 0.f || b.int_ior == b.ext_ior)
                return fail(MSK_ERR_ARG, "The interior and exterior indices of refraction must be positive and differ!");
        }
        if ((b.type == MSK_BSDF_ROUGHCONDUCTOR || b.type == MSK_BSDF_ROUGHDIELECTRIC) && b.distribution != 1)
            return fail(MSK_ERR_UNSUPPORTED,
                        "bsdf %u: only the \"ggx\" distribution is defined (beckmann is a stub in the reference, microfacet.h:113-115)", i);
        if (b.twosided && (b.type == MSK_BSDF_ROUGHDIELECTRIC || b.type == MSK_BSDF_DIELECTRIC))
            return fail(MSK_ERR_ARG, "Only materials without a transmission component can be nested!");
    }
    if (d->nmedia > 254) return fail(MSK_ERR_UNSUPPORTED, "more than 254 media");
    if (d->nmedia && !d->media) return fail(MSK_ERR_ARG, "media: null buffer");
    for (uint32_t i = 0; i < d->nmedia; ++i) {
        const MskMedium &m = d->media[i];
        int rc;
        if ((rc = check_spectrum(d, m.sigma_a, "medium sigma_a")) || (rc = check_spectrum(d, m.sigma_s, "medium sigma_s"))) return rc;
        if (m.phase != MSK_PHASE_ISOTROPIC) return fail(MSK_ERR_UNSUPPORTED, "medium %u: unknown phase function %d", i, m.phase);
    }
    if (d->nmedia && (d->sensor_medium < -1 || d->sensor_medium >= (int32_t) d->nmedia)) return fail(MSK_ERR_ARG, "sensor medium out of range");
    int env_count = 0;
    for (uint32_t i = 0; i < d->nemitters; ++i) {
        const MskEmitter &e = d->emitters[i];
        int rc = check_spectrum(d, e.radiance, "emitter radiance");
        if (rc) return rc;
        if (e.type == MSK_EMITTER_AREA) {
            if (e.shape < 0 || (uint32_t) e.shape >= d->nmeshes) return fail(MSK_ERR_ARG, "emitter %u: shape out of range", i);
            if (d->meshes[e.shape].emitter != (int32_t) i) return fail(MSK_ERR_ARG, "emitter %u: shape does not point back", i);
            if (d->meshes[e.shape].ntris == 0) return fail(MSK_ERR_ARG, "emitter %u: empty mesh", i);
        } else if (e.type == MSK_EMITTER_CONSTANT) {
            env_count++;
            if (d->environment != (int32_t) i) return fail(MSK_ERR_ARG, "Can only have one environment light");
        } else
            return fail(MSK_ERR_ARG, "emitter %u: unknown type %d", i, e.type);
    }
    if (d->environment >= (int32_t) d->nemitters || (d->environment >= 0 && env_count != 1))
        return fail(MSK_ERR_ARG, "environment index out of range");

    MskScene *s = new (std::nothrow) MskScene;
    if (!s) return fail(MSK_ERR_OOM, "out of host memory");
    s->ctx = ctx;
    auto bail = [&](int rc) { cudaStreamSynchronize(ctx->stream); delete s; return rc; };

    // ---- geometry: concatenate meshes, per-mesh info, emitter area CDFs (mesh.cpp:39-48, distribution.h:84-93)
    std::vector<DMeshInfo> infos(d->nmeshes);
    size_t nverts = 0, ntris = 0;
    for (uint32_t i = 0; i < d->nmeshes; ++i) {
        const MskMesh &m = d->meshes[i];
        if (m.bsdf < 0 || (uint32_t) m.bsdf >= d->nbsdfs) return bail(fail(MSK_ERR_ARG, "mesh %u: bsdf id out of range", i));
        if (m.emitter >= (int32_t) d->nemitters) return bail(fail(MSK_ERR_ARG, "mesh %u: emitter id out of range", i));
        if (m.emitter >= 0 && (d->emitters[m.emitter].type != MSK_EMITTER_AREA || d->emitters[m.emitter].shape != (int32_t) i))
            return bail(fail(MSK_ERR_ARG, "mesh %u: emitter %d is not an area emitter of this mesh", i, m.emitter));
        if ((m.nverts && !m.verts) || (m.ntris && !m.tris)) return bail(fail(MSK_ERR_ARG, "mesh %u: null buffer", i));
        infos[i].vert_offset = (uint32_t) nverts; infos[i].tri_offset = (uint32_t) ntris; infos[i].ntris = m.ntris;
        infos[i].bsdf = m.bsdf; infos[i].emitter = m.emitter;
        infos[i].flags = (m.has_normals ? 1u : 0u) | (m.has_uvs ? 2u : 0u);
        if (d->nmedia) { // medium ids are only meaningful when the description carries media
            if (m.interior_medium < -1 || m.interior_medium >= (int32_t) d->nmedia || m.exterior_medium < -1 || m.exterior_medium >= (int32_t) d->nmedia)
                return bail(fail(MSK_ERR_ARG, "mesh %u: medium id out of range", i));
            infos[i].flags |= (uint32_t) (m.interior_medium + 1) << 8 | (uint32_t) (m.exterior_medium + 1) << 16;
        }
        infos[i].inv_area = 0.f; infos[i].cdf_offset = 0;
        nverts += m.nverts; ntris += m.ntris;
        if (nverts > 0xffffffffull || ntris > 0x7ffffff0ull) return bail(fail(MSK_ERR_UNSUPPORTED, "scene too large"));
    }
    float bmin[3] = { INFINITY, INFINITY, INFINITY }, bmax[3] = { -INFINITY, -INFINITY, -INFINITY };
    std::vector<float> cdfs;
    for (uint32_t i = 0; i < d->nmeshes; ++i) {
        const MskMesh &m = d->meshes[i];
        for (size_t t = 0; t < (size_t) m.ntris * 3; ++t)
            if (m.tris[t] >= m.nverts) return bail(fail(MSK_ERR_ARG, "mesh %u: vertex index %u out of range", i, m.tris[t]));
        for (uint32_t v = 0; v < m.nverts; ++v) // Mesh::recompute_bbox, mesh.cpp:22-26
            for (int a = 0; a < 3; ++a) {
                float x = m.verts[(size_t) v * 8 + a];
                if (!std::isfinite(x)) return bail(fail(MSK_ERR_ARG, "mesh %u: non-finite vertex position", i));
                bmin[a] = std::min(bmin[a], x); bmax[a] = std::max(bmax[a], x);
            }
        if (m.emitter >= 0) {
            infos[i].cdf_offset = (uint32_t) cdfs.size();
            float area_sum = 0.f, acc = 0.f;
            size_t base = cdfs.size();
            cdfs.push_back(0.f);
            for (uint32_t t = 0; t < m.ntris; ++t) {
                const float *p0 = m.verts + (size_t) m.tris[3 * t] * 8, *p1 = m.verts + (size_t) m.tris[3 * t + 1] * 8,
                            *p2 = m.verts + (size_t) m.tris[3 * t + 2] * 8;
                float e0[3] = { p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2] }, e1[3] = { p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2] };
                float cx = e0[1] * e1[2] - e0[2] * e1[1], cy = e0[2] * e1[0] - e0[0] * e1[2], cz = e0[0] * e1[1] - e0[1] * e1[0];
                float a = 0.5f * std::sqrt(cx * cx + cy * cy + cz * cz); // mesh.h:51-57
                area_sum += a;
                acc = t == 0 ? a : acc + a;
                cdfs.push_back(acc);
            }
            const float inv_sum = 1.f / cdfs.back();
            for (size_t k = base; k < cdfs.size(); ++k) cdfs[k] *= inv_sum;
            if (!(area_sum > 0.f)) return bail(fail(MSK_ERR_ARG, "mesh %u: area emitter with zero surface area", i));
            infos[i].inv_area = 1.f / area_sum;
        }
    }
    const double ms_validate = since(t_begin);
    // one contiguous upload per array
    float4 *d_verts = nullptr;
    uint32_t *d_indices = nullptr;
    {
        cudaError_t e1 = cudaMalloc((void **) &d_verts, std::max<size_t>(nverts, 1) * 2 * sizeof(float4));
        if (e1 != cudaSuccess) return bail(cuda_fail(e1, "cudaMalloc(verts)", __FILE__, __LINE__));
        s->allocs.push_back(d_verts);
        e1 = cudaMalloc((void **) &d_indices, std::max<size_t>(ntris, 1) * 3 * sizeof(uint32_t));
        if (e1 != cudaSuccess) return bail(cuda_fail(e1, "cudaMalloc(indices)", __FILE__, __LINE__));
        s->allocs.push_back(d_indices);
        for (uint32_t i = 0; i < d->nmeshes; ++i) {
            const MskMesh &m = d->meshes[i];
            if (m.nverts) {
                e1 = cudaMemcpyAsync(d_verts + 2 * (size_t) infos[i].vert_offset, m.verts, (size_t) m.nverts * 32, cudaMemcpyHostToDevice, ctx->stream);
                if (e1 != cudaSuccess) return bail(cuda_fail(e1, "cudaMemcpy(verts)", __FILE__, __LINE__));
            }
            if (m.ntris) {
                e1 = cudaMemcpyAsync(d_indices + 3 * (size_t) infos[i].tri_offset, m.tris, (size_t) m.ntris * 12, cudaMemcpyHostToDevice, ctx->stream);
                if (e1 != cudaSuccess) return bail(cuda_fail(e1, "cudaMemcpy(indices)", __FILE__, __LINE__));
            }
        }
    }
    s->d.verts = d_verts; s->d.indices = d_indices;
    s->ntris = ntris;

    int rc;
    if ((rc = upload(s, infos.data(), infos.size(), &s->d.meshes))) return bail(rc);
    if ((rc = upload(s, d->bsdfs, d->nbsdfs, &s->d.bsdfs))) return bail(rc);
    if ((rc = upload(s, d->emitters, d->nemitters, &s->d.emitters))) return bail(rc);
    std::vector<DSpectrum> spectra(d->nspectra);
    for (uint32_t i = 0; i < d->nspectra; ++i) {
        const MskSpectrum &m = d->spectra[i];
        DSpectrum &o = spectra[i];
        o.kind = m.kind; o.c0 = m.c[0]; o.c1 = m.c[1]; o.c2 = m.c[2]; o.value = m.value;
        o.table_offset = m.table_offset; o.table_size = m.table_size; o.lambda_min = m.lambda_min; o.inv_interval = 0.f;
        if (m.kind == MSK_SPEC_REGULAR || m.kind == MSK_SPEC_SRGB_D65) {
            // regular.cpp:38-39,58: interval size in double, its reciprocal stored in a float
            double range = double(m.lambda_max) - double(m.lambda_min), interval = range / (m.table_size - 1);
            o.inv_interval = float(1. / interval);
        }
        if (m.kind == MSK_SPEC_CHECKERBOARD) {
            o.table_offset = (uint32_t) m.child0; o.table_size = (uint32_t) m.child1;
            o.c0 = m.to_uv[0]; o.c1 = m.to_uv[1]; o.c2 = m.to_uv[2];
            o.value = m.to_uv[3]; o.lambda_min = m.to_uv[4]; o.inv_interval = m.to_uv[5];
            s->d.has_textures = 1;
        }
    }
    if ((rc = upload(s, spectra.data(), spectra.size(), &s->d.spectra))) return bail(rc);
    if ((rc = upload(s, d->media, d->nmedia, &s->d.media))) return bail(rc);
    s->d.sensor_medium = d->nmedia ? d->sensor_medium : -1;
    if ((rc = upload(s, d->spectrum_tables, d->ntable_floats, &s->d.tables))) return bail(rc);
    if ((rc = upload(s, cdfs.data(), cdfs.size(), &s->d.cdfs))) return bail(rc);
    if ((rc = upload(s, d->camera.filter_table, 33, &s->d.filter_table))) return bail(rc);
    static_assert(sizeof(msk_cie_d65_rows) == 95 * sizeof(float4), "table layout");
    if ((rc = upload(s, reinterpret_cast<const float4 *>(&msk_cie_d65_rows[0][0]), 95, &s->d.cie))) return bail(rc);
    s->d.nemitters = d->nemitters; s->d.environment = d->environment; s->d.nmeshes = d->nmeshes;
    s->d.bsdf_type_mask = 0;
    for (uint32_t i = 0; i < d->nmeshes; ++i) s->d.bsdf_type_mask |= 1u << d->bsdfs[d->meshes[i].bsdf].type;
    // constant.cpp:21-28 + bbox.h:109-112
    s->d.env_radius = 0.f;
    if (nverts) {
        float c[3] = { (bmax[0] + bmin[0]) * 0.5f, (bmax[1] + bmin[1]) * 0.5f, (bmax[2] + bmin[2]) * 0.5f };
        float dx = c[0] - bmax[0], dy = c[1] - bmax[1], dz = c[2] - bmax[2];
        float radius = std::sqrt(dx * dx + dy * dy + dz * dz);
        const float ray_eps = std::numeric_limits<float>::epsilon() / 2 * 1500;
        s->d.env_radius = std::max(ray_eps, radius * (1.f + ray_eps));
    }
    memcpy(s->d.cam.s2c, d->camera.sample_to_camera, sizeof(float) * 16);
    memcpy(s->d.cam.c2w, d->camera.to_world, sizeof(float) * 16);
    s->d.cam.near_clip = d->camera.near_clip; s->d.cam.far_clip = d->camera.far_clip;
    s->d.cam.width = d->camera.width; s->d.cam.height = d->camera.height;
    s->d.cam.filter_radius = d->camera.filter_radius;
    s->d.cam.filter_scale = 32.f / d->camera.filter_radius; // rfilter.cpp:21

    // MSK_BVH_BUILDER=ploc selects the SAH-driven clustering builder; default LBVH, which measured better on the
    // uniformly tessellated BASELINE meshes (C5 primary rays: 15.3 vs 18.5 wide nodes per ray; fuller 8-wide nodes
    // after the collapse: 1.47 M vs 1.62 M), although PLOC lowers the binary SAH cost (66.7 -> 62.4).  A PLOC tree too
    // deep for the traversal stack (degenerate input) falls back to the balanced LBVH.
    const double ms_upload = since(t_begin) - ms_validate;
    const char *bsel = getenv("MSK_BVH_BUILDER");
    int builder = (bsel && !strcmp(bsel, "ploc")) ? MSK_BVH_PLOC : MSK_BVH_LBVH;
    if ((rc = bvh_build(ctx->stream, d_verts, d_indices, infos, &s->bvh, builder))) return bail(rc);
    if (builder == MSK_BVH_PLOC && 2 * s->bvh.depth + 2 > (uint32_t) 64) {
        bvh_free(&s->bvh);
        if ((rc = bvh_build(ctx->stream, d_verts, d_indices, infos, &s->bvh, MSK_BVH_LBVH))) return bail(rc);
    }
    if (2 * s->bvh.depth + 2 > (uint32_t) 64) // node groups + postponed triangle groups, msk_traverse.cuh kStackSize
        return bail(fail(MSK_ERR_UNSUPPORTED, "BVH depth %u exceeds the traversal stack", s->bvh.depth));
    s->d.nodes = s->bvh.nodes; s->d.tris = s->bvh.tris;
    s->d.k47 = 0x47000000u;
    s->d.nnodes = (uint32_t) s->bvh.nnodes;
    for (int a = 0; a < 3; ++a) {
        const float ext = s->bvh.hi[a] - s->bvh.lo[a];
        s->d.bb_lo[a] = s->bvh.lo[a];
        s->d.bb_scale[a] = ext > 0.f ? 0.999f / ext : 0.f;
    }
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (es != cudaSuccess) return bail(cuda_fail(es, "cudaStreamSynchronize", __FILE__, __LINE__));
    if (debug_setup) bvh_print_shape(ctx->stream, s->bvh);
    if (debug_setup)
        fprintf(stderr, "[msk] scene_create: %zu tris | validate + CDFs %.2f ms | malloc + upload %.2f ms | BVH %.2f ms host (%.2f ms device) | total %.2f ms\n",
                ntris, ms_validate, ms_upload, since(t_begin) - ms_validate - ms_upload, s->bvh.ms_build, since(t_begin));
    *out = s;
    return MSK_OK;
}

void msk_gpu_scene_destroy(MskScene *scene) {
    if (!scene) return;
    DeviceGuard guard(scene->ctx->device);
    cudaStreamSynchronize(scene->ctx->stream);
    delete scene;
}

int msk_gpu_accel_info(MskScene *s, MskAccelInfo *out) {
    if (!s || !out) return fail(MSK_ERR_ARG, "null argument");
    *out = MskAccelInfo{};
    out->ntris = s->bvh.ntris; out->nnodes = s->bvh.nnodes;
    out->node_bytes = s->bvh.nnodes * 80ull; out->tri_bytes = s->bvh.tri_slots * 48ull;
    out->ms_build = s->bvh.ms_build; out->sah_cost = s->bvh.sah_cost; out->max_depth = s->bvh.depth;
    return MSK_OK;
}

int msk_gpu_intersect_dev(MskScene *s, const MskRay *d_rays, MskHit *d_hits, size_t n) {
    if (!s || (n && (!d_rays || !d_hits))) return fail(MSK_ERR_ARG, "null argument");
    DeviceGuard guard(s->ctx->device);
    return s->ctx->renderer.intersect(s->ctx->stream, s->d, d_rays, d_hits, n);
}

int msk_gpu_occluded_dev(MskScene *s, const MskRay *d_rays, uint8_t *d_occ, size_t n) {
    if (!s || (n && (!d_rays || !d_occ))) return fail(MSK_ERR_ARG, "null argument");
    DeviceGuard guard(s->ctx->device);
    return s->ctx->renderer.occluded(s->ctx->stream, s->d, d_rays, d_occ, n);
}

int msk_gpu_intersect(MskScene *s, const MskRay *rays, MskHit *hits, size_t n) {
    if (!s || (n && (!rays || !hits))) return fail(MSK_ERR_ARG, "null argument");
    if (!n) return MSK_OK;
    DeviceGuard guard(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    MskRay *d_rays = nullptr; MskHit *d_hits = nullptr;
    MSK_CUDA_CHECK(cudaMalloc((void **) &d_rays, n * sizeof(MskRay)));
    cudaError_t e = cudaMalloc((void **) &d_hits, n * sizeof(MskHit));
    if (e != cudaSuccess) { cudaFree(d_rays); return cuda_fail(e, "cudaMalloc(hits)", __FILE__, __LINE__); }
    int rc = MSK_OK;
    e = cudaMemcpyAsync(d_rays, rays, n * sizeof(MskRay), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = s->ctx->renderer.intersect(st, s->d, d_rays, d_hits, n);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(hits, d_hits, n * sizeof(MskHit), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(st);
    cudaFree(d_rays); cudaFree(d_hits);
    if (e != cudaSuccess) return cuda_fail(e, "msk_gpu_intersect", __FILE__, __LINE__);
    return rc;
}

int msk_gpu_occluded(MskScene *s, const MskRay *rays, uint8_t *occ, size_t n) {
    if (!s || (n && (!rays || !occ))) return fail(MSK_ERR_ARG, "null argument");
    if (!n) return MSK_OK;
    DeviceGuard guard(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    MskRay *d_rays = nullptr; uint8_t *d_occ = nullptr;
    MSK_CUDA_CHECK(cudaMalloc((void **) &d_rays, n * sizeof(MskRay)));
    cudaError_t e = cudaMalloc((void **) &d_occ, n);
    if (e != cudaSuccess) { cudaFree(d_rays); return cuda_fail(e, "cudaMalloc(occ)", __FILE__, __LINE__); }
    int rc = MSK_OK;
    e = cudaMemcpyAsync(d_rays, rays, n * sizeof(MskRay), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = s->ctx->renderer.occluded(st, s->d, d_rays, d_occ, n);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(occ, d_occ, n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(st);
    cudaFree(d_rays); cudaFree(d_occ);
    if (e != cudaSuccess) return cuda_fail(e, "msk_gpu_occluded", __FILE__, __LINE__);
    return rc;
}

int msk_gpu_intersect_stats(MskScene *s, const MskRay *rays, size_t n, uint32_t *nodes_visited, uint32_t *tris_tested) {
    if (!s || (n && (!rays || !nodes_visited || !tris_tested))) return fail(MSK_ERR_ARG, "null argument");
    if (!n) return MSK_OK;
    DeviceGuard guard(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    MskRay *d_rays = nullptr; uint32_t *d_cnt = nullptr;
    MSK_CUDA_CHECK(cudaMalloc((void **) &d_rays, n * sizeof(MskRay)));
    cudaError_t e = cudaMalloc((void **) &d_cnt, 2 * n * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(d_rays); return cuda_fail(e, "cudaMalloc(stats)", __FILE__, __LINE__); }
    int rc = MSK_OK;
    e = cudaMemcpyAsync(d_rays, rays, n * sizeof(MskRay), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = s->ctx->renderer.intersect_stats(st, s->d, d_rays, n, d_cnt, d_cnt + n);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(nodes_visited, d_cnt, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(tris_tested, d_cnt + n, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(st);
    cudaFree(d_rays); cudaFree(d_cnt);
    if (e != cudaSuccess) return cuda_fail(e, "msk_gpu_intersect_stats", __FILE__, __LINE__);
    return rc;
}

int msk_gpu_render_reserve(MskScene *s, const MskRenderDesc *rd) {
    if (!s || !rd) return fail(MSK_ERR_ARG, "null argument");
    DeviceGuard guard(s->ctx->device);
    return s->ctx->renderer.reserve(s->d, *rd);
}

int msk_gpu_render_dev(MskScene *s, const MskRenderDesc *rd, float *d_film, MskStats *stats) {
    if (!s || !rd || !d_film) return fail(MSK_ERR_ARG, "null argument");
    DeviceGuard guard(s->ctx->device);
    return s->ctx->renderer.render(s->ctx->stream, s->d, *rd, d_film, stats);
}

int msk_gpu_render(MskScene *s, const MskRenderDesc *rd, float *film_host, MskStats *stats) {
    if (!s || !rd || !film_host) return fail(MSK_ERR_ARG, "null argument");
    DeviceGuard guard(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    size_t bytes = (size_t) s->d.cam.width * s->d.cam.height * 5 * sizeof(float);
    MskCtx *ctx = s->ctx;
    if (ctx->film_cache_bytes < bytes) { // staging film kept across calls: no allocator traffic per render
        cudaFree(ctx->film_cache); ctx->film_cache = nullptr; ctx->film_cache_bytes = 0;
        MSK_CUDA_CHECK(cudaMalloc((void **) &ctx->film_cache, bytes));
        ctx->film_cache_bytes = bytes;
    }
    float *d_film = ctx->film_cache;
    cudaError_t e = cudaSuccess;
    if (!rd->clear_film) e = cudaMemcpyAsync(d_film, film_host, bytes, cudaMemcpyHostToDevice, st);
    int rc = MSK_OK;
    if (e == cudaSuccess) rc = ctx->renderer.render(st, s->d, *rd, d_film, stats);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(film_host, d_film, bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "msk_gpu_render", __FILE__, __LINE__);
    return rc;
}

// Several GPUs of one process (include/misaki_b200.h).  One host thread per GPU: CUDA calls of different devices do not
// serialise behind each other, and the polls of unbounded-depth jobs (Renderer::render) block only their own thread.
int msk_gpu_render_multi(MskScene *const *scenes, uint32_t nscenes, const MskRenderDesc *rd, float *film_host, MskStats *stats) {
    if (!scenes || !nscenes || !rd || !film_host) return fail(MSK_ERR_ARG, "null argument");
    if (nscenes > 16) return fail(MSK_ERR_UNSUPPORTED, "at most 16 GPUs per render");
    for (uint32_t i = 0; i < nscenes; ++i) {
        if (!scenes[i]) return fail(MSK_ERR_ARG, "scene %u is null", i);
        if (scenes[i]->d.cam.width != scenes[0]->d.cam.width || scenes[i]->d.cam.height != scenes[0]->d.cam.height || scenes[i]->ntris != scenes[0]->ntris)
            return fail(MSK_ERR_ARG, "scene %u differs from scene 0: every GPU needs the same description", i);
        for (uint32_t j = 0; j < i; ++j)
            if (scenes[j]->ctx == scenes[i]->ctx) return fail(MSK_ERR_ARG, "scenes %u and %u share a context: one MskCtx per entry", j, i);
    }
    if (nscenes == 1) return msk_gpu_render(scenes[0], rd, film_host, stats);
    if (rd->sample_end < rd->sample_begin || rd->sample_end > rd->spp) return fail(MSK_ERR_ARG, "bad sample range");
    if (!rd->clear_film) return fail(MSK_ERR_UNSUPPORTED, "msk_gpu_render_multi accumulates into a cleared film only");
    const size_t nfloats = (size_t) scenes[0]->d.cam.width * scenes[0]->d.cam.height * 5;
    const size_t padded = (nfloats + 3) & ~(size_t) 3; // the reduction adds float4s; the tail past the film stays zero
    std::vector<MskFilmShare *> shares(nscenes, nullptr);
    std::vector<int> rcs(nscenes, MSK_OK);
    std::vector<std::string> errs(nscenes);
    std::vector<MskStats> st(nscenes);
    auto destroy_all = [&]() { for (auto *s : shares) msk_gpu_film_share_destroy(s); };
    for (uint32_t i = 0; i < nscenes; ++i) {
        int rc = msk_gpu_film_share_create(scenes[i]->ctx, padded, &shares[i]);
        if (rc) { destroy_all(); return rc; }
    }
    {
        int rc = msk_gpu_film_share_attach(shares[0], shares.data() + 1, nscenes - 1);
        if (rc) { destroy_all(); return rc; }
    }
    static std::atomic<uint32_t> g_epoch{ 0 };
    uint32_t epoch = ++g_epoch;
    if (!epoch) epoch = ++g_epoch;
    const uint32_t total = rd->sample_end - rd->sample_begin, base = total / nscenes, rem = total % nscenes;
    auto shard = [&](uint32_t i) {
        MskRenderDesc r = *rd;
        r.sample_begin = rd->sample_begin + i * base + std::min(i, rem);
        r.sample_end = r.sample_begin + base + (i < rem ? 1u : 0u);
        r.clear_film = 1;
        return r;
    };
    for (uint32_t i = 0; i < nscenes; ++i) { // every allocation happens before any GPU starts (see Renderer::reserve)
        DeviceGuard guard(scenes[i]->ctx->device);
        const MskRenderDesc r = shard(i);
        int rc = scenes[i]->ctx->renderer.reserve(scenes[i]->d, r);
        if (rc) { destroy_all(); return rc; }
    }
    auto worker = [&](uint32_t i) {
        const MskRenderDesc r = shard(i);
        int rc = msk_gpu_render_dev(scenes[i], &r, msk_gpu_film_share_ptr(shares[i]), &st[i]);
        if (!rc) rc = msk_gpu_reduce_film(shares[i], i == 0, epoch);
        if (!rc) rc = msk_gpu_film_share_check(shares[i]); // drains this GPU's stream
        if (!rc && i == 0) {
            DeviceGuard guard(scenes[0]->ctx->device);
            cudaError_t e = cudaMemcpyAsync(film_host, msk_gpu_film_share_ptr(shares[0]), nfloats * sizeof(float), cudaMemcpyDeviceToHost, scenes[0]->ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(scenes[0]->ctx->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "msk_gpu_render_multi (film copy)", __FILE__, __LINE__);
        }
        rcs[i] = rc;
        if (rc) errs[i] = g_error; // the error string is thread-local
    };
    std::vector<std::thread> threads;
    for (uint32_t i = 1; i < nscenes; ++i) threads.emplace_back(worker, i);
    worker(0);
    for (auto &t : threads) t.join();
    destroy_all();
    for (uint32_t i = 0; i < nscenes; ++i)
        if (rcs[i]) return fail(rcs[i], "GPU %d: %s", scenes[i]->ctx->device, errs[i].c_str());
    if (stats) {
        *stats = st[0];
        for (uint32_t i = 1; i < nscenes; ++i) {
            stats->paths += st[i].paths; stats->rays_closest += st[i].rays_closest; stats->rays_shadow += st[i].rays_shadow;
            stats->shaded_vertices += st[i].shaded_vertices; stats->kernel_launches += st[i].kernel_launches;
            stats->batches += st[i].batches; stats->bounces = std::max(stats->bounces, st[i].bounces);
            stats->tail_rays_closest += st[i].tail_rays_closest; stats->tail_rays_shadow += st[i].tail_rays_shadow;
            stats->ms_render = std::max(stats->ms_render, st[i].ms_render);
        }
    }
    return MSK_OK;
}

// AOVIntegrator (integrators/aov.cpp): channel count of a description, or a negative status
int msk_gpu_aov_channels(const MskAovDesc *aov) {
    if (!aov) return fail(MSK_ERR_ARG, "null argument");
    uint32_t nch = 0;
    int rc = Renderer::aov_plan(*aov, &nch);
    return rc ? rc : (int) nch;
}

int msk_gpu_render_aov_dev(MskScene *s, const MskRenderDesc *rd, const MskAovDesc *aov, float *d_film, MskStats *stats) {
    if (!s || !rd || !aov || !d_film) return fail(MSK_ERR_ARG, "null argument");
    DeviceGuard guard(s->ctx->device);
    return s->ctx->renderer.render(s->ctx->stream, s->d, *rd, d_film, stats, aov);
}

int msk_gpu_render_aov(MskScene *s, const MskRenderDesc *rd, const MskAovDesc *aov, float *film_host, MskStats *stats) {
    if (!s || !rd || !aov || !film_host) return fail(MSK_ERR_ARG, "null argument");
    uint32_t nch = 0;
    int rc = Renderer::aov_plan(*aov, &nch);
    if (rc) return rc;
    DeviceGuard guard(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    size_t bytes = (size_t) s->d.cam.width * s->d.cam.height * (5 + nch) * sizeof(float);
    MskCtx *ctx = s->ctx;
    if (ctx->film_cache_bytes < bytes) {
        cudaFree(ctx->film_cache); ctx->film_cache = nullptr; ctx->film_cache_bytes = 0;
        MSK_CUDA_CHECK(cudaMalloc((void **) &ctx->film_cache, bytes));
        ctx->film_cache_bytes = bytes;
    }
    float *d_film = ctx->film_cache;
    cudaError_t e = cudaSuccess;
    if (!rd->clear_film) e = cudaMemcpyAsync(d_film, film_host, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = ctx->renderer.render(st, s->d, *rd, d_film, stats, aov);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(film_host, d_film, bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "msk_gpu_render_aov", __FILE__, __LINE__);
    return rc;
}

} // extern "C"

// A floor quad under a quad light, diffuse, 8 x 4 pixels: touches the BVH builder, every wavefront stage (two bounces, one
// of them through the specialised shade kernels), the tail kernel, the film gather and the batch queries.  Failures are
// ignored: the real call that follows reports them.
static void warm_up(MskCtx *ctx) {
    static const float verts[2][4 * 8] = {
        { -1, 0, -1, 0, 1, 0, 0, 0,   1, 0, -1, 0, 1, 0, 1, 0,   1, 0, 1, 0, 1, 0, 1, 1,   -1, 0, 1, 0, 1, 0, 0, 1 },
        { -0.5f, 2, -0.5f, 0, -1, 0, 0, 0,   0.5f, 2, -0.5f, 0, -1, 0, 1, 0,   0.5f, 2, 0.5f, 0, -1, 0, 1, 1,   -0.5f, 2, 0.5f, 0, -1, 0, 0, 1 } };
    static const uint32_t tris_up[6] = { 0, 2, 1, 0, 3, 2 }, tris_down[6] = { 0, 1, 2, 0, 2, 3 };
    MskSpectrum spec[2]{};
    spec[0].kind = MSK_SPEC_UNIFORM; spec[0].value = 0.5f; spec[0].child0 = spec[0].child1 = -1;
    spec[1] = spec[0]; spec[1].value = 5.f;
    MskBsdf bsdf{};
    bsdf.type = MSK_BSDF_DIFFUSE; bsdf.reflectance = 0; bsdf.transmittance = bsdf.eta = bsdf.k = -1;
    MskEmitter em{};
    em.type = MSK_EMITTER_AREA; em.radiance = 1; em.shape = 1;
    MskMesh meshes[2]{};
    meshes[0].verts = verts[0]; meshes[0].nverts = 4; meshes[0].tris = tris_up; meshes[0].ntris = 2; meshes[0].bsdf = 0; meshes[0].emitter = -1;
    meshes[0].interior_medium = meshes[0].exterior_medium = -1;
    meshes[1] = meshes[0]; meshes[1].verts = verts[1]; meshes[1].tris = tris_down; meshes[1].emitter = 0;
    MskSceneDesc d{};
    d.meshes = meshes; d.nmeshes = 2; d.bsdfs = &bsdf; d.nbsdfs = 1; d.emitters = &em; d.nemitters = 1; d.spectra = spec; d.nspectra = 2;
    d.environment = -1; d.sensor_medium = -1;
    MskCamera &c = d.camera;
    c.width = 8; c.height = 4; c.near_clip = 0.01f; c.far_clip = 100.f; c.filter_radius = 2.f;
    for (int i = 0; i < 33; ++i) c.filter_table[i] = i < 32 ? 1.f - i / 32.f : 0.f;
    // sample_to_camera: pixel (x, y) -> a point on the plane z = 1 in front of a camera at (0, 1, -3) looking along +z
    const float s2c[16] = { 0.25f, 0, 0, -1,   0, -0.25f, 0, 0.5f,   0, 0, 0, 1,   0, 0, 0, 1 };
    const float c2w[16] = { 1, 0, 0, 0,   0, 1, 0, 1,   0, 0, 1, -3,   0, 0, 0, 1 };
    memcpy(c.sample_to_camera, s2c, sizeof(s2c)); memcpy(c.to_world, c2w, sizeof(c2w));
    MskScene *sc = nullptr;
    if (msk_gpu_scene_create(ctx, &d, &sc) != MSK_OK) return;
    float film[8 * 4 * 5];
    for (int32_t depth : { 3, -1 }) { // bounded: wavefront only; unbounded: polls the queue and finishes with k_tail
        MskRenderDesc rd{};
        rd.spp = 2; rd.sample_end = 2; rd.max_depth = depth; rd.rr_depth = 2; rd.clear_film = 1;
        msk_gpu_render(sc, &rd, film, nullptr);
    }
    MskRay ray{ { 0, 1, 0 }, 1e-3f, { 0, -1, 0 }, 10.f };
    MskHit hit; uint8_t occ;
    msk_gpu_intersect(sc, &ray, &hit, 1);
    msk_gpu_occluded(sc, &ray, &occ, 1);
    msk_gpu_scene_destroy(sc);
}

// HDRFilm::image on the device film: one thread per pixel, 20 B in, 16 B out.  Explicit round-to-nearest multiplies and
// adds in the host loop's order (no FMA contraction), IEEE division: bit-identical to msk_gpu_develop and to the
// reference's compiled hdrfilm.cpp.
__global__ void __launch_bounds__(256) k_develop(const float *__restrict__ film, float4 *__restrict__ rgba, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = film + i * 5;
    const float X = p[0], Y = p[1], Z = p[2], A = p[3], W = p[4];
    const float r = __fadd_rn(__fadd_rn(__fmul_rn(3.240479f, X), __fmul_rn(-1.537150f, Y)), __fmul_rn(-0.498535f, Z));
    const float g = __fadd_rn(__fadd_rn(__fmul_rn(-0.969256f, X), __fmul_rn(1.875991f, Y)), __fmul_rn(0.041556f, Z));
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(0.055648f, X), __fmul_rn(-0.204043f, Y)), __fmul_rn(1.057311f, Z));
    const float inv = W != 0.f ? __fdiv_rn(1.f, W) : 0.f;
    rgba[i] = make_float4(__fmul_rn(r, inv), __fmul_rn(g, inv), __fmul_rn(b, inv), __fmul_rn(A, inv));
}

extern "C" {

int msk_gpu_develop_dev(MskScene *s, const float *d_film, float *d_rgba) {
    if (!s || !d_film || !d_rgba) return fail(MSK_ERR_ARG, "null argument");
    DeviceGuard guard(s->ctx->device);
    const size_t n = (size_t) s->d.cam.width * s->d.cam.height;
    k_develop<<<(unsigned) ((n + 255) / 256), 256, 0, s->ctx->stream>>>(d_film, reinterpret_cast<float4 *>(d_rgba), n);
    MSK_CUDA_CHECK(cudaGetLastError());
    return MSK_OK;
}

// HDRFilm::image, hdrfilm.cpp:48-90, for a film that already sits in host memory (the host plugin's Film::develop path:
// a 1 MB image is not worth a round trip over PCIe; msk_gpu_develop_dev is the same arithmetic on a device film)
int msk_gpu_develop(MskScene *s, const float *film, float *rgba) {
    if (!s || !film || !rgba) return fail(MSK_ERR_ARG, "null argument");
    size_t n = (size_t) s->d.cam.width * s->d.cam.height;
    for (size_t i = 0; i < n; ++i) {
        const float *p = film + i * 5;
        float r = 3.240479f * p[0] + -1.537150f * p[1] + -0.498535f * p[2];
        float g = -0.969256f * p[0] + 1.875991f * p[1] + 0.041556f * p[2];
        float b = 0.055648f * p[0] + -0.204043f * p[1] + 1.057311f * p[2];
        float inv = p[4] != 0.f ? 1.f / p[4] : 0.f;
        rgba[i * 4 + 0] = r * inv; rgba[i * 4 + 1] = g * inv; rgba[i * 4 + 2] = b * inv; rgba[i * 4 + 3] = p[3] * inv;
    }
    return MSK_OK;
}

} // extern "C"
