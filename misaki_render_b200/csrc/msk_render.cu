// Wavefront path tracer: replaces the TBB tile loop of SamplingIntegrator::render
// (reference src/librender/integrator.cpp:31-126) and PathTracer::sample
// (src/librender/integrators/path.cpp:23-131).
//
// One batch = every pixel of the film x a run of consecutive samples.  Stages,
// connected by queues in HBM (all launches on one stream, no host round trips
// inside a bounce):
//
//   k_raygen      camera sample -> primary ray + path state             (dense)
//   k_intersect   persistent-thread closest hit; classifies each hit by BSDF type and
//                 appends its queue index to that type's segment (warp match + 1 atomic)
//   k_shade       runs over the type-sorted index list: emission/MIS, Russian roulette,
//                 NEE sample (-> shadow queue), BSDF sample (-> next ray queue,
//                 warp-ballot compaction), state moves to its new compacted slot
//   k_shadow      persistent-thread any hit; unoccluded => L[path] += contribution
//   k_film_*      per-sample XYZ records, then a per-pixel GATHER of the filtered
//                 splats (deterministic, no float atomics)
//
// Sample (pixel p = y*W+x, index s) is seeded with Sampler::seed(p*spp + s) and draws
// in source order, so any partition of the sample range over batches or GPUs
// produces the same per-sample radiance.
#include "msk_render.h"
#include "msk_shading.cuh"
#include "msk_traverse.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace msk {

namespace {

#ifndef MSK_TRAV_MIN_BLOCKS
#define MSK_TRAV_MIN_BLOCKS 6 /* resident 128-thread CTAs per SM the traversal kernels are compiled for (<= 80 registers) */
#endif
constexpr int kNumKeys = 1 + MSK_BSDF_TYPE_COUNT; // 0 = miss, 1 + bsdf type

struct Ctrl {
    uint32_t n_rays[2];
    uint32_t n_shadow;
    uint32_t cursor_isect, cursor_shadow;
    uint32_t type_count[kNumKeys];
    uint32_t pad_;
    unsigned long long total_closest, total_shadow;
    unsigned long long nodes_closest, tris_closest, nodes_shadow, tris_shadow; // MSK_RENDER_TRAVERSAL_STATS
    unsigned long long shaded; // path vertices processed by k_shade
    unsigned long long tail_closest, tail_shadow; // rays traced by k_tail (also included in total_*)
    unsigned long long tail_depth;                // deepest vertex k_tail reached
};

struct Pool {
    uint32_t capacity;
    MskRay  *rays[2];
    float4  *hit;       // t, u, v, prim
    uint32_t *hit_geom;
    float4  *T[2], *WL[2], *AUX[2]; // throughput | wavelengths | eta, prev_pdf, stale_pdf, -
    uint4   *MISC[2];               // rng lo, rng hi, path, depth | flags << 16
    float4  *L;                     // accumulated radiance per path of the batch
    MskRay  *sh_ray;
    float4  *sh_contrib;
    float4  *rec;                   // X, Y, Z, pos.x
    float   *rec_py;
    uint32_t *sorted;               // kNumKeys segments of `capacity` queue indices
    float   *aov;                   // AOV integrator: channel-major per-sample values, aov[c * capacity + i]
    Ctrl    *ctrl;
    // MSK_RAY_SORT (experiment): queue indices of the current ray queue ordered by (origin Morton cell, direction octant)
    uint32_t *rs_keys[2], *rs_vals[2];
};

// AOVIntegrator (aov.cpp:22-29): the requested outputs in channel order
struct AovPlan {
    uint32_t ntypes, nch;
    int32_t  rgba_channel;          // first channel of the nested integrator's RGBA, or -1
    uint8_t  types[MSK_AOV_MAX_CHANNELS];
};

struct BatchParams {
    uint32_t npix, width;
    uint32_t s0, ns;       // first sample index, samples per pixel in this batch
    uint32_t spp;          // of the whole job (seeding)
    uint64_t base_seed;
    int32_t  max_depth, rr_depth, hide_emitters;
    uint32_t tiled;        // path slots enumerate the film in 8x4 tiles (film size a multiple of 8x4), see slot_decode
    uint32_t sw_log2;      // tiled: a warp holds 2^sw_log2 consecutive samples of 32 >> sw_log2 pixels (divides ns)
};

// Path slot i of a batch (0 <= i < npix * ns) -> (linear pixel index, sample offset within the batch).
//
// Untiled: sample-major, pixels linear.  With `tiled` (film a multiple of 8x4) the slots of an 8x4 pixel tile are
// contiguous and a warp -- of k_raygen, of the static traversal of bounce 0, of the first shade and shadow passes --
// covers 2^sw_log2 consecutive SAMPLES of a compact sub-tile of 32 >> sw_log2 pixels (8x4, 4x4, 4x2, 2x2, 2x1, 1x1:
// the low bits of the pixel's Morton code inside the tile).  With 32 samples of ONE pixel per warp the camera rays of
// a warp differ by sub-pixel jitter only: they visit the same nodes and leaves (a static warp of 8x4 pixels of one
// sample ran at 17.9 of 32 lanes per instruction on C2, profiles/r02d), hit the same material and send their shadow
// rays from the same spot.  Only the enumeration order of the paths changes: seeds depend on (pixel, sample) and the
// film records are written per (sample, pixel), so the film is bit-identical.
__device__ __forceinline__ void slot_decode(const BatchParams &bp, uint32_t i, uint32_t &pixel, uint32_t &s_off) {
    if (!bp.tiled) { pixel = i % bp.npix; s_off = i / bp.npix; return; }
    const uint32_t tiles_x = bp.width >> 3, l = i & 31u, w = i >> 5;
    if (bp.sw_log2 > 5u) { // MSK_SAMPLES_PER_WARP=0 (A/B): the round-1 order, sample-major, one 8x4 tile of one sample per warp
        const uint32_t p = i % bp.npix, t = p >> 5;
        s_off = i / bp.npix;
        pixel = ((t / tiles_x) * 4u + (l >> 3)) * bp.width + (t % tiles_x) * 8u + (l & 7u);
        return;
    }
    const uint32_t sw = bp.sw_log2, pw = 5u - sw;          // log2 of samples / pixels per warp
    const uint32_t k = w & ((1u << sw) - 1u), tg = w >> sw; // sub-tile of the 8x4 tile | tile * groups + sample group
    const uint32_t groups = bp.ns >> sw, g = tg % groups;
    uint32_t t = tg / groups, tx, ty;
    if (bp.tiled == 2u) { // the four warps of a 128-thread block cover a 16x8 block of pixels (2x2 tiles)
        const uint32_t groups_x = tiles_x >> 1, q = t >> 2, c = t & 3u;
        tx = (q % groups_x) * 2u + (c & 1u); ty = (q / groups_x) * 2u + (c >> 1);
    } else {
        tx = t % tiles_x; ty = t / tiles_x;
    }
    const uint32_t m = (k << pw) | (l & ((1u << pw) - 1u)); // Morton code in the tile: x0 y0 x1 y1 x2 from bit 0
    const uint32_t x = (m & 1u) | ((m >> 1) & 2u) | ((m >> 2) & 4u), y = ((m >> 1) & 1u) | ((m >> 2) & 2u);
    s_off = (g << sw) + (l >> pw);
    pixel = (ty * 4u + y) * bp.width + tx * 8u + x;
}

constexpr uint32_t kFlagDelta = 1u << 16;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---------------------------------------------------------------------------------------
__global__ void k_begin_batch(Ctrl *c, uint32_t n) {
    c->n_rays[0] = n; c->n_rays[1] = 0; c->n_shadow = 0; c->cursor_isect = 0; c->cursor_shadow = 0;
    for (int i = 0; i < kNumKeys; ++i) c->type_count[i] = 0;
}

// after k_shade + k_shadow of bounce b: make queue `next` current
__global__ void k_end_bounce(Ctrl *c, int cur) {
    c->total_closest += c->n_rays[cur];
    c->total_shadow += c->n_shadow;
    c->shaded += c->n_rays[cur];
    c->n_rays[cur] = 0; c->n_shadow = 0; c->cursor_isect = 0; c->cursor_shadow = 0;
    for (int i = 0; i < kNumKeys; ++i) c->type_count[i] = 0;
}

// ---------------------------------------------------------------------------------------
// render_sample: integrator.cpp:103-126 (first half), perspective.cpp:22-41
// write_state == 0: the first shade pass knows a camera path's throughput (1) and (eta, pdfs) = (1, 0, 0) without reading
// them (k_shade's `first`), so they are not written either: 32 of the 112 bytes per camera sample, and 32 of the ~280 a
// first-bounce vertex moves.
__global__ void __launch_bounds__(256) k_raygen(const __grid_constant__ DScene sc, Pool pool, BatchParams bp, int write_state) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = bp.npix * bp.ns;
    if (i >= n) return;
    uint32_t pixel, s_off;
    slot_decode(bp, i, pixel, s_off);
    const uint32_t s = bp.s0 + s_off;
    uint32_t gx = pixel % bp.width, gy = pixel / bp.width;
    uint64_t rng = pcg_seed((uint64_t) pixel * bp.spp + s + bp.base_seed);
    float jx = next1d(rng), jy = next1d(rng);
    float wav = next1d(rng);
    next1d(rng); next1d(rng); // aperture sample: consumed, unused
    float4 wl, weight;
    sample_wavelength(wav, wl, weight);
    V3 o, d;
    float mint, maxt;
    camera_ray(sc.cam, (float) gx + jx, (float) gy + jy, o, d, mint, maxt);
    float4 *rp = reinterpret_cast<float4 *>(pool.rays[0] + i);
    rp[0] = make_float4(o.x, o.y, o.z, mint);
    rp[1] = make_float4(d.x, d.y, d.z, maxt);
    pool.WL[0][i]   = wl;
    pool.MISC[0][i] = make_uint4((uint32_t) rng, (uint32_t) (rng >> 32), i, 1u);
    if (write_state) {
        pool.T[0][i]   = f4(1.f);
        pool.AUX[0][i] = make_float4(1.f, 0.f, 0.f, 0.f);
    }
    pool.L[i]       = f4(0.f);
}

// ---------------------------------------------------------------------------------------
// Closest hit over the current ray queue (Scene::ray_intersect, scene.cpp:216-253) + classification.
__device__ __forceinline__ void add_traversal_stats(unsigned long long *nodes, unsigned long long *tris, uint32_t cn, uint32_t ct) {
    for (int o = 16; o; o >>= 1) { cn += __shfl_xor_sync(0xffffffffu, cn, o); ct += __shfl_xor_sync(0xffffffffu, ct, o); }
    if (lane_id() == 0) { atomicAdd(nodes, (unsigned long long) cn); atomicAdd(tris, (unsigned long long) ct); }
}

// IO adaptor of trace_queue: reads the current ray queue, writes the hit record.
#ifndef MSK_SORT_IN_COMMIT
#define MSK_SORT_IN_COMMIT 0
#endif
template <bool STATS>
struct IntersectIO {
    const Pool &pool;
    const MskRay *rays;
    const DScene *scp = nullptr;
    const uint32_t *perm = nullptr; // MSK_RAY_SORT: position in the traversal order -> queue index
    uint32_t cn_total = 0, ct_total = 0;
    __device__ __forceinline__ void load(uint32_t q, float4 &ro, float4 &rd) const {
        if (perm) q = __ldg(perm + q);
        const float4 *rp = reinterpret_cast<const float4 *>(rays + q);
        ro = __ldcs(rp); rd = __ldcs(rp + 1);
    }
    __device__ __forceinline__ uint32_t tag(float4 &) const { return 0u; }
    __device__ __forceinline__ void commit(bool have, uint32_t q, const Traversal &s) {
        if (perm && have) q = __ldg(perm + q);
#if MSK_SORT_IN_COMMIT
        const uint32_t done = __ballot_sync(0xffffffffu, have);
#endif
        if (!have) return;
        const bool found = s.is_hit(); // hit <=> tfar != maxt, scene.cpp:234
        pool.hit[q]      = make_float4(found ? s.hit.t : MSK_INF, s.hit.u, s.hit.v, __uint_as_float(s.hit.prim));
        pool.hit_geom[q] = found ? s.hit.geom : 0xffffffffu;
        if (STATS) { cn_total += s.cnt_nodes; ct_total += s.cnt_tris; }
#if MSK_SORT_IN_COMMIT
        const DScene &sc = *scp;
        const uint32_t key = found ? 1u + (uint32_t) sc.bsdfs[sc.meshes[s.hit.geom].bsdf].type : 0u;
        const uint32_t peers = __match_any_sync(done, key);
        const uint32_t leader = __ffs(peers) - 1u;
        const uint32_t rank = __popc(peers & ((1u << lane_id()) - 1u));
        uint32_t slot = 0;
        if (lane_id() == leader) slot = atomicAdd(&pool.ctrl->type_count[key], (uint32_t) __popc(peers));
        slot = __shfl_sync(peers, slot, leader);
        pool.sorted[(size_t) key * pool.capacity + slot + rank] = q;
#endif
    }
};

template <bool STATS, bool PACKETS = false>
__global__ void __launch_bounds__(128, MSK_TRAV_MIN_BLOCKS) k_intersect(const __grid_constant__ DScene sc, Pool pool, int cur, int coherent, const uint32_t *perm) {
    MSK_TRAV_SHARED;
    const Accel ac{ sc.nodes, sc.tris, sc.k47, perm_lut_init(msk_s_perm) };
    Ctrl *c = pool.ctrl;
    IntersectIO<STATS> io{ pool, pool.rays[cur], &sc, perm };
    trace_queue<false, STATS, PACKETS>(ac, MSK_TRAV_SMEM, c->n_rays[cur], &c->cursor_isect, io, coherent);
    if (STATS) add_traversal_stats(&c->nodes_closest, &c->tris_closest, io.cn_total, io.ct_total);
}

// MSK_RAY_SORT (experiment, off by default): key of every ray of the current queue = 8-bit-per-axis Morton cell of its
// origin, then the octant of its direction; entries past the queue length sort to the end.
__device__ __forceinline__ uint32_t spread8(uint32_t v) { // 8 bits -> every third bit
    v = (v | (v << 8)) & 0x00f00fu; v = (v | (v << 4)) & 0x0c30c3u; v = (v | (v << 2)) & 0x249249u;
    return v;
}
__global__ void __launch_bounds__(256) k_ray_keys(const __grid_constant__ DScene sc, Pool pool, int cur, uint32_t n_max, uint32_t *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max) return;
    uint32_t key = 0xffffffffu;
    if (i < pool.ctrl->n_rays[cur]) {
        const float4 *rp = reinterpret_cast<const float4 *>(pool.rays[cur] + i);
        const float4 o = rp[0], d = rp[1];
        const uint32_t qx = (uint32_t) fminf(fmaxf((o.x - sc.bb_lo[0]) * sc.bb_scale[0], 0.f) * 256.f, 255.f);
        const uint32_t qy = (uint32_t) fminf(fmaxf((o.y - sc.bb_lo[1]) * sc.bb_scale[1], 0.f) * 256.f, 255.f);
        const uint32_t qz = (uint32_t) fminf(fmaxf((o.z - sc.bb_lo[2]) * sc.bb_scale[2], 0.f) * 256.f, 255.f);
        const uint32_t oct = (d.x < 0.f ? 4u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 1u : 0u);
        key = (((spread8(qx) << 2) | (spread8(qy) << 1) | spread8(qz)) << 3) | oct;
    }
    keys[i] = key; vals[i] = i;
}

// ---------------------------------------------------------------------------------------
// One path vertex: path.cpp:33-123 re-ordered so that everything that follows the hit of the
// ray spawned at the previous vertex (emitter MIS term :82-108, Russian roulette :116-122)
// runs at the start of the next vertex, in the original order of random draws.
// Material sort (north_star: "sorted by material to curb divergence"): a counting sort of the queue indices by
// key (0 = miss, 1 + BSDF type).  Each block ranks a tile of kSortTile entries in shared memory and reserves
// its share of every key's segment with ONE global atomic per key, so the pass costs a hit_geom read and an
// index write per ray.  Besides making k_shade's warps uniform in material it groups the NEXT bounce's rays by
// the surface they leave (measured: the static traversal of C2 is 1.3x slower on an unsorted queue).
constexpr int kSortThreads = 256, kSortItems = 8, kSortTile = kSortThreads * kSortItems;
__global__ void __launch_bounds__(kSortThreads) k_sort(const __grid_constant__ DScene sc, Pool pool, int cur) {
    Ctrl *c = pool.ctrl;
    const uint32_t total = c->n_rays[cur];
    __shared__ uint32_t s_cnt[kNumKeys], s_base[kNumKeys];
    const uint32_t ntiles = (total + kSortTile - 1) / kSortTile;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t tbase = tile * kSortTile, tn = min((uint32_t) kSortTile, total - tbase);
        if (threadIdx.x < kNumKeys) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t keyrank[kSortItems]; // key << 16 | rank within (tile, key)
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t e = i * kSortThreads + threadIdx.x;
            const bool in = e < tn;
            uint32_t key = 0;
            if (in) {
                const uint32_t geom = __ldcs(pool.hit_geom + tbase + e);
                if (geom != 0xffffffffu) key = 1u + (uint32_t) sc.bsdfs[sc.meshes[geom].bsdf].type;
            }
            const uint32_t active = __ballot_sync(0xffffffffu, in);
            keyrank[i] = 0;
            if (in) {
                const uint32_t peers = __match_any_sync(active, key);
                const uint32_t leader = __ffs(peers) - 1u;
                uint32_t base = 0;
                if (lane_id() == leader) base = atomicAdd(&s_cnt[key], (uint32_t) __popc(peers));
                base = __shfl_sync(peers, base, leader);
                keyrank[i] = (key << 16) | (base + __popc(peers & ((1u << lane_id()) - 1u)));
            }
        }
        __syncthreads();
        if (threadIdx.x < kNumKeys) {
            const uint32_t cnt = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = cnt ? atomicAdd(&c->type_count[threadIdx.x], cnt) : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t e = i * kSortThreads + threadIdx.x;
            if (e < tn) {
                const uint32_t key = keyrank[i] >> 16;
                pool.sorted[(size_t) key * pool.capacity + s_base[key] + (keyrank[i] & 0xffffu)] = tbase + e;
            }
        }
        __syncthreads();
    }
}

#ifndef MSK_SHADE_MIN_BLOCKS
#define MSK_SHADE_MIN_BLOCKS 4 /* 128 registers (28 B of spills) instead of 152: 6.2 -> 5.4 ms on C2, the stage is latency-bound */
#endif
// KEY < 0: one launch walks every key segment of the sorted index list (used for short queues, where launch
// count matters more than code size).  KEY >= 0: the launch handles segment KEY only (0 = miss, 1 + BSDF type),
// the BSDF switch folds to one case, and the kernel gets the register budget of that material alone -- the
// all-materials kernel needs 152 registers and its 6197 static FFMAs thrash the instruction cache
// (stall no_instruction = 3.8 per issue, profiles/r01c_ncu_k_shade.txt).
#ifndef MSK_SHADE_DIFFUSE_BLOCKS
#define MSK_SHADE_DIFFUSE_BLOCKS 4
#endif
#ifndef MSK_SHADE_GLOSSY_BLOCKS
#define MSK_SHADE_GLOSSY_BLOCKS 4
#endif
constexpr int shade_min_blocks(int key) { return key < 0 ? MSK_SHADE_MIN_BLOCKS : (key == 0 ? 8 : (key == 1 ? MSK_SHADE_DIFFUSE_BLOCKS : MSK_SHADE_GLOSSY_BLOCKS)); }
// One path vertex of PathTracer::sample given the hit record of the ray that arrived (shared by the wavefront stage
// k_shade and the per-path tail kernel k_tail).  TYPE >= 0: BSDF type known at compile time.
struct VertexOut {
    bool emit_ray, emit_shadow, add_L;
    MskRay nray, sray;
    float4 nT, nAUX, contrib, L;
    uint4 nMISC;
};
template <int TYPE>
__device__ __forceinline__ void shade_vertex(const DScene &sc, const BatchParams &bp, bool miss, uint32_t geom, float4 hit, float4 rd,
                                             float4 T, float4 wl, uint4 misc, float4 aux, VertexOut &o) {
    bool emit_ray = false, emit_shadow = false;
    MskRay nray, sray;
    float4 nT, nAUX, contrib;
    uint4 nMISC;
    uint64_t rng = (uint64_t) misc.x | ((uint64_t) misc.y << 32);
    const uint32_t path = misc.z;
    const int depth = (int) (misc.w & 0xffffu);
    const bool prev_delta = (misc.w & kFlagDelta) != 0;
    float eta = aux.x;
    const float prev_pdf = aux.y, stale_pdf = aux.z;
    float4 L = f4(0.f);
    bool add_L = false, alive = true;
    const V3 rdir = v3(rd.x, rd.y, rd.z);

    if (miss) { // miss: path.cpp:34-41 (depth 1) / :90-98,:103-108 (BSDF-sampled ray escaped)
        if (sc.environment >= 0) {
            int radiance = sc.emitters[sc.environment].radiance; // constant.cpp:79-81; a miss carries no uv
            if (sc.has_textures) radiance = texture_resolve(sc, radiance, 0.f, 0.f);
            float4 le = spectrum_eval(sc, radiance, wl);
            if (depth == 1) { if (!bp.hide_emitters) { L = T * le; add_L = true; } }
            else { L = T * le * mis_weight(prev_pdf, prev_delta ? 0.f : stale_pdf); add_L = true; }
        }
        alive = false;
    } else {
        const DMeshInfo mi = sc.meshes[geom];
        const Surface sf = make_surface(sc, mi, __float_as_uint(hit.w), hit.y, hit.z);
        const V3 wi = to_local(sf.sh, -rdir);
        // emitter term of the ray that arrived here: weight now, radiance after the single spectrum pass below
        int id_le = -1;
        float le_weight = 1.f;
        if (mi.emitter >= 0 && wi.z > 0.f) { // area.cpp:51-54
            if (depth == 1) { // path.cpp:44-47
                if (!bp.hide_emitters) id_le = sc.emitters[mi.emitter].radiance;
            } else {          // path.cpp:82-88,103-108 with ds.set_query (records.cpp:7-14)
                float emitter_pdf = 0.f;
                if (!prev_delta) {
                    float dp = fabsf(dot(rdir, sf.sh.n));
                    emitter_pdf = mi.inv_area * ((dp != 0.f) ? (hit.x * hit.x) / dp : 0.f);
                    if (sc.nemitters > 1) emitter_pdf *= 1.f / (float) sc.nemitters;
                }
                le_weight = mis_weight(prev_pdf, emitter_pdf);
                id_le = sc.emitters[mi.emitter].radiance;
            }
            if (id_le >= 0 && sc.has_textures) id_le = texture_resolve(sc, id_le, sf.uvx, sf.uvy);
        }
        const float4 T_in = T; // the emitter term uses the throughput before this vertex's roulette division
        if (depth > 1 && depth >= bp.rr_depth) { // path.cpp:116-122 of the previous iteration
            float qq = fminf(hmax(T) * eta * eta, 0.95f);
            if (next1d(rng) >= qq) alive = false;
            else T = T / qq;
        }
        if (alive && bp.max_depth > 0 && depth >= bp.max_depth) alive = false; // path.cpp:48-49
        MskBsdf bsdf = sc.bsdfs[mi.bsdf];
        NeeSample ns;
        ns.pdf = 0.f; ns.radiance = -1; ns.stale_pdf = 0.f;
        bool nee = false;
        if (alive) {
            if (sc.has_textures) bsdf_resolve_textures(sc, bsdf, sf.uvx, sf.uvy); // Texture::eval(si), checkerboard.cpp:25-31
            if (bsdf_is_smooth(TYPE >= 0 ? TYPE : bsdf.type)) { // path.cpp:56-67
                float sx = next1d(rng), sy = next1d(rng);
                ns = sample_emitter_direct(sc, sf.p, sx, sy);
                nee = ns.pdf != 0.f;
            }
        }
        BsdfSpectra sp;
        float4 le, ln;
        eval_vertex_spectra<TYPE>(sc, bsdf, alive, id_le, nee ? ns.radiance : -1, wl, sp, le, ln);
        if (id_le >= 0) { L = T_in * le * le_weight; add_L = true; } // le_weight == 1 at depth 1
        if (alive) {
            const float new_stale = ns.stale_pdf;
            const float tmin_spawn = (1.f + max_abs(sf.p)) * kRayEpsilon; // interaction.h:40-44, scene.cpp:91-93
            if (nee) {
                V3 wo = to_local(sf.sh, ns.d);
                float4 bval; float bpdf;
                bsdf_eval_pdf<TYPE>(sp, bsdf, wi, wo, bval, bpdf);
                float w = mis_weight(ns.pdf, bpdf);
                contrib = T * nee_value(ns, ln) * bval * w;
                if (!is_zero(contrib)) {
                    emit_shadow = true;
                    sray.o[0] = sf.p.x; sray.o[1] = sf.p.y; sray.o[2] = sf.p.z;
                    sray.tmin = shadow_tmin(sf.p.x, sf.p.y, sf.p.z);
                    sray.d[0] = ns.d.x; sray.d[1] = ns.d.y; sray.d[2] = ns.d.z;
                    sray.tmax = ns.dist * (1.f - kShadowEpsilon);
                }
            }
            float s1 = next1d(rng), s2x = next1d(rng), s2y = next1d(rng); // path.cpp:71-72, left to right
            BsdfSample bs = bsdf_sample<TYPE>(sp, bsdf, wi, s1, s2x, s2y);
            if (is_zero(bs.weight)) alive = false; // failed sample: nothing downstream can contribute
            else {
                V3 wo = to_world(sf.sh, bs.wo);
                T = T * bs.weight;
                eta *= bs.eta;
                emit_ray = true;
                nray.o[0] = sf.p.x; nray.o[1] = sf.p.y; nray.o[2] = sf.p.z; nray.tmin = tmin_spawn;
                nray.d[0] = wo.x; nray.d[1] = wo.y; nray.d[2] = wo.z; nray.tmax = MSK_INF;
                nT = T;
                nMISC = make_uint4((uint32_t) rng, (uint32_t) (rng >> 32), path,
                                   (uint32_t) (depth + 1) | ((bs.type & BF_Delta) ? kFlagDelta : 0u));
                nAUX = make_float4(eta, bs.pdf, new_stale, 0.f);
            }
        }
    }
    o.emit_ray = emit_ray; o.emit_shadow = emit_shadow; o.add_L = add_L;
    o.nray = nray; o.sray = sray; o.nT = nT; o.nAUX = nAUX; o.contrib = contrib; o.L = L; o.nMISC = nMISC;
}

template <int KEY>
__global__ void __launch_bounds__(128, shade_min_blocks(KEY)) k_shade(const __grid_constant__ DScene sc, Pool pool, BatchParams bp, int cur, int first) {
    Ctrl *c = pool.ctrl;
    const int nxt = cur ^ 1;
    constexpr int TYPE = KEY > 0 ? KEY - 1 : -1;
    uint32_t counts[kNumKeys], total = 0;
    if (KEY >= 0) total = c->type_count[KEY];
    else {
#pragma unroll
        for (int k = 0; k < kNumKeys; ++k) { counts[k] = c->type_count[k]; total += counts[k]; }
    }
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (total + stride - 1) / stride;
    // (Tried: reading this thread's next queue index one iteration ahead and pulling that vertex's state lines into L2 with
    // prefetch.global.L2 while the current vertex is shaded -- the stage runs at 25 % occupancy with DRAM at 46 % of peak,
    // profiles/r02d_ncu_k_shade.txt.  C2 shade 2.31 -> 2.41 ms: the extra index load and seven prefetches per vertex cost
    // more issue slots and registers than the L2 hits save.)
    const uint32_t tid0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t it = 0; it < rounds; ++it) {
        uint32_t idx = it * stride + tid0;
        bool valid = idx < total;
        bool emit_ray = false, emit_shadow = false;
        MskRay nray, sray;
        float4 nT, nWL, nAUX, contrib;
        uint4 nMISC;
        uint32_t path = 0;
        if (valid) {
            uint32_t key = KEY >= 0 ? (uint32_t) KEY : 0u, j = idx;
            if (KEY < 0) {
#pragma unroll
                for (int k = 0; k < kNumKeys - 1; ++k)
                    if (key == (uint32_t) k && j >= counts[k]) { j -= counts[k]; key = k + 1; }
            }
            const uint32_t q = pool.sorted[(size_t) key * pool.capacity + j];
            const float4 hit = pool.hit[q];
            const float4 rd  = reinterpret_cast<const float4 *>(pool.rays[cur] + q)[1];
            // (first: the vertices of camera rays -- k_raygen did not write their constant throughput and eta / pdf record)
            const float4 T = first ? f4(1.f) : pool.T[cur][q];
            const float4 wl = pool.WL[cur][q];
            const uint4 misc = pool.MISC[cur][q];
            const float4 aux = first ? make_float4(1.f, 0.f, 0.f, 0.f) : pool.AUX[cur][q];
            path = misc.z;
            VertexOut vo;
            const bool miss = KEY == 0 || (KEY < 0 && key == 0);
            shade_vertex<TYPE>(sc, bp, miss, miss ? 0xffffffffu : pool.hit_geom[q], hit, rd, T, wl, misc, aux, vo);
            emit_ray = vo.emit_ray; emit_shadow = vo.emit_shadow;
            nray = vo.nray; sray = vo.sray; nT = vo.nT; nWL = wl; nAUX = vo.nAUX; contrib = vo.contrib; nMISC = vo.nMISC;
            if (vo.add_L) { float4 acc = pool.L[path]; pool.L[path] = acc + vo.L; }
        }
        // compaction: one atomic per warp and queue
        uint32_t m_ray = __ballot_sync(0xffffffffu, emit_ray), m_sh = __ballot_sync(0xffffffffu, emit_shadow);
        uint32_t base_ray = 0, base_sh = 0;
        if (lane_id() == 0) {
            if (m_ray) base_ray = atomicAdd(&c->n_rays[nxt], (uint32_t) __popc(m_ray));
            if (m_sh) base_sh = atomicAdd(&c->n_shadow, (uint32_t) __popc(m_sh));
        }
        base_ray = __shfl_sync(0xffffffffu, base_ray, 0);
        base_sh  = __shfl_sync(0xffffffffu, base_sh, 0);
        const uint32_t below = (1u << lane_id()) - 1u;
        if (emit_ray) {
            uint32_t o = base_ray + __popc(m_ray & below);
            float4 *rp = reinterpret_cast<float4 *>(pool.rays[nxt] + o);
            rp[0] = make_float4(nray.o[0], nray.o[1], nray.o[2], nray.tmin);
            rp[1] = make_float4(nray.d[0], nray.d[1], nray.d[2], nray.tmax);
            pool.T[nxt][o] = nT; pool.WL[nxt][o] = nWL; pool.MISC[nxt][o] = nMISC; pool.AUX[nxt][o] = nAUX;
        }
        if (emit_shadow) {
            uint32_t o = base_sh + __popc(m_sh & below);
            float4 *rp = reinterpret_cast<float4 *>(pool.sh_ray + o);
            rp[0] = make_float4(sray.o[0], sray.o[1], sray.o[2], __uint_as_float(path)); // tmin = shadow_tmin(o): ShadowIO::tag
            rp[1] = make_float4(sray.d[0], sray.d[1], sray.d[2], sray.tmax);
            pool.sh_contrib[o] = contrib;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Volumetric path tracer: one iteration of VolumetricPathTracer::sample (integrators/volpath.cpp:40-165) per queue
// entry -- free-flight sampling in the current medium (media/homogeneous.cpp:21-53), then either a medium
// scattering event (attenuated NEE + isotropic phase sample, phase/isotropic.cpp) or the surface interaction
// (emitter term, attenuated NEE, BSDF sample, medium transition), then Russian roulette.  The restatement decisions
// for the stale RGB API are listed above oracle.cpp's volpath_sample and in DESIGN.md; the kernel mirrors the
// oracle statement by statement, including the order of random draws.
// Path state beyond the path tracer's: AUX.w carries (medium + 1) | channel << 8 | scattered << 10 | emitted << 11.
constexpr uint32_t kVolChannelShift = 8, kVolScattered = 1u << 10, kVolEmitted = 1u << 11;

__device__ __forceinline__ float spec_mean(float4 v) { return ((v.x + v.z) + (v.y + v.w)) / 4.f; } // Eigen packet reduction
__device__ __forceinline__ float tr1(float st, float d) { return st == 0.f ? 1.f : expf(st * (-d)); }
__device__ __forceinline__ float4 medium_tr(float4 st, float d) { // homogeneous.cpp:48,56-59
    return make_float4(tr1(st.x, d), tr1(st.y, d), tr1(st.z, d), tr1(st.w, d));
}
__device__ __forceinline__ float4 medium_sigma_t(const DScene &sc, int medium, float4 wl, float4 &sigma_s) {
    const MskMedium m = sc.media[medium];
    const float4 sa = spectrum_eval(sc, m.sigma_a, wl);
    sigma_s = spectrum_eval(sc, m.sigma_s, wl);
    return sigma_s + sa; // homogeneous.cpp:17
}

// One iteration of VolumetricPathTracer::sample given the hit record of the current ray (shared by k_shade_vol and
// the per-path tail kernel).
__device__ __forceinline__ void shade_vertex_vol(const DScene &sc, const BatchParams &bp, bool miss, uint32_t geom, float4 hit, float4 ro,
                                                 float4 rd, float4 T, float4 wl, uint4 misc, float4 aux, VertexOut &o) {
    bool emit_ray = false, emit_shadow = false;
    MskRay nray, sray;
    float4 nT, nAUX, contrib;
    uint4 nMISC;
    uint64_t rng = (uint64_t) misc.x | ((uint64_t) misc.y << 32);
    const uint32_t path = misc.z;
    const int depth = (int) (misc.w & 0xffffu);
    float eta = aux.x;
    uint32_t bits = __float_as_uint(aux.w);
    if (depth == 1) { // volpath.cpp:35-39: initial medium = sensor->medium(), emitted_radiance = true, channel draw
        const uint32_t channel = min((uint32_t) (next1d(rng) * 4.f), 3u);
        bits = (uint32_t) (sc.sensor_medium + 1) | (channel << kVolChannelShift) | kVolEmitted;
    }
    int medium = (int) (bits & 0xffu) - 1;
    const uint32_t channel = (bits >> kVolChannelShift) & 3u;
    bool scattered = (bits & kVolScattered) != 0, emitted = (bits & kVolEmitted) != 0;
    float4 L = f4(0.f);
    bool add_L = false, alive = true;
    const V3 rorg = v3(ro.x, ro.y, ro.z), rdir = v3(rd.x, rd.y, rd.z);

    // ---- free flight: medium->sample_distance(ray::spawn(ray, 0, si.t), next1d, channel), volpath.cpp:41-43
    bool ms_flag = false;
    float4 sigma_t = f4(0.f), sigma_s = f4(0.f), tr = f4(1.f);
    float ms_pdf = 1.f;
    V3 msp = rorg;
    if (medium >= 0) {
        const float sample = next1d(rng);
        sigma_t = medium_sigma_t(sc, medium, wl, sigma_s);
        const float st_c = channel == 0 ? sigma_t.x : (channel == 1 ? sigma_t.y : (channel == 2 ? sigma_t.z : sigma_t.w));
        float dist = -logf(1.f - sample) / st_c;
        if (dist < hit.x) { // ray.maxt - ray.mint with mint = 0, maxt = si.t
            msp = rorg + rdir * dist;
            if (msp.x == rorg.x && msp.y == rorg.y && msp.z == rorg.z) ms_pdf = spec_mean(medium_tr(sigma_t, dist));
            else { ms_pdf = spec_mean(medium_tr(sigma_t, dist) * sigma_t); ms_flag = true; }
        } else {
            dist = hit.x;
            ms_pdf = spec_mean(medium_tr(sigma_t, dist));
        }
        tr = medium_tr(sigma_t, dist);
        if (hmax(tr) < 1e-20f) tr = f4(0.f);
    }

    if (ms_flag) { // ---- medium scattering event, volpath.cpp:44-74
        T = T * (sigma_s * tr / ms_pdf);
        const float sx = next1d(rng), sy = next1d(rng);
        NeeSample ns = sample_emitter_direct(sc, msp, sx, sy);
        if (ns.pdf != 0.f) { // scene.cpp:136-138 + isotropic phase value 1 / 4 pi
            const float4 ln = spectrum_eval(sc, ns.radiance, wl);
            contrib = T * (nee_value(ns, ln) * medium_tr(sigma_t, ns.dist)) * kInvFourPi;
            if (!is_zero(contrib)) {
                emit_shadow = true;
                sray.o[0] = msp.x; sray.o[1] = msp.y; sray.o[2] = msp.z;
                sray.tmin = shadow_tmin(msp.x, msp.y, msp.z); // scaled like scene.cpp:91-93 (see oracle.cpp, volpath notes)
                sray.d[0] = ns.d.x; sray.d[1] = ns.d.y; sray.d[2] = ns.d.z;
                sray.tmax = ns.dist * (1.f - kShadowEpsilon);
            }
        }
        if (depth + 1 >= bp.max_depth && bp.max_depth > 0) alive = false;
        else {
            const float px = next1d(rng), py = next1d(rng);
            const V3 wo = square_to_uniform_sphere(px, py);
            emit_ray = true;
            nray.o[0] = msp.x; nray.o[1] = msp.y; nray.o[2] = msp.z; nray.tmin = kRayEpsilon;
            nray.d[0] = wo.x; nray.d[1] = wo.y; nray.d[2] = wo.z; nray.tmax = MSK_INF;
            scattered = true;
        }
    } else {       // ---- surface interaction, volpath.cpp:75-155
        if (medium >= 0) T = T * (tr / ms_pdf);
        if (miss) { // escaped, :82-93
            if (emitted && (!bp.hide_emitters || scattered)) {
                float4 le = f4(0.f);
                if (sc.environment >= 0) {
                    int radiance = sc.emitters[sc.environment].radiance;
                    if (sc.has_textures) radiance = texture_resolve(sc, radiance, 0.f, 0.f);
                    le = spectrum_eval(sc, radiance, wl);
                }
                L = T * le;
                if (medium >= 0) L = L * medium_tr(sigma_t, rd.w - ro.w); // eval_transmittance(ray): exp(sigma_t (mint - maxt))
                add_L = true;
            }
            alive = false;
        } else {
            const DMeshInfo mi = sc.meshes[geom];
            const Surface sf = make_surface(sc, mi, __float_as_uint(hit.w), hit.y, hit.z);
            const V3 wi = to_local(sf.sh, -rdir);
            int id_le = -1;
            if (mi.emitter >= 0 && emitted && (!bp.hide_emitters || scattered) && wi.z > 0.f) { // :95-98, area.cpp:51-54
                id_le = sc.emitters[mi.emitter].radiance;
                if (sc.has_textures) id_le = texture_resolve(sc, id_le, sf.uvx, sf.uvy);
            }
            MskBsdf bsdf = sc.bsdfs[mi.bsdf];
            if (sc.has_textures) bsdf_resolve_textures(sc, bsdf, sf.uvx, sf.uvy);
            NeeSample ns;
            ns.pdf = 0.f; ns.radiance = -1;
            if (bsdf_is_smooth(bsdf.type)) { // :104-115: attenuated NEE, added without the MIS weight
                const float sx = next1d(rng), sy = next1d(rng);
                ns = sample_emitter_direct(sc, sf.p, sx, sy);
            }
            const bool nee = ns.pdf != 0.f;
            BsdfSpectra sp;
            float4 le, ln;
            eval_vertex_spectra<-1, true>(sc, bsdf, true, id_le, nee ? ns.radiance : -1, wl, sp, le, ln);
            if (id_le >= 0) { L = T * le; add_L = true; }
            if (nee) {
                const V3 wo = to_local(sf.sh, ns.d);
                float4 bval; float bpdf;
                bsdf_eval_pdf<-1>(sp, bsdf, wi, wo, bval, bpdf);
                float4 ev = nee_value(ns, ln);
                if (medium >= 0) ev = ev * medium_tr(sigma_t, ns.dist);
                contrib = T * ev * bval;
                if (!is_zero(contrib)) {
                    emit_shadow = true;
                    sray.o[0] = sf.p.x; sray.o[1] = sf.p.y; sray.o[2] = sf.p.z;
                    sray.tmin = shadow_tmin(sf.p.x, sf.p.y, sf.p.z);
                    sray.d[0] = ns.d.x; sray.d[1] = ns.d.y; sray.d[2] = ns.d.z;
                    sray.tmax = ns.dist * (1.f - kShadowEpsilon);
                }
            }
            const float s1 = next1d(rng), s2x = next1d(rng), s2y = next1d(rng);
            BsdfSample bs = bsdf_sample<-1>(sp, bsdf, wi, s1, s2x, s2y);
            if (is_zero(bs.weight)) alive = false; // :122-123
            else {
                emitted = false;
                bool recursive = depth + 1 < bp.max_depth || bp.max_depth < 0;
                if ((depth < bp.max_depth || bp.max_depth < 0) && (bs.type & BF_Delta)) { emitted = true; recursive = true; } // :129-135
                if (!recursive) alive = false;
                else {
                    const V3 wo = to_world(sf.sh, bs.wo);
                    T = T * bs.weight;
                    eta *= bs.eta;
                    const int interior = (int) ((mi.flags >> 8) & 0xffu) - 1, exterior = (int) ((mi.flags >> 16) & 0xffu) - 1;
                    if (interior >= 0 || exterior >= 0) medium = dot(wo, sf.n) > 0.f ? exterior : interior; // interaction.cpp:10-13
                    emit_ray = true;
                    nray.o[0] = sf.p.x; nray.o[1] = sf.p.y; nray.o[2] = sf.p.z; nray.tmin = (1.f + max_abs(sf.p)) * kRayEpsilon;
                    nray.d[0] = wo.x; nray.d[1] = wo.y; nray.d[2] = wo.z; nray.tmax = MSK_INF;
                    scattered = true;
                }
            }
        }
    }
    if (alive && emit_ray && depth + 1 >= bp.rr_depth) { // :158-164
        const float qq = fminf(hmax(T) * eta * eta, 0.95f);
        if (next1d(rng) >= qq) emit_ray = false;
        else T = T / qq;
    }
    if (bp.max_depth > 0 && depth + 1 > bp.max_depth) emit_ray = false; // loop condition :40
    if (!alive) emit_ray = false;
    if (emit_ray) {
        bits = (uint32_t) (medium + 1) | (channel << kVolChannelShift) | (scattered ? kVolScattered : 0u) | (emitted ? kVolEmitted : 0u);
        nT = T;
        nMISC = make_uint4((uint32_t) rng, (uint32_t) (rng >> 32), path, (uint32_t) (depth + 1));
        nAUX = make_float4(eta, 0.f, 0.f, __uint_as_float(bits));
    }
    o.emit_ray = emit_ray; o.emit_shadow = emit_shadow; o.add_L = add_L;
    o.nray = nray; o.sray = sray; o.nT = nT; o.nAUX = nAUX; o.contrib = contrib; o.L = L; o.nMISC = nMISC;
}

#ifndef MSK_VOL_MIN_BLOCKS
#define MSK_VOL_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, MSK_VOL_MIN_BLOCKS) k_shade_vol(const __grid_constant__ DScene sc, Pool pool, BatchParams bp, int cur) {
    Ctrl *c = pool.ctrl;
    const int nxt = cur ^ 1;
    uint32_t counts[kNumKeys], total = 0;
#pragma unroll
    for (int k = 0; k < kNumKeys; ++k) { counts[k] = c->type_count[k]; total += counts[k]; }
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (total + stride - 1) / stride;
    for (uint32_t it = 0; it < rounds; ++it) {
        uint32_t idx = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool valid = idx < total;
        bool emit_ray = false, emit_shadow = false;
        MskRay nray, sray;
        float4 nT, nWL, nAUX, contrib;
        uint4 nMISC;
        uint32_t path = 0;
        if (valid) {
            uint32_t key = 0, j = idx;
#pragma unroll
            for (int k = 0; k < kNumKeys - 1; ++k)
                if (key == (uint32_t) k && j >= counts[k]) { j -= counts[k]; key = k + 1; }
            const uint32_t q = pool.sorted[(size_t) key * pool.capacity + j];
            const float4 hit = pool.hit[q];
            const float4 ro = reinterpret_cast<const float4 *>(pool.rays[cur] + q)[0];
            const float4 rd = reinterpret_cast<const float4 *>(pool.rays[cur] + q)[1];
            float4 T = pool.T[cur][q];
            const float4 wl = pool.WL[cur][q];
            const uint4 misc = pool.MISC[cur][q];
            const float4 aux = pool.AUX[cur][q];
            path = misc.z;
            VertexOut vo;
            shade_vertex_vol(sc, bp, key == 0, key == 0 ? 0xffffffffu : pool.hit_geom[q], hit, ro, rd, T, wl, misc, aux, vo);
            emit_ray = vo.emit_ray; emit_shadow = vo.emit_shadow;
            nray = vo.nray; sray = vo.sray; nT = vo.nT; nWL = wl; nAUX = vo.nAUX; contrib = vo.contrib; nMISC = vo.nMISC;
            if (vo.add_L) { float4 acc = pool.L[path]; pool.L[path] = acc + vo.L; }
        }
        uint32_t m_ray = __ballot_sync(0xffffffffu, emit_ray), m_sh = __ballot_sync(0xffffffffu, emit_shadow);
        uint32_t base_ray = 0, base_sh = 0;
        if (lane_id() == 0) {
            if (m_ray) base_ray = atomicAdd(&c->n_rays[nxt], (uint32_t) __popc(m_ray));
            if (m_sh) base_sh = atomicAdd(&c->n_shadow, (uint32_t) __popc(m_sh));
        }
        base_ray = __shfl_sync(0xffffffffu, base_ray, 0);
        base_sh  = __shfl_sync(0xffffffffu, base_sh, 0);
        const uint32_t below = (1u << lane_id()) - 1u;
        if (emit_ray) {
            uint32_t o = base_ray + __popc(m_ray & below);
            float4 *rp = reinterpret_cast<float4 *>(pool.rays[nxt] + o);
            rp[0] = make_float4(nray.o[0], nray.o[1], nray.o[2], nray.tmin);
            rp[1] = make_float4(nray.d[0], nray.d[1], nray.d[2], nray.tmax);
            pool.T[nxt][o] = nT; pool.WL[nxt][o] = nWL; pool.MISC[nxt][o] = nMISC; pool.AUX[nxt][o] = nAUX;
        }
        if (emit_shadow) {
            uint32_t o = base_sh + __popc(m_sh & below);
            float4 *rp = reinterpret_cast<float4 *>(pool.sh_ray + o);
            rp[0] = make_float4(sray.o[0], sray.o[1], sray.o[2], __uint_as_float(path)); // tmin = shadow_tmin(o): ShadowIO::tag
            rp[1] = make_float4(sray.d[0], sray.d[1], sray.d[2], sray.tmax);
            pool.sh_contrib[o] = contrib;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Tail of an unbounded-depth job.  Once Russian roulette has thinned the queue, a bounce is a handful of
// latency-bound launches over a few hundred thousand rays: on C2 the 17 bounces after depth 6 carry < 5 % of the
// rays and took 1.28 ms of the 17.6 ms step (75 us per bounce, profiles/r01g_bounces_c2.txt).  k_tail finishes
// those paths in ONE launch: every thread takes a surviving path and runs intersect -> shade -> shadow ray in a loop
// until the path dies, with the same device functions (traverse, shade_vertex) and the same order of additions to
// L[path] as the wavefront stages, so the film is bit-identical.  Divergence is irrelevant at this size; what
// matters is that there is no queue traffic, no sort and no launch boundary between the vertices of a path.
template <bool STATS, bool VOL>
__global__ void __launch_bounds__(kTravThreads) k_tail(const __grid_constant__ DScene sc, Pool pool, BatchParams bp, int cur) {
    MSK_TRAV_SHARED;
    const Accel ac{ sc.nodes, sc.tris, sc.k47, perm_lut_init(msk_s_perm) };
    MSK_TRAV_LOCAL_STACK;
    TravStack stack(msk_local_stack, msk_s_stack);
    Ctrl *c = pool.ctrl;
    const uint32_t n = c->n_rays[cur];
    uint32_t n_closest = 0, n_shadow = 0, max_depth = 0, cn = 0, ct = 0, sn = 0, stt = 0;
    // persistent lanes: a lane whose path has died takes the next surviving path (path lengths under Russian roulette
    // are geometric, so a static assignment leaves most lanes of a warp idle behind its longest path)
    bool have = false, more = true;
    float4 ro, rd, T, aux, wl, L;
    uint4 misc;
    uint32_t path = 0;
    const uint32_t below = (1u << lane_id()) - 1u;
    for (;;) {
        const uint32_t want = __ballot_sync(0xffffffffu, !have && more);
        if (want) {
            uint32_t base = 0;
            if (lane_id() == (uint32_t) (__ffs(want) - 1)) base = atomicAdd(&c->cursor_isect, (uint32_t) __popc(want));
            base = __shfl_sync(0xffffffffu, base, __ffs(want) - 1);
            if (!have && more) {
                const uint32_t i = base + __popc(want & below);
                if (i < n) {
                    const float4 *rp = reinterpret_cast<const float4 *>(pool.rays[cur] + i);
                    ro = rp[0]; rd = rp[1];
                    T = pool.T[cur][i]; aux = pool.AUX[cur][i]; wl = pool.WL[cur][i]; misc = pool.MISC[cur][i];
                    path = misc.z;
                    L = pool.L[path];
                    have = true;
                } else {
                    more = false;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, have) == 0u) break;
        if (have) {
            RayHit h;
            uint32_t a = 0, b = 0;
            const bool found = traverse<false, STATS>(ac, stack, ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w, h, &a, &b);
            n_closest++;
            if (STATS) { cn += a; ct += b; }
            max_depth = max(max_depth, misc.w & 0xffffu);
            // the hit record exactly as IntersectIO::commit stores it
            const float4 hit = make_float4(found ? h.t : MSK_INF, found ? h.u : 0.f, found ? h.v : 0.f, __uint_as_float(found ? h.prim : 0xffffffffu));
            VertexOut vo;
            if (VOL) shade_vertex_vol(sc, bp, !found, found ? h.geom : 0xffffffffu, hit, ro, rd, T, wl, misc, aux, vo);
            else shade_vertex<-1>(sc, bp, !found, found ? h.geom : 0xffffffffu, hit, rd, T, wl, misc, aux, vo);
            if (vo.add_L) L = L + vo.L;
            if (vo.emit_shadow) { // k_shadow: unoccluded => L += contribution
                RayHit sh;
                const bool occluded = traverse<true, STATS>(ac, stack, vo.sray.o[0], vo.sray.o[1], vo.sray.o[2], vo.sray.d[0], vo.sray.d[1],
                                                            vo.sray.d[2], vo.sray.tmin, vo.sray.tmax, sh, &a, &b);
                n_shadow++;
                if (STATS) { sn += a; stt += b; }
                if (!occluded) L = L + vo.contrib;
            }
            if (vo.emit_ray) {
                ro = make_float4(vo.nray.o[0], vo.nray.o[1], vo.nray.o[2], vo.nray.tmin);
                rd = make_float4(vo.nray.d[0], vo.nray.d[1], vo.nray.d[2], vo.nray.tmax);
                T = vo.nT; aux = vo.nAUX; misc = vo.nMISC;
            } else {
                pool.L[path] = L;
                have = false;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        n_closest += __shfl_xor_sync(0xffffffffu, n_closest, o); n_shadow += __shfl_xor_sync(0xffffffffu, n_shadow, o);
        max_depth = max(max_depth, __shfl_xor_sync(0xffffffffu, max_depth, o));
    }
    if (lane_id() == 0 && (n_closest | n_shadow)) {
        atomicAdd(&c->total_closest, (unsigned long long) n_closest); atomicAdd(&c->total_shadow, (unsigned long long) n_shadow);
        atomicAdd(&c->tail_closest, (unsigned long long) n_closest); atomicAdd(&c->tail_shadow, (unsigned long long) n_shadow);
        atomicAdd(&c->shaded, (unsigned long long) n_closest);
        atomicMax(&c->tail_depth, (unsigned long long) max_depth);
    }
    if (STATS) { add_traversal_stats(&c->nodes_closest, &c->tris_closest, cn, ct); add_traversal_stats(&c->nodes_shadow, &c->tris_shadow, sn, stt); }
}

// after k_tail: the queue it consumed is empty
__global__ void k_end_tail(Ctrl *c, int cur) {
    c->n_rays[cur] = 0; c->n_rays[cur ^ 1] = 0; c->n_shadow = 0; c->cursor_isect = 0; c->cursor_shadow = 0;
    for (int i = 0; i < kNumKeys; ++i) c->type_count[i] = 0;
}

// ---------------------------------------------------------------------------------------
// NEE visibility (Scene::ray_test, scene.cpp:90-98,255-273) fused with the accumulation of
// the NEE term (path.cpp:63-66).
template <bool STATS>
struct ShadowIO {
    const Pool &pool;
    uint32_t cn_total = 0, ct_total = 0;
    __device__ __forceinline__ void load(uint32_t q, float4 &ro, float4 &rd) const {
        const float4 *rp = reinterpret_cast<const float4 *>(pool.sh_ray + q);
        ro = __ldcs(rp); rd = __ldcs(rp + 1);
    }
    // The queue record carries the path index where the ray's tmin would be: every producer sets tmin to the spawn offset
    // of its origin (shadow_tmin), which is recomputed here, so commit() has the address of L[path] without a dependent
    // load (ncu: the sh_path -> L[path] chain was 11 % of the stall samples of the bounce-0 launch, profiles/r03d_ncu_k_shadow_b0.txt).
    __device__ __forceinline__ uint32_t tag(float4 &ro) const {
        const uint32_t path = __float_as_uint(ro.w);
        ro.w = shadow_tmin(ro.x, ro.y, ro.z);
        return path;
    }
    __device__ __forceinline__ void commit(bool have, uint32_t q, const Traversal &s) {
        if (!have) return;
        if (STATS) { cn_total += s.cnt_nodes; ct_total += s.cnt_tris; }
        if (!s.is_hit()) { // unoccluded (scene.cpp:272): one shadow ray per path and bounce, so no atomics
            const float4 acc = pool.L[s.tag];
            pool.L[s.tag] = acc + __ldcs(pool.sh_contrib + q);
        }
    }
};

template <bool STATS>
__global__ void __launch_bounds__(128, MSK_TRAV_MIN_BLOCKS) k_shadow(const __grid_constant__ DScene sc, Pool pool, int coherent) {
    MSK_TRAV_SHARED;
    const Accel ac{ sc.nodes, sc.tris, sc.k47, perm_lut_init(msk_s_perm) };
    Ctrl *c = pool.ctrl;
    ShadowIO<STATS> io{ pool };
    trace_queue<true, STATS>(ac, MSK_TRAV_SMEM, c->n_shadow, &c->cursor_shadow, io, coherent);
    if (STATS) add_traversal_stats(&c->nodes_shadow, &c->tris_shadow, io.cn_total, io.ct_total);
}

// ---------------------------------------------------------------------------------------
// Film.  render_sample tail (integrator.cpp:115-125): xyz = spectrum_to_xyz(result * ray_weight).
__global__ void __launch_bounds__(256) k_film_records(const __grid_constant__ DScene sc, Pool pool, BatchParams bp, int rgba_channel) {
    // Records are stored by (sample, pixel) for the gather; with several samples of a pixel in one warp (slot_decode)
    // the 32 lanes of a warp would write 32 different sample planes (measured: film 1.04 -> 1.23 ms on C2, 14.5 -> 21.5
    // on C3).  The 256 records of a block -- a compact pixel block times a run of samples -- are therefore exchanged
    // through shared memory so that consecutive threads write the pixels of a row of one sample plane.
    __shared__ float4 s_rec[256];
    __shared__ float s_py[256];
    __shared__ uint32_t s_r[256];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = bp.npix * bp.ns;
    const bool exchange = bp.tiled && bp.sw_log2 >= 1u && bp.sw_log2 <= 5u;
    float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
    float py = 0.f;
    uint32_t r = 0xffffffffu;
    if (i < n) {
        uint32_t pixel, s_off;
        slot_decode(bp, i, pixel, s_off);
        const uint32_t s = bp.s0 + s_off;
        uint32_t gx = pixel % bp.width, gy = pixel / bp.width;
        r = s_off * bp.npix + pixel; // record index: the film gathers read records by (sample, pixel)
        uint64_t rng = pcg_seed((uint64_t) pixel * bp.spp + s + bp.base_seed);
        float jx = next1d(rng), jy = next1d(rng);
        float wav = next1d(rng);
        float4 wl, weight;
        sample_wavelength(wav, wl, weight);
        const float4 spec = pool.L[i];
        float4 result = spec * weight;
        float X, Y, Z;
        spectrum_to_xyz(sc, result, wl, X, Y, Z);
        rec = make_float4(X, Y, Z, (float) gx + jx);
        py = (float) gy + jy;
        if (rgba_channel >= 0) { // aov.cpp:124-140: xyz_to_srgb(spectrum_to_xyz(spec)) of the nested integrator, before ray_weight
            float x, y, z;
            spectrum_to_xyz(sc, spec, wl, x, y, z);
            float *a = pool.aov + (size_t) rgba_channel * pool.capacity + r;
            a[0]                         = 3.240479f * x + -1.537150f * y + -0.498535f * z;
            a[pool.capacity]             = -0.969256f * x + 1.875991f * y + 0.041556f * z;
            a[2 * (size_t) pool.capacity] = 0.055648f * x + -0.204043f * y + 1.057311f * z;
            a[3 * (size_t) pool.capacity] = 1.f;
        }
    }
    if (exchange) {
        s_rec[threadIdx.x] = rec; s_py[threadIdx.x] = py; s_r[threadIdx.x] = r;
        __syncthreads();
        // output thread j -> the thread that computed the j-th record in (group, sample, row-major pixel) order.  The 8
        // warps of the block hold, per group, `pb` pixels (an aligned Morton block: 8x4, 4x4 or 4x2) times `sw` samples.
        const uint32_t sw = bp.sw_log2, pw = 5u - sw, pb_log2 = min(5u, 3u + pw), bw_log2 = pb_log2 == 5u ? 3u : 2u;
        const uint32_t j = threadIdx.x, per_group = 1u << (pb_log2 + sw);
        const uint32_t grp = j / per_group, q = j % per_group, s_local = q >> pb_log2, pm = q & ((1u << pb_log2) - 1u);
        const uint32_t x = pm & ((1u << bw_log2) - 1u), y = pm >> bw_log2;
        const uint32_t m = (x & 1u) | ((y & 1u) << 1) | ((x & 2u) << 1) | ((y & 2u) << 2) | ((x & 4u) << 2);
        const uint32_t wg_log2 = pb_log2 - pw; // warps per group
        const uint32_t src = (((grp << wg_log2) + (m >> pw)) << 5) | (s_local << pw) | (m & ((1u << pw) - 1u));
        rec = s_rec[src]; py = s_py[src]; r = s_r[src];
    }
    if (r != 0xffffffffu) {
        pool.rec[r]    = rec;
        pool.rec_py[r] = py;
    }
}

// AOVIntegrator::sample (aov.cpp:87-122): geometric outputs of the primary hit, read from the hit records of
// bounce 0 (queue index == sample index there) before the next bounce overwrites them.
__global__ void __launch_bounds__(256) k_aov_capture(const __grid_constant__ DScene sc, Pool pool, BatchParams bp, AovPlan plan) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = bp.npix * bp.ns;
    if (i >= n) return;
    const float4 hit = pool.hit[i];
    const uint32_t geom = pool.hit_geom[i];
    Surface sf;
    sf.p = v3(0, 0, 0); sf.n = v3(0, 0, 0); sf.sh.n = v3(0, 0, 0); sf.uvx = 0.f; sf.uvy = 0.f;
    const bool valid = geom != 0xffffffffu;
    if (valid) sf = make_surface(sc, sc.meshes[geom], __float_as_uint(hit.w), hit.y, hit.z);
    uint32_t pixel, s_off;
    slot_decode(bp, i, pixel, s_off);
    float *a = pool.aov + s_off * bp.npix + pixel; // by record index, like k_film_records
    const size_t cs = pool.capacity;
    uint32_t c = 0;
    for (uint32_t k = 0; k < plan.ntypes; ++k) {
        switch (plan.types[k]) {
            case MSK_AOV_DEPTH: a[c * cs] = valid ? hit.x : 0.f; c += 1; break;
            case MSK_AOV_POSITION: a[c * cs] = sf.p.x; a[(c + 1) * cs] = sf.p.y; a[(c + 2) * cs] = sf.p.z; c += 3; break;
            case MSK_AOV_UV: a[c * cs] = sf.uvx; a[(c + 1) * cs] = sf.uvy; c += 2; break;
            case MSK_AOV_GEO_NORMAL: a[c * cs] = sf.n.x; a[(c + 1) * cs] = sf.n.y; a[(c + 2) * cs] = sf.n.z; c += 3; break;
            case MSK_AOV_SH_NORMAL: a[c * cs] = sf.sh.n.x; a[(c + 1) * cs] = sf.sh.n.y; a[(c + 2) * cs] = sf.sh.n.z; c += 3; break;
            default: c += 4; break; // MSK_AOV_INTEGRATOR_RGBA: written by k_film_records once the paths are complete
        }
    }
}

// ImageBlock::put (imageblock.cpp:55-114) + Film::put, as a gather: pixel (x, y) sums
// w_x * w_y * {X,Y,Z,1,1} over the samples of the 5x5 pixel neighbourhood.  A sample outside
// the filter support looks up table[32] == 0 and adds nothing, exactly like the reference's
// lo/hi clipping (rfilter.h:13-16).
constexpr int kFilmTileX = 32, kFilmTileY = 8;
constexpr int kBlockSize = 32; // MSK_BLOCK_SIZE, imageblock.h:8
__global__ void __launch_bounds__(kFilmTileX *kFilmTileY) k_film_gather(const __grid_constant__ DScene sc, Pool pool,
                                                                         BatchParams bp, float *__restrict__ film,
                                                                         uint32_t height, uint32_t stride) {
    const int x = blockIdx.x * kFilmTileX + threadIdx.x, y = blockIdx.y * kFilmTileY + threadIdx.y;
    const int W = (int) bp.width, H = (int) height;
    if (x >= W || y >= H) return;
    const int r = (int) ceilf(sc.cam.filter_radius - 0.5f); // border size, rfilter.cpp:22
    const float scale = sc.cam.filter_scale, radius = sc.cam.filter_radius;
    float aX = 0.f, aY = 0.f, aZ = 0.f, aW = 0.f;
    for (uint32_t s = 0; s < bp.ns; ++s) {
        const size_t sbase = (size_t) s * bp.npix;
        for (int ny = max(y - r, 0); ny <= min(y + r, H - 1); ++ny)
            for (int nx = max(x - r, 0); nx <= min(x + r, W - 1); ++nx) {
                size_t i = sbase + (size_t) ny * W + nx;
                float4 rec = __ldg(pool.rec + i);
                float py = __ldg(pool.rec_py + i);
                // The sample was splatted into the 32x32 ImageBlock of its own pixel (nx, ny), in
                // block-relative float coordinates (imageblock.cpp:86-98): reproduce that rounding,
                // because the 33-entry weight table is a step function of |x - pos|.
                const int bx = (nx & ~(kBlockSize - 1)) - r, by = (ny & ~(kBlockSize - 1)) - r; // m_offset - m_border_size
                const float posx = rec.w - 0.5f - (float) bx, posy = py - 0.5f - (float) by;
                const float xb = (float) (x - bx), yb = (float) (y - by);
                // lo = ceil(pos - radius) <= x <= floor(pos + radius) = hi
                if (xb < posx - radius || xb > posx + radius || yb < posy - radius || yb > posy + radius) continue;
                float wx = __ldg(sc.filter_table + min((int) fabsf((xb - posx) * scale), 32));
                float wy = __ldg(sc.filter_table + min((int) fabsf((yb - posy) * scale), 32));
                float w = wx * wy;
                aX += w * rec.x; aY += w * rec.y; aZ += w * rec.z; aW += w;
            }
    }
    float *p = film + ((size_t) y * W + x) * stride;
    p[0] += aX; p[1] += aY; p[2] += aZ; p[3] += aW; p[4] += aW;
}

// The AOV channels of the film (channels 5.. of every pixel): the same gather, same weights, over the channel-major
// per-sample AOV values.  Not a hot path (the AOV integrator renders one bounce), so it reads through L1/L2.
constexpr int kAovChunk = 8;
__global__ void __launch_bounds__(kFilmTileX *kFilmTileY) k_film_gather_aov(const __grid_constant__ DScene sc, Pool pool,
                                                                             BatchParams bp, float *__restrict__ film,
                                                                             uint32_t height, uint32_t stride, uint32_t nch) {
    const int x = blockIdx.x * kFilmTileX + threadIdx.x, y = blockIdx.y * kFilmTileY + threadIdx.y;
    const int W = (int) bp.width, H = (int) height;
    if (x >= W || y >= H) return;
    const int r = (int) ceilf(sc.cam.filter_radius - 0.5f);
    const float scale = sc.cam.filter_scale, radius = sc.cam.filter_radius;
    for (uint32_t c0 = 0; c0 < nch; c0 += kAovChunk) {
        float acc[kAovChunk];
#pragma unroll
        for (int k = 0; k < kAovChunk; ++k) acc[k] = 0.f;
        for (uint32_t s = 0; s < bp.ns; ++s) {
            const size_t sbase = (size_t) s * bp.npix;
            for (int ny = max(y - r, 0); ny <= min(y + r, H - 1); ++ny)
                for (int nx = max(x - r, 0); nx <= min(x + r, W - 1); ++nx) {
                    const size_t i = sbase + (size_t) ny * W + nx;
                    const float px = __ldg(&pool.rec[i].w), py = __ldg(pool.rec_py + i);
                    const int bx = (nx & ~(kBlockSize - 1)) - r, by = (ny & ~(kBlockSize - 1)) - r;
                    const float posx = px - 0.5f - (float) bx, posy = py - 0.5f - (float) by;
                    const float xb = (float) (x - bx), yb = (float) (y - by);
                    if (xb < posx - radius || xb > posx + radius || yb < posy - radius || yb > posy + radius) continue;
                    const float wx = __ldg(sc.filter_table + min((int) fabsf((xb - posx) * scale), 32));
                    const float wy = __ldg(sc.filter_table + min((int) fabsf((yb - posy) * scale), 32));
                    const float w = wx * wy;
#pragma unroll
                    for (int k = 0; k < kAovChunk; ++k)
                        if (c0 + k < nch) acc[k] += w * __ldg(pool.aov + (size_t) (c0 + k) * pool.capacity + i);
                }
        }
        float *p = film + ((size_t) y * W + x) * stride + 5 + c0;
#pragma unroll
        for (int k = 0; k < kAovChunk; ++k)
            if (c0 + k < nch) p[k] += acc[k];
    }
}

// The same gather with the records of the tile's neighbourhood staged in shared memory, for filters whose border
// (ceil(radius - 0.5), rfilter.cpp:22) is at most kFilmMaxBorder pixels -- the reference's only filter, gaussian with
// radius 2, has border 2.  The global-memory version above reads every record 25 times through L1/L2 (1.3 ms of
// the 18 ms C2 step); here each record is read once per tile (+ halo) and the 33-entry weight table sits in shared
// memory too.  Same neighbour order (s, ny, nx), same float arithmetic: the film is bit-identical to the
// global-memory gather.
constexpr int kFilmMaxBorder = 2;
constexpr int kFilmHaloX = kFilmTileX + 2 * kFilmMaxBorder, kFilmHaloY = kFilmTileY + 2 * kFilmMaxBorder;
__global__ void __launch_bounds__(kFilmTileX *kFilmTileY) k_film_gather_tiled(const __grid_constant__ DScene sc, Pool pool,
                                                                               BatchParams bp, float *__restrict__ film,
                                                                               uint32_t height, uint32_t stride) {
    // Per halo cell (one sample of one pixel): X, Y, Z and the cell's filter weights towards the five pixel columns and the
    // five pixel rows it can reach, computed ONCE by the thread that stages the cell -- the reference's own arithmetic per
    // (sample, target pixel): block-relative positions (imageblock.cpp:86-98), the lo / hi range test, the 33-entry table --
    // with 0 outside the range (adding 0 * value is what skipping the pixel is).  The first version evaluated both weights
    // in the consumer, per (pixel, candidate): 25 candidates x (two range tests, two table look-ups) per pixel and sample,
    // 870 thread instructions, the kernel 67 % issue-bound (profiles/ncu_counters.json: 588 M warp instructions per C2
    // step); staging computes 10 weights per cell instead of 50 per pixel, the consumer is 3 shared loads and 5 FMAs per
    // candidate.  Same neighbour order, same products and sums: bit-identical film.
    constexpr int kWin = 2 * kFilmMaxBorder + 1;
    __shared__ float4 s_xyz[kFilmHaloY][kFilmHaloX];
    __shared__ float s_wx[kWin][kFilmHaloY][kFilmHaloX], s_wy[kWin][kFilmHaloY][kFilmHaloX];
    __shared__ float s_tab[33];
    const int W = (int) bp.width, H = (int) height;
    const int x0 = blockIdx.x * kFilmTileX, y0 = blockIdx.y * kFilmTileY;
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const int r = (int) ceilf(sc.cam.filter_radius - 0.5f);
    const float scale = sc.cam.filter_scale, radius = sc.cam.filter_radius;
    const int tid = threadIdx.y * kFilmTileX + threadIdx.x;
    if (tid < 33) s_tab[tid] = __ldg(sc.filter_table + tid);
    const bool inside = x < W && y < H;
    // this thread's neighbour window in halo coordinates (clipped to the film like the reference's lo/hi clamp)
    const int hx_lo = max(x - r, 0) - (x0 - kFilmMaxBorder), hx_hi = min(x + r, W - 1) - (x0 - kFilmMaxBorder);
    const int hy_lo = max(y - r, 0) - (y0 - kFilmMaxBorder), hy_hi = min(y + r, H - 1) - (y0 - kFilmMaxBorder);
    const int hx_first = (int) threadIdx.x, hy_first = (int) threadIdx.y; // halo coordinate of (x - kFilmMaxBorder, y - kFilmMaxBorder)
    float aX = 0.f, aY = 0.f, aZ = 0.f, aW = 0.f;
    for (uint32_t s = 0; s < bp.ns; ++s) {
        const size_t sbase = (size_t) s * bp.npix;
        __syncthreads();
        for (int e = tid; e < kFilmHaloX * kFilmHaloY; e += kFilmTileX * kFilmTileY) {
            const int hy = e / kFilmHaloX, hx = e - hy * kFilmHaloX;
            const int nx = x0 - kFilmMaxBorder + hx, ny = y0 - kFilmMaxBorder + hy;
            if (nx >= 0 && nx < W && ny >= 0 && ny < H) {
                const size_t i = sbase + (size_t) ny * W + nx;
                const float4 rec = __ldcs(pool.rec + i);
                const float py = __ldcs(pool.rec_py + i);
                const int bx = (nx & ~(kBlockSize - 1)) - r, by = (ny & ~(kBlockSize - 1)) - r; // m_offset - m_border_size
                const float posx = rec.w - 0.5f - (float) bx, posy = py - 0.5f - (float) by;
                s_xyz[hy][hx] = make_float4(rec.x, rec.y, rec.z, 0.f);
#pragma unroll
                for (int d = 0; d < kWin; ++d) { // target pixel (nx - border + d, .) / (., ny - border + d), relative to THIS sample's block
                    const float xb = (float) (nx - kFilmMaxBorder + d - bx), yb = (float) (ny - kFilmMaxBorder + d - by);
                    // lo = ceil(pos - radius) <= x <= floor(pos + radius) = hi
                    const bool outx = xb < posx - radius || xb > posx + radius, outy = yb < posy - radius || yb > posy + radius;
                    s_wx[d][hy][hx] = outx ? 0.f : s_tab[min((int) fabsf((xb - posx) * scale), 32)];
                    s_wy[d][hy][hx] = outy ? 0.f : s_tab[min((int) fabsf((yb - posy) * scale), 32)];
                }
            }
        }
        __syncthreads();
        if (!inside) continue;
#pragma unroll
        for (int dy = 0; dy < kWin; ++dy) {
            const int hy = hy_first + dy;
            if (hy < hy_lo || hy > hy_hi) continue;
#pragma unroll
            for (int dx = 0; dx < kWin; ++dx) {
                const int hx = hx_first + dx;
                if (hx < hx_lo || hx > hx_hi) continue;
                // this pixel is target kWin - 1 - dx of the cell dx columns into its window (x = nx - border + (kWin - 1 - dx))
                const float w = s_wx[kWin - 1 - dx][hy][hx] * s_wy[kWin - 1 - dy][hy][hx];
                const float4 a = s_xyz[hy][hx];
                aX += w * a.x; aY += w * a.y; aZ += w * a.z; aW += w;
            }
        }
    }
    if (!inside) return;
    float *p = film + ((size_t) y * W + x) * stride;
    p[0] += aX; p[1] += aY; p[2] += aZ; p[3] += aW; p[4] += aW;
}

// ---------------------------------------------------------------------------------------
// Stand-alone batch queries (msk_gpu_intersect / msk_gpu_occluded)
template <bool STATS>
struct QueryClosestIO {
    const MskRay *rays;
    MskHit *hits;
    uint32_t *nnodes, *ntris;
    __device__ __forceinline__ void load(uint32_t q, float4 &ro, float4 &rd) const {
        const float4 *rp = reinterpret_cast<const float4 *>(rays + q);
        ro = __ldcs(rp); rd = __ldcs(rp + 1);
    }
    __device__ __forceinline__ uint32_t tag(float4 &) const { return 0u; }
    __device__ __forceinline__ void commit(bool have, uint32_t q, const Traversal &s) {
        if (!have) return;
        if (STATS) { nnodes[q] = s.cnt_nodes; ntris[q] = s.cnt_tris; }
        else {
            MskHit o;
            const bool found = s.is_hit();
            o.t = found ? s.hit.t : MSK_INF; o.u = found ? s.hit.u : 0.f; o.v = found ? s.hit.v : 0.f;
            o.prim = found ? s.hit.prim : 0xffffffffu; o.geom = found ? s.hit.geom : 0xffffffffu;
            hits[q] = o;
        }
    }
};

template <bool STATS>
__global__ void __launch_bounds__(128, MSK_TRAV_MIN_BLOCKS) k_query_closest(const __grid_constant__ DScene sc, const MskRay *__restrict__ rays,
                                                       MskHit *__restrict__ hits, uint32_t n, uint32_t *cursor,
                                                       uint32_t *nnodes, uint32_t *ntris) {
    MSK_TRAV_SHARED;
    const Accel ac{ sc.nodes, sc.tris, sc.k47, perm_lut_init(msk_s_perm) };
    QueryClosestIO<STATS> io{ rays, hits, nnodes, ntris };
    trace_queue<false, STATS>(ac, MSK_TRAV_SMEM, n, cursor, io, 0);
}

struct QueryAnyIO {
    const MskRay *rays;
    uint8_t *occ;
    __device__ __forceinline__ void load(uint32_t q, float4 &ro, float4 &rd) const {
        const float4 *rp = reinterpret_cast<const float4 *>(rays + q);
        ro = __ldcs(rp); rd = __ldcs(rp + 1);
    }
    __device__ __forceinline__ uint32_t tag(float4 &) const { return 0u; }
    __device__ __forceinline__ void commit(bool have, uint32_t q, const Traversal &s) {
        if (have) occ[q] = s.is_hit() ? 1 : 0;
    }
};

__global__ void __launch_bounds__(128, MSK_TRAV_MIN_BLOCKS) k_query_any(const __grid_constant__ DScene sc, const MskRay *__restrict__ rays,
                                                   uint8_t *__restrict__ occ, uint32_t n, uint32_t *cursor) {
    MSK_TRAV_SHARED;
    const Accel ac{ sc.nodes, sc.tris, sc.k47, perm_lut_init(msk_s_perm) };
    QueryAnyIO io{ rays, occ };
    trace_queue<true, false>(ac, MSK_TRAV_SMEM, n, cursor, io, 0);
}

template <typename T> cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **) p, std::max<size_t>(n, 1) * sizeof(T)); }

} // namespace

// =========================================================================================
// One in-flight batch: its own path pool, stream and poll buffers (see Renderer::render).
constexpr int kMaxLanes = 4;
struct Lane {
    Pool pool{};
    cudaStream_t stream = nullptr;            // lane 0 runs on the caller's stream; the others own one
    Ctrl *h_poll[2] = { nullptr, nullptr };   // pinned: queue-length polls of unbounded-depth jobs, double-buffered
    Ctrl *h_ctrl = nullptr;                   // pinned: the lane's counters at the end of a job
    cudaEvent_t poll_ev[2]{};
    cudaEvent_t film_done = nullptr;          // the lane's last film kernel of a batch
    size_t aov_floats = 0;                    // allocated size of pool.aov
};

struct Renderer::Impl {
    Lane lanes[kMaxLanes];
    int npools = 0;        // lanes whose pool is allocated (all with `capacity` paths)
    uint32_t capacity = 0;
    uint32_t *query_cursor = nullptr;
    cudaEvent_t ev[8]{};
    cudaEvent_t fork_ev = nullptr;
    // Tuning knobs (environment, read once in init(); defaults are the measured best on C2, tools/ab_knobs.sh)
    uint32_t batch_paths = 32u << 20; // MSK_BATCH_PATHS: paths per wavefront batch (C2: 8 Mi -> 16 Mi = 804 -> 919 Mpaths/s, the per-batch tail of short bounces is paid once)
    int inflight = 2;                 // MSK_INFLIGHT: batches in flight, each on its own stream and pool (1: one batch after the other)
    uint32_t split_min = 4u << 20;    // MSK_SPLIT_MIN_PATHS: a job of fewer batches than lanes is split further while a batch keeps this many paths
    int spec_shade = 1;               // MSK_SPEC_SHADE: one k_shade launch per material key present in the scene
    uint32_t spec_min = 1u << 18;     // MSK_SPEC_MIN: ... while the queue may hold at least this many vertices
    int shadow_static_bounces = 0;    // MSK_SHADOW_STATIC_BOUNCES: bounces whose shadow queue counts as coherent
    int static_bounces = 1;           // MSK_STATIC_BOUNCES: bounces whose closest-hit queue counts as coherent (camera rays)
    int packet_camera = 1;            // MSK_PACKET_CAMERA: the warps of the camera-ray queue (samples of one pixel group) are traversed as packets
                                      // (C2 1549 -> 1570 Mpaths/s, C3 946 -> 955, bit-identical films; not for tiny scenes: C4 1802 -> 1797.  Bounce-0
                                      // shadow rays as packets -- shafts from a pixel's footprint to the light -- were 3x slower, profiles/r03g_ab_packets.txt)
    int debug_bounces = 0;            // MSK_DEBUG_BOUNCES: print polled queue lengths and per-launch stage times to stderr
    uint32_t tail_threshold = 1u << 18; // MSK_TAIL_THRESHOLD: finish an unbounded job with k_tail once the queue is this short (0: never)
    int tiled_slots = 1;              // MSK_TILED_SLOTS: enumerate the film in 8x4 tiles (see slot_decode)
    int samples_per_warp = 32;        // MSK_SAMPLES_PER_WARP: upper bound of the samples of one pixel group a warp holds (0: sample-major order)
    int poll_min_depth = 8;           // MSK_POLL_MIN_DEPTH: jobs with max_depth >= this (or unbounded) poll the queue length from bounce 4 on
    int async_poll = 1;               // MSK_ASYNC_POLL: poll the queue length one bounce late, without draining the stream
    uint32_t static_nodes = 64;       // MSK_STATIC_NODES: scenes with at most this many wide nodes always use the static traversal (every ray does the same few steps: nothing to re-balance)
    int first_elide = 1;              // MSK_FIRST_ELIDE: camera paths' constant initial state is neither written by k_raygen nor read by the first k_shade
    int use_graph = 1;                // MSK_GRAPH: bounded-depth jobs (a fixed launch sequence) replay a cached CUDA graph
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<unsigned char> graph_key; // everything the captured launches depend on
    uint64_t graph_launches = 0, graph_extra_closest = 0; uint32_t graph_bounces = 0, graph_batches = 0;
    int ray_sort = 0;                 // MSK_RAY_SORT (experiment): 1 = reorder the closest-hit queue of bounces >= 1 by (origin cell, octant)
    void *rs_tmp = nullptr; size_t rs_tmp_bytes = 0;
    std::vector<cudaEvent_t> timer_events; // MSK_RENDER_STAGE_TIMERS
    std::vector<int> timer_stage;
    int persistent_blocks = 0, sm_count = 0;

    void free_pools() {
        for (int l = 0; l < kMaxLanes; ++l) {
            Pool &p = lanes[l].pool;
            for (int i = 0; i < 2; ++i) { cudaFree(p.rays[i]); cudaFree(p.T[i]); cudaFree(p.WL[i]); cudaFree(p.AUX[i]); cudaFree(p.MISC[i]); }
            cudaFree(p.hit); cudaFree(p.hit_geom); cudaFree(p.L); cudaFree(p.sh_ray); cudaFree(p.sh_contrib);
            cudaFree(p.sorted); cudaFree(p.rec); cudaFree(p.rec_py); cudaFree(p.ctrl); cudaFree(p.aov);
            for (int i = 0; i < 2; ++i) { cudaFree(p.rs_keys[i]); cudaFree(p.rs_vals[i]); }
            p = Pool{};
            lanes[l].aov_floats = 0;
        }
        cudaFree(rs_tmp); rs_tmp = nullptr; rs_tmp_bytes = 0;
        npools = 0; capacity = 0;
    }
    void drop_graph() {
        if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; graph_key.clear(); }
    }
    // samples per pixel and batch, lanes in flight for a job of `nsamples` samples per pixel over `npix` pixels
    void plan(uint32_t npix, uint32_t nsamples, uint32_t paths_per_batch, bool single_lane, uint32_t *per_batch, int *nlanes) const {
        const uint32_t target = paths_per_batch ? paths_per_batch : batch_paths;
        uint32_t pb = std::max(1u, target / npix);
        int nl = single_lane ? 1 : std::min(std::max(inflight, 1), kMaxLanes);
        nsamples = std::max(nsamples, 1u);
        if (nl > 1 && !paths_per_batch) { // fewer batches than lanes: split while a batch stays large enough to fill the machine
            const uint32_t want = (nsamples + (uint32_t) nl - 1u) / (uint32_t) nl;
            const uint32_t floor_ = std::max(1u, split_min / npix);
            if (want < pb) pb = std::max(want, std::min(floor_, pb));
        }
        // a warp holds up to 32 samples of a pixel (slot_decode): keep the batch a multiple of 32, else a power of two
        if (pb >= 32u) pb &= ~31u;
        else { uint32_t q = 1; while (q * 2u <= pb) q *= 2u; pb = q; }
        pb = std::min(pb, nsamples);
        const uint32_t nbatches = (nsamples + pb - 1u) / pb;
        *per_batch = pb;
        *nlanes = (int) std::min<uint32_t>((uint32_t) nl, nbatches);
    }
    // What a job runs as (shared by reserve() and render(), which must agree on the pools).  A bounded-depth job below the
    // polling depth is a FIXED sequence of launches (queue lengths live on the device): C1 is 34 launches of 5-90 us
    // each, and the host-side launch cost and the gaps between them were ~7 % of its 1 ms step.  The sequence is captured
    // once into a CUDA graph and replayed while nothing it depends on changes (scene pointers and parameters, the render
    // description, the film pointer, the pool); anything else runs launch by launch.
    // One lane: profiling modes (stage timers and traversal counters describe one batch at a time), the AOV integrator
    // (one bounce), graph replay, the ray-sort experiment (one scratch buffer).
    void job_plan(uint32_t npix, const MskRenderDesc &rd, bool has_aov, uint32_t bound, uint32_t *per_batch, int *nlanes, bool *graphable) const {
        const uint32_t nsamples = rd.sample_end - rd.sample_begin;
        const bool timers = (rd.flags & MSK_RENDER_STAGE_TIMERS) != 0, tstats = (rd.flags & MSK_RENDER_TRAVERSAL_STATS) != 0;
        plan(npix, nsamples, rd.paths_per_batch, true, per_batch, nlanes);
        const uint64_t nbatches1 = nsamples ? ((uint64_t) nsamples + *per_batch - 1) / *per_batch : 0;
        *graphable = use_graph && !timers && !tstats && !debug_bounces && bound < (uint32_t) poll_min_depth && nbatches1 * (6ull * bound + 8) <= 1024;
        const bool single = timers || tstats || debug_bounces || ray_sort || has_aov || *graphable || bound == 0;
        if (!single) plan(npix, nsamples, rd.paths_per_batch, false, per_batch, nlanes);
    }
};

Renderer::Renderer() : impl_(new Impl) {}
Renderer::~Renderer() { release(); delete impl_; }

void Renderer::release() {
    impl_->drop_graph();
    impl_->free_pools();
    cudaFree(impl_->query_cursor); impl_->query_cursor = nullptr;
    for (int l = 0; l < kMaxLanes; ++l) {
        Lane &ln = impl_->lanes[l];
        if (l > 0 && ln.stream) { cudaStreamDestroy(ln.stream); }
        ln.stream = nullptr;
        if (ln.h_ctrl) { cudaFreeHost(ln.h_ctrl); ln.h_ctrl = nullptr; }
        for (auto &h : ln.h_poll) if (h) { cudaFreeHost(h); h = nullptr; }
        for (auto &e : ln.poll_ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (ln.film_done) { cudaEventDestroy(ln.film_done); ln.film_done = nullptr; }
    }
    for (auto &e : impl_->ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (impl_->fork_ev) { cudaEventDestroy(impl_->fork_ev); impl_->fork_ev = nullptr; }
    for (auto &e : impl_->timer_events) cudaEventDestroy(e);
    impl_->timer_events.clear();
}

int Renderer::init(int sm_count) {
    impl_->sm_count = sm_count;
    impl_->persistent_blocks = sm_count * 8; // 128-thread CTAs, 8 resident per SM
    MSK_CUDA_CHECK(dalloc(&impl_->query_cursor, 1));
    for (int l = 0; l < kMaxLanes; ++l) {
        Lane &ln = impl_->lanes[l];
        if (l > 0) MSK_CUDA_CHECK(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        MSK_CUDA_CHECK(cudaMallocHost((void **) &ln.h_ctrl, sizeof(Ctrl)));
        for (auto &h : ln.h_poll) MSK_CUDA_CHECK(cudaMallocHost((void **) &h, sizeof(Ctrl)));
        for (auto &e : ln.poll_ev) MSK_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        MSK_CUDA_CHECK(cudaEventCreateWithFlags(&ln.film_done, cudaEventDisableTiming));
    }
    for (auto &e : impl_->ev) MSK_CUDA_CHECK(cudaEventCreate(&e));
    MSK_CUDA_CHECK(cudaEventCreateWithFlags(&impl_->fork_ev, cudaEventDisableTiming));
    auto env_u = [](const char *name, long long dflt) -> long long {
        const char *v = getenv(name);
        return (v && *v) ? atoll(v) : dflt;
    };
    impl_->batch_paths = (uint32_t) std::max<long long>(1, env_u("MSK_BATCH_PATHS", impl_->batch_paths));
    impl_->inflight = (int) env_u("MSK_INFLIGHT", impl_->inflight);
    impl_->split_min = (uint32_t) std::max<long long>(1, env_u("MSK_SPLIT_MIN_PATHS", impl_->split_min));
    impl_->spec_shade = (int) env_u("MSK_SPEC_SHADE", impl_->spec_shade);
    impl_->spec_min = (uint32_t) env_u("MSK_SPEC_MIN", impl_->spec_min);
    impl_->shadow_static_bounces = (int) env_u("MSK_SHADOW_STATIC_BOUNCES", impl_->shadow_static_bounces);
    impl_->static_bounces = (int) env_u("MSK_STATIC_BOUNCES", impl_->static_bounces);
    impl_->packet_camera = (int) env_u("MSK_PACKET_CAMERA", impl_->packet_camera);
    impl_->async_poll = (int) env_u("MSK_ASYNC_POLL", impl_->async_poll);
    impl_->tail_threshold = (uint32_t) env_u("MSK_TAIL_THRESHOLD", impl_->tail_threshold);
    impl_->poll_min_depth = (int) env_u("MSK_POLL_MIN_DEPTH", impl_->poll_min_depth);
    impl_->tiled_slots = (int) env_u("MSK_TILED_SLOTS", impl_->tiled_slots);
    impl_->samples_per_warp = (int) env_u("MSK_SAMPLES_PER_WARP", impl_->samples_per_warp);
    impl_->ray_sort = (int) env_u("MSK_RAY_SORT", impl_->ray_sort);
    impl_->use_graph = (int) env_u("MSK_GRAPH", impl_->use_graph);
    impl_->first_elide = (int) env_u("MSK_FIRST_ELIDE", impl_->first_elide);
    impl_->static_nodes = (uint32_t) env_u("MSK_STATIC_NODES", impl_->static_nodes);
    impl_->debug_bounces = (int) env_u("MSK_DEBUG_BOUNCES", 0);
    // the packet kernel is not reached by the warm-up render of msk_gpu_init (a tiny scene): load it now, not in the first render
    cudaFuncAttributes fa;
    MSK_CUDA_CHECK(cudaFuncGetAttributes(&fa, k_intersect<false, true>));
    return MSK_OK;
}

// `count` pools of `capacity` paths each (pools only grow; a larger capacity re-allocates them all)
int Renderer::ensure_pool(uint32_t capacity, int count) {
    Impl &im = *impl_;
    count = std::max(count, 1);
    if (capacity <= im.capacity && count <= im.npools) return MSK_OK;
    im.drop_graph(); // captured the old pool
    if (capacity > im.capacity) { count = std::max(count, im.npools); im.free_pools(); }
    else capacity = im.capacity;
    const size_t n = capacity;
    for (int l = im.npools; l < count; ++l) {
        Pool &p = im.lanes[l].pool;
        for (int i = 0; i < 2; ++i) {
            MSK_CUDA_CHECK(dalloc(&p.rays[i], n)); MSK_CUDA_CHECK(dalloc(&p.T[i], n)); MSK_CUDA_CHECK(dalloc(&p.WL[i], n));
            MSK_CUDA_CHECK(dalloc(&p.AUX[i], n)); MSK_CUDA_CHECK(dalloc(&p.MISC[i], n));
        }
        MSK_CUDA_CHECK(dalloc(&p.hit, n)); MSK_CUDA_CHECK(dalloc(&p.hit_geom, n)); MSK_CUDA_CHECK(dalloc(&p.L, n));
        MSK_CUDA_CHECK(dalloc(&p.sh_ray, n)); MSK_CUDA_CHECK(dalloc(&p.sh_contrib, n));
        MSK_CUDA_CHECK(dalloc(&p.sorted, n * kNumKeys)); MSK_CUDA_CHECK(dalloc(&p.rec, n)); MSK_CUDA_CHECK(dalloc(&p.rec_py, n));
        MSK_CUDA_CHECK(dalloc(&p.ctrl, 1));
        MSK_CUDA_CHECK(cudaMemset(p.ctrl, 0, sizeof(Ctrl)));
        if (im.ray_sort) {
            for (int i = 0; i < 2; ++i) { MSK_CUDA_CHECK(dalloc(&p.rs_keys[i], n)); MSK_CUDA_CHECK(dalloc(&p.rs_vals[i], n)); }
            if (!im.rs_tmp) {
                MSK_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, im.rs_tmp_bytes, p.rs_keys[0], p.rs_keys[1], p.rs_vals[0], p.rs_vals[1], (int) n, 0, 27));
                MSK_CUDA_CHECK(cudaMalloc(&im.rs_tmp, std::max<size_t>(im.rs_tmp_bytes, 16)));
            }
        }
        p.capacity = capacity;
        im.npools = l + 1;
    }
    im.capacity = capacity;
    return MSK_OK;
}

int Renderer::aov_plan(const MskAovDesc &aov, uint32_t *nch) {
    static const uint32_t per[6] = { 1, 3, 2, 3, 3, 4 };
    if (aov.ntypes && !aov.types) return fail(MSK_ERR_ARG, "AOV description without types");
    uint32_t n = 0, nested = 0;
    for (uint32_t i = 0; i < aov.ntypes; ++i) {
        if (aov.types[i] < 0 || aov.types[i] > MSK_AOV_INTEGRATOR_RGBA) return fail(MSK_ERR_ARG, "Invalid AOV type %d!", aov.types[i]);
        nested += aov.types[i] == MSK_AOV_INTEGRATOR_RGBA;
        n += per[aov.types[i]];
    }
    if (nested > 1) return fail(MSK_ERR_UNSUPPORTED, "at most one nested integrator per AOV integrator");
    if (n > MSK_AOV_MAX_CHANNELS) return fail(MSK_ERR_UNSUPPORTED, "more than %d AOV channels", MSK_AOV_MAX_CHANNELS);
    if (nch) *nch = n;
    return MSK_OK;
}

// Allocates the path pools a render of `rd` will use, so that the render itself makes no allocation.  cudaMalloc
// synchronises the whole device; msk_gpu_render_multi calls this for every context before any of them starts, because one
// context's reduction kernel may already be spinning on a flag that another context OF THE SAME DEVICE publishes only after
// its render (two contexts on one GPU: the one-GPU test of the multi-device path).
int Renderer::reserve(const DScene &sc, const MskRenderDesc &rd) {
    const uint64_t npix64 = (uint64_t) sc.cam.width * sc.cam.height;
    if (!npix64 || npix64 > (1ull << 27)) return fail(MSK_ERR_UNSUPPORTED, "film size %ux%u unsupported", sc.cam.width, sc.cam.height);
    if (rd.sample_end < rd.sample_begin || rd.sample_end > rd.spp) return fail(MSK_ERR_ARG, "bad sample range");
    const uint32_t npix = (uint32_t) npix64;
    const uint32_t bound = rd.max_depth > 0 ? (uint32_t) rd.max_depth : (rd.max_depth == 0 ? 0u : 0xffffffffu);
    uint32_t per_batch; int nlanes; bool graphable;
    impl_->job_plan(npix, rd, false, bound, &per_batch, &nlanes, &graphable);
    return ensure_pool(npix * per_batch, nlanes);
}

// The batch loop.  A batch is all pixels x a run of consecutive samples; its bounces are a chain of dependent launches
// whose queues thin out geometrically, so the late bounces of a batch cannot fill the machine (C3: bounces 8-16 run over
// < 0.5 M rays each, 5 launches per bounce; the fog workload has ~60 such bounces per batch).  Up to MSK_INFLIGHT batches
// are therefore IN FLIGHT at once, each on its own stream with its own path pool: the thin tail of one batch runs under
// the fat bounces of the next, which is what a regenerating path pool buys without mixing path depths in one
// queue -- the sorted, depth-uniform wavefront and the per-batch film gather stay as they are.  The film stays
// bit-deterministic: the film kernels of batch b wait (event) for those of batch b - 1, so every pixel accumulates its
// batches in order.
int Renderer::render(cudaStream_t stream0, const DScene &sc, const MskRenderDesc &rd, float *d_film, MskStats *stats, const MskAovDesc *aov) {
    const uint32_t W = sc.cam.width, H = sc.cam.height;
    const uint64_t npix64 = (uint64_t) W * H;
    if (!W || !H || npix64 > (1ull << 27)) return fail(MSK_ERR_UNSUPPORTED, "film size %ux%u unsupported", W, H);
    if (rd.sample_end < rd.sample_begin || rd.sample_end > rd.spp) return fail(MSK_ERR_ARG, "bad sample range");
    if (rd.integrator > MSK_INTEGRATOR_VOLPATH) return fail(MSK_ERR_ARG, "unknown integrator %u", rd.integrator);
    if (rd.rr_depth <= 0) return fail(MSK_ERR_ARG, "\"rr_depth\" must be set to a value greater than zero!");
    if (rd.max_depth < 0 && rd.max_depth != -1) return fail(MSK_ERR_ARG, "\"max_depth\" must be set to -1 (infinite) or a value >= 0");
    Impl &im = *impl_;
    const uint32_t npix = (uint32_t) npix64;
    const uint32_t nsamples = rd.sample_end - rd.sample_begin;
    const bool tstats = (rd.flags & MSK_RENDER_TRAVERSAL_STATS) != 0;
    // MSK_RENDER_STAGE_TIMERS: bracket every launch with a pair of events (a profiling aid used by bench.py
    // for the per-kernel roofline; the extra event records perturb ms_render slightly, so it is off by default)
    const bool timers = (rd.flags & MSK_RENDER_STAGE_TIMERS) != 0;
    // AOV integrator: film stride 5 + nch; per-sample AOV values live beside the path pool
    AovPlan plan{};
    plan.rgba_channel = -1;
    int rc;
    if (aov) {
        if ((rc = aov_plan(*aov, &plan.nch))) return rc;
        plan.ntypes = aov->ntypes;
        static const uint32_t per[6] = { 1, 3, 2, 3, 3, 4 };
        for (uint32_t i = 0, c = 0; i < aov->ntypes; ++i) {
            plan.types[i] = (uint8_t) aov->types[i];
            if (aov->types[i] == MSK_AOV_INTEGRATOR_RGBA) plan.rgba_channel = (int32_t) c;
            c += per[aov->types[i]];
        }
    }
    const uint32_t stride = 5 + plan.nch;
    const bool trace_paths = !aov || plan.rgba_channel >= 0; // an AOV integrator without a nested one traces primary rays only
    // a path reaches vertex `depth` only while depth <= max_depth, and the vertex at max_depth still
    // needs its shade pass (emission), so a bounded job runs exactly max_depth iterations
    const uint32_t bound = !trace_paths ? 0u : (rd.max_depth > 0 ? (uint32_t) rd.max_depth : (rd.max_depth == 0 ? 0u : 0xffffffffu));
    uint32_t per_batch; int nlanes; bool graphable;
    im.job_plan(npix, rd, aov != nullptr, bound, &per_batch, &nlanes, &graphable);
    if ((rc = ensure_pool(npix * per_batch, nlanes))) return rc;
    if (aov) {
        Lane &l0 = im.lanes[0];
        const size_t need = (size_t) im.capacity * std::max(plan.nch, 1u);
        if (l0.aov_floats < need) {
            cudaFree(l0.pool.aov); l0.pool.aov = nullptr; l0.aov_floats = 0;
            MSK_CUDA_CHECK(dalloc(&l0.pool.aov, need));
            l0.aov_floats = need;
        }
    }
    im.lanes[0].stream = stream0;

    uint64_t launches = 0;
    uint32_t max_bounces = 0, batches = 0;
    bool tail_used = false;
    const int pb = im.persistent_blocks;
    enum { ST_RAYGEN, ST_INTERSECT, ST_SORT, ST_SHADE, ST_SHADOW, ST_FILM, ST_TAIL, ST_COUNT };
    uint64_t extra_closest = 0;
    std::vector<int> &tstage = im.timer_stage;
    tstage.clear();
    size_t tev = 0;
    auto stage_begin = [&](int st) -> int { // (timers imply one lane: the caller's stream)
        if (!timers) return MSK_OK;
        if (tev + 2 > im.timer_events.size()) {
            for (int k = 0; k < 2; ++k) { cudaEvent_t e; MSK_CUDA_CHECK(cudaEventCreate(&e)); im.timer_events.push_back(e); }
        }
        tstage.push_back(st);
        MSK_CUDA_CHECK(cudaEventRecord(im.timer_events[tev], stream0));
        return MSK_OK;
    };
    auto stage_end = [&]() -> int {
        if (!timers) return MSK_OK;
        MSK_CUDA_CHECK(cudaEventRecord(im.timer_events[tev + 1], stream0));
        tev += 2;
        return MSK_OK;
    };
#define MSK_STAGE(st, launch) do { int rc__ = stage_begin(st); if (rc__) return rc__; launch; launches++; rc__ = stage_end(); if (rc__) return rc__; } while (0)

    // the path tracer's first shade pass takes the camera paths' initial state as known (k_raygen, k_shade); the volumetric
    // kernel reads it
    const bool first_elide = im.first_elide && rd.integrator != MSK_INTEGRATOR_VOLPATH;
    // state of the batch a lane is working on
    struct Run {
        bool active = false, done = false;
        uint64_t index = 0;      // batch number within the job
        BatchParams bp{};
        uint32_t n = 0, bounce = 0, n_est = 0;
        int cur = 0;
        bool poll = false, use_tail = false, poll_pending = false;
    };
    Run runs[kMaxLanes];

    // raygen of a new batch
    auto start = [&](int li, uint64_t index, uint32_t s0) -> int {
        Lane &ln = im.lanes[li];
        Pool &pool = ln.pool;
        cudaStream_t stream = ln.stream;
        Run &r = runs[li];
        r = Run{};
        r.active = true; r.index = index;
        BatchParams &bp = r.bp;
        bp.npix = npix; bp.width = W; bp.s0 = s0; bp.ns = std::min(per_batch, rd.sample_end - s0);
        bp.spp = rd.spp; bp.base_seed = rd.base_seed;
        bp.max_depth = rd.max_depth; bp.rr_depth = rd.rr_depth; bp.hide_emitters = rd.hide_emitters;
        bp.tiled = (im.tiled_slots && W % 8u == 0 && H % 4u == 0) ? 1u : 0u;
        if (bp.tiled && im.tiled_slots >= 2 && W % 16u == 0 && H % 8u == 0) bp.tiled = 2u;
        bp.sw_log2 = 0;
        if (im.samples_per_warp <= 0) bp.sw_log2 = 0xffffffffu;
        else while (bp.sw_log2 < 5u && (2u << bp.sw_log2) <= (uint32_t) im.samples_per_warp && bp.ns % (2u << bp.sw_log2) == 0u) bp.sw_log2++;
        const uint32_t n = npix * bp.ns;
        r.n = n; r.n_est = n; // n_est: upper bound of the current queue length known to the host (queues only shrink)
        k_begin_batch<<<1, 1, 0, stream>>>(pool.ctrl, n);
        launches++;
        MSK_STAGE(ST_RAYGEN, (k_raygen<<<(n + 255) / 256, 256, 0, stream>>>(sc, pool, bp, first_elide ? 0 : 1)));
        if (aov && bound == 0) { // the AOV integrator's own ray_intersect (aov.cpp:90) when no path bounce runs
            MSK_STAGE(ST_INTERSECT, (k_intersect<false><<<pb, 128, 0, stream>>>(sc, pool, 0, 1, nullptr)));
            MSK_STAGE(ST_FILM, (k_aov_capture<<<(n + 255) / 256, 256, 0, stream>>>(sc, pool, bp, plan)));
            extra_closest += n;
        }
        // (not with MSK_RENDER_TRAVERSAL_STATS: the per-stage node / triangle counters then describe the wavefront kernels alone)
        r.poll = bound >= (uint32_t) im.poll_min_depth; // unbounded jobs, and bounded ones deep enough to have a thin tail
        r.use_tail = im.tail_threshold > 0 && r.poll && !tstats;
        r.done = bound == 0;
        return MSK_OK;
    };

    // one bounce of a lane's batch (or the switch to the tail kernel); sets done once the batch needs no further bounce
    auto step = [&](int li) -> int {
        Lane &ln = im.lanes[li];
        Pool &pool = ln.pool;
        cudaStream_t stream = ln.stream;
        Run &r = runs[li];
        const BatchParams &bp = r.bp;
        const uint32_t n = r.n;
        const int cur = r.cur;
        const uint32_t bounce = r.bounce;
        const uint32_t *perm = nullptr;
        if (im.ray_sort && bounce >= 1 && r.n_est >= (1u << 18)) { // the reordering pass is timed with the material sort
            MSK_STAGE(ST_SORT, (k_ray_keys<<<(r.n_est + 255) / 256, 256, 0, stream>>>(sc, pool, cur, r.n_est, pool.rs_keys[0], pool.rs_vals[0])));
            size_t tmp_bytes = im.rs_tmp_bytes;
            if (stage_begin(ST_SORT)) return MSK_ERR_CUDA;
            MSK_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(im.rs_tmp, tmp_bytes, pool.rs_keys[0], pool.rs_keys[1], pool.rs_vals[0], pool.rs_vals[1], (int) r.n_est, 0, 27, stream));
            if (stage_end()) return MSK_ERR_CUDA;
            perm = pool.rs_vals[1];
        }
        const int tiny_scene = sc.nnodes <= im.static_nodes;
        // 0: lockstep driver, 1: static warps of 32 consecutive rays, 2: static warps traversed as packets (camera rays)
        const int isect_coherent = bounce == 0 && im.static_bounces > 0 && im.packet_camera && !tiny_scene && !perm ? 2 : ((int) bounce < im.static_bounces || tiny_scene);
        const int first = first_elide && bounce == 0;
        if (isect_coherent == 2 && tstats) MSK_STAGE(ST_INTERSECT, (k_intersect<true, true><<<pb, 128, 0, stream>>>(sc, pool, cur, isect_coherent, perm)));
        else if (isect_coherent == 2) MSK_STAGE(ST_INTERSECT, (k_intersect<false, true><<<pb, 128, 0, stream>>>(sc, pool, cur, isect_coherent, perm)));
        else if (tstats) MSK_STAGE(ST_INTERSECT, (k_intersect<true><<<pb, 128, 0, stream>>>(sc, pool, cur, isect_coherent, perm)));
        else MSK_STAGE(ST_INTERSECT, (k_intersect<false><<<pb, 128, 0, stream>>>(sc, pool, cur, isect_coherent, perm)));
        // the AOV integrator shares the primary hit with the nested path tracer (the reference intersects twice)
        if (aov && bounce == 0) MSK_STAGE(ST_FILM, (k_aov_capture<<<(n + 255) / 256, 256, 0, stream>>>(sc, pool, bp, plan)));
        if (!MSK_SORT_IN_COMMIT) MSK_STAGE(ST_SORT, (k_sort<<<im.sm_count * 4, kSortThreads, 0, stream>>>(sc, pool, cur)));
        if (rd.integrator == MSK_INTEGRATOR_VOLPATH) {
            MSK_STAGE(ST_SHADE, (k_shade_vol<<<pb, 128, 0, stream>>>(sc, pool, bp, cur)));
        } else if (im.spec_shade && r.n_est >= im.spec_min) {
            const uint32_t keys = 1u | (sc.bsdf_type_mask << 1);
#define MSK_SHADE_KEY(K) if (keys & (1u << K)) MSK_STAGE(ST_SHADE, (k_shade<K><<<pb, 128, 0, stream>>>(sc, pool, bp, cur, first)))
            MSK_SHADE_KEY(0); MSK_SHADE_KEY(1); MSK_SHADE_KEY(2); MSK_SHADE_KEY(3); MSK_SHADE_KEY(4); MSK_SHADE_KEY(5);
#undef MSK_SHADE_KEY
            static_assert(kNumKeys == 6, "one specialised k_shade launch per key");
        } else {
            MSK_STAGE(ST_SHADE, (k_shade<-1><<<pb, 128, 0, stream>>>(sc, pool, bp, cur, first)));
        }
        const int sh_coherent = (int) bounce < im.shadow_static_bounces || tiny_scene;
        if (tstats) MSK_STAGE(ST_SHADOW, (k_shadow<true><<<pb, 128, 0, stream>>>(sc, pool, sh_coherent)));
        else MSK_STAGE(ST_SHADOW, (k_shadow<false><<<pb, 128, 0, stream>>>(sc, pool, sh_coherent)));
        k_end_bounce<<<1, 1, 0, stream>>>(pool.ctrl, cur);
        launches++;
        r.cur ^= 1;
        r.bounce++;
        if (r.bounce >= bound) { r.done = true; return MSK_OK; }
        // unbounded paths (Russian roulette only): poll the queue length once it is likely short.  The poll of
        // bounce b is read after bounce b+1 has been enqueued, so the stream never drains; the price is one
        // bounce over an empty queue at the very end.
        if (r.poll && r.bounce >= 4) {
            const int slot = (int) (r.bounce & 1u);
            MSK_CUDA_CHECK(cudaMemcpyAsync(ln.h_poll[slot], pool.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
            MSK_CUDA_CHECK(cudaEventRecord(ln.poll_ev[slot], stream));
            if (im.async_poll) {
                if (r.poll_pending) {
                    MSK_CUDA_CHECK(cudaEventSynchronize(ln.poll_ev[slot ^ 1]));
                    r.n_est = ln.h_poll[slot ^ 1]->n_rays[r.cur ^ 1]; // queue the bounce just enqueued ran over
                    if (im.debug_bounces) fprintf(stderr, "[msk] bounce %u ran over %u rays\n", r.bounce - 1, r.n_est);
                    if (r.n_est == 0) { r.done = true; return MSK_OK; }
                }
                r.poll_pending = true;
            } else {
                MSK_CUDA_CHECK(cudaEventSynchronize(ln.poll_ev[slot]));
                r.n_est = ln.h_poll[slot]->n_rays[r.cur];
                if (r.n_est == 0) { r.done = true; return MSK_OK; }
            }
            // queues only shrink: once the last polled length is below the threshold, one k_tail launch runs
            // every surviving path of the current queue to completion instead of ~5 launches per further bounce
            if (r.use_tail && r.n_est <= im.tail_threshold) {
                if (rd.integrator == MSK_INTEGRATOR_VOLPATH) MSK_STAGE(ST_TAIL, (k_tail<false, true><<<pb, 128, 0, stream>>>(sc, pool, bp, r.cur)));
                else MSK_STAGE(ST_TAIL, (k_tail<false, false><<<pb, 128, 0, stream>>>(sc, pool, bp, r.cur)));
                k_end_tail<<<1, 1, 0, stream>>>(pool.ctrl, r.cur);
                launches++;
                tail_used = true;
                r.done = true;
                return MSK_OK;
            }
            if (r.bounce > 100000) return fail(MSK_ERR_CUDA, "path queue did not drain");
        }
        return MSK_OK;
    };

    // film accumulation of a finished batch, after the previous batch's (prev_lane < 0: the first batch of the job)
    auto finish = [&](int li, int prev_lane) -> int {
        Lane &ln = im.lanes[li];
        Pool &pool = ln.pool;
        cudaStream_t stream = ln.stream;
        Run &r = runs[li];
        const BatchParams &bp = r.bp;
        const uint32_t n = r.n;
        max_bounces = std::max(max_bounces, r.bounce);
        MSK_STAGE(ST_FILM, (k_film_records<<<(n + 255) / 256, 256, 0, stream>>>(sc, pool, bp, plan.rgba_channel)));
        if (prev_lane >= 0 && prev_lane != li) MSK_CUDA_CHECK(cudaStreamWaitEvent(stream, im.lanes[prev_lane].film_done, 0));
        dim3 fg((W + kFilmTileX - 1) / kFilmTileX, (H + kFilmTileY - 1) / kFilmTileY), fb(kFilmTileX, kFilmTileY);
        if ((int) std::ceil(sc.cam.filter_radius - 0.5f) <= kFilmMaxBorder)
            MSK_STAGE(ST_FILM, (k_film_gather_tiled<<<fg, fb, 0, stream>>>(sc, pool, bp, d_film, H, stride)));
        else
            MSK_STAGE(ST_FILM, (k_film_gather<<<fg, fb, 0, stream>>>(sc, pool, bp, d_film, H, stride)));
        if (plan.nch) MSK_STAGE(ST_FILM, (k_film_gather_aov<<<fg, fb, 0, stream>>>(sc, pool, bp, d_film, H, stride, plan.nch)));
        if (nlanes > 1) MSK_CUDA_CHECK(cudaEventRecord(ln.film_done, stream));
        batches++;
        r.active = false;
        return MSK_OK;
    };

    // every launch of the job (replayed from a CUDA graph when the sequence is fixed, see above)
    auto enqueue = [&]() -> int {
        if (rd.clear_film) MSK_CUDA_CHECK(cudaMemsetAsync(d_film, 0, (size_t) npix * stride * sizeof(float), stream0));
        if (nlanes > 1) { // the other lanes start after whatever precedes this job on the caller's stream
            MSK_CUDA_CHECK(cudaEventRecord(im.fork_ev, stream0));
            for (int l = 1; l < nlanes; ++l) MSK_CUDA_CHECK(cudaStreamWaitEvent(im.lanes[l].stream, im.fork_ev, 0));
        }
        for (int l = 0; l < nlanes; ++l)
            MSK_CUDA_CHECK(cudaMemsetAsync(&im.lanes[l].pool.ctrl->total_closest, 0, 10 * sizeof(unsigned long long), im.lanes[l].stream));
        uint32_t next_s0 = rd.sample_begin;
        uint64_t next_index = 0, next_finish = 0;
        int last_finished_lane = -1;
        for (;;) {
            bool any = false;
            for (int l = 0; l < nlanes; ++l) {
                Run &r = runs[l];
                if (!r.active && next_s0 < rd.sample_end) {
                    if ((rc = start(l, next_index++, next_s0))) return rc;
                    next_s0 += per_batch;
                }
                if (!r.active) continue;
                any = true;
                if (!r.done && (rc = step(l))) return rc;
                if (r.done && r.index == next_finish) { // the film accumulates its batches in order
                    if ((rc = finish(l, last_finished_lane))) return rc;
                    last_finished_lane = l;
                    next_finish++;
                }
            }
            if (!any) break;
        }
        for (int l = 1; l < nlanes; ++l) { // join: everything of the job precedes what follows on the caller's stream
            MSK_CUDA_CHECK(cudaEventRecord(im.lanes[l].film_done, im.lanes[l].stream));
            MSK_CUDA_CHECK(cudaStreamWaitEvent(stream0, im.lanes[l].film_done, 0));
        }
        return MSK_OK;
    };
#undef MSK_STAGE
    if (graphable) {
        std::vector<unsigned char> key(sizeof(DScene) + sizeof(MskRenderDesc) + sizeof(float *) + sizeof(AovPlan) + sizeof(int));
        unsigned char *k = key.data();
        memcpy(k, &sc, sizeof(DScene)); k += sizeof(DScene);
        memcpy(k, &rd, sizeof(MskRenderDesc)); k += sizeof(MskRenderDesc);
        memcpy(k, &d_film, sizeof(float *)); k += sizeof(float *);
        memcpy(k, &plan, sizeof(AovPlan)); k += sizeof(AovPlan);
        const int has_aov = aov != nullptr;
        memcpy(k, &has_aov, sizeof(int));
        if (!im.graph_exec || key != im.graph_key) {
            im.drop_graph();
            cudaGraph_t graph = nullptr;
            MSK_CUDA_CHECK(cudaStreamBeginCapture(stream0, cudaStreamCaptureModeThreadLocal));
            rc = enqueue();
            cudaError_t ce = cudaStreamEndCapture(stream0, &graph); // always end the capture, also after a failed enqueue
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture", __FILE__, __LINE__);
            ce = cudaGraphInstantiate(&im.graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { im.graph_exec = nullptr; return cuda_fail(ce, "cudaGraphInstantiate", __FILE__, __LINE__); }
            im.graph_key = key; im.graph_launches = launches; im.graph_bounces = max_bounces; im.graph_batches = batches; im.graph_extra_closest = extra_closest;
        }
        launches = im.graph_launches; max_bounces = im.graph_bounces; batches = im.graph_batches; extra_closest = im.graph_extra_closest;
        MSK_CUDA_CHECK(cudaEventRecord(im.ev[0], stream0));
        MSK_CUDA_CHECK(cudaGraphLaunch(im.graph_exec, stream0));
    } else {
        MSK_CUDA_CHECK(cudaEventRecord(im.ev[0], stream0));
        if ((rc = enqueue())) return rc;
    }
    MSK_CUDA_CHECK(cudaEventRecord(im.ev[1], stream0));
    MSK_CUDA_CHECK(cudaGetLastError());
    if (stats) {
        for (int l = 0; l < nlanes; ++l) // (after the join: every lane's counters are final in stream0's order)
            MSK_CUDA_CHECK(cudaMemcpyAsync(im.lanes[l].h_ctrl, im.lanes[l].pool.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream0));
        MSK_CUDA_CHECK(cudaStreamSynchronize(stream0));
        *stats = MskStats{};
        stats->paths = (uint64_t) npix * nsamples;
        stats->rays_closest = extra_closest;
        uint64_t tail_depth = 0;
        for (int l = 0; l < nlanes; ++l) {
            const Ctrl &hc = *im.lanes[l].h_ctrl;
            stats->rays_closest += hc.total_closest;
            stats->rays_shadow += hc.total_shadow;
            stats->shaded_vertices += hc.shaded;
            stats->nodes_closest += hc.nodes_closest; stats->tris_closest += hc.tris_closest;
            stats->nodes_shadow += hc.nodes_shadow; stats->tris_shadow += hc.tris_shadow;
            stats->tail_rays_closest += hc.tail_closest; stats->tail_rays_shadow += hc.tail_shadow;
            tail_depth = std::max<uint64_t>(tail_depth, hc.tail_depth);
        }
        stats->kernel_launches = launches;
        stats->bounces = tail_used ? std::max(max_bounces, (uint32_t) tail_depth) : max_bounces;
        stats->batches = batches;
        cudaEventElapsedTime(&stats->ms_render, im.ev[0], im.ev[1]);
        if (timers) {
            float acc[ST_COUNT] = {};
            uint32_t cnt[ST_COUNT] = {};
            for (size_t k = 0; k < tstage.size(); ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, im.timer_events[2 * k], im.timer_events[2 * k + 1]);
                if (im.debug_bounces) fprintf(stderr, "[msk] launch %zu stage %d: %.1f us\n", k, tstage[k], ms * 1e3f);
                acc[tstage[k]] += ms; cnt[tstage[k]]++;
            }
            stats->ms_raygen = acc[ST_RAYGEN]; stats->ms_intersect = acc[ST_INTERSECT]; stats->ms_shade = acc[ST_SHADE];
            stats->ms_shadow = acc[ST_SHADOW]; stats->ms_film = acc[ST_FILM]; stats->ms_sort = acc[ST_SORT];
            stats->n_intersect_launches = cnt[ST_INTERSECT]; stats->n_shade_launches = cnt[ST_SHADE];
            stats->n_shadow_launches = cnt[ST_SHADOW];
            stats->ms_tail = acc[ST_TAIL]; stats->n_tail_launches = cnt[ST_TAIL];
        }
    }
    return MSK_OK;
}

int Renderer::intersect(cudaStream_t stream, const DScene &sc, const MskRay *d_rays, MskHit *d_hits, size_t n) {
    if (n > 0xfffffff0ull) return fail(MSK_ERR_UNSUPPORTED, "too many rays in one call");
    if (!n) return MSK_OK;
    MSK_CUDA_CHECK(cudaMemsetAsync(impl_->query_cursor, 0, sizeof(uint32_t), stream));
    int blocks = (int) std::min<size_t>((size_t) impl_->persistent_blocks, (n + 127) / 128);
    k_query_closest<false><<<blocks, 128, 0, stream>>>(sc, d_rays, d_hits, (uint32_t) n, impl_->query_cursor, nullptr, nullptr);
    MSK_CUDA_CHECK(cudaGetLastError());
    return MSK_OK;
}

int Renderer::intersect_stats(cudaStream_t stream, const DScene &sc, const MskRay *d_rays, size_t n, uint32_t *d_nodes, uint32_t *d_tris) {
    if (n > 0xfffffff0ull) return fail(MSK_ERR_UNSUPPORTED, "too many rays in one call");
    if (!n) return MSK_OK;
    MSK_CUDA_CHECK(cudaMemsetAsync(impl_->query_cursor, 0, sizeof(uint32_t), stream));
    int blocks = (int) std::min<size_t>((size_t) impl_->persistent_blocks, (n + 127) / 128);
    k_query_closest<true><<<blocks, 128, 0, stream>>>(sc, d_rays, nullptr, (uint32_t) n, impl_->query_cursor, d_nodes, d_tris);
    MSK_CUDA_CHECK(cudaGetLastError());
    return MSK_OK;
}

int Renderer::occluded(cudaStream_t stream, const DScene &sc, const MskRay *d_rays, uint8_t *d_occ, size_t n) {
    if (n > 0xfffffff0ull) return fail(MSK_ERR_UNSUPPORTED, "too many rays in one call");
    if (!n) return MSK_OK;
    MSK_CUDA_CHECK(cudaMemsetAsync(impl_->query_cursor, 0, sizeof(uint32_t), stream));
    int blocks = (int) std::min<size_t>((size_t) impl_->persistent_blocks, (n + 127) / 128);
    k_query_any<<<blocks, 128, 0, stream>>>(sc, d_rays, d_occ, (uint32_t) n, impl_->query_cursor);
    MSK_CUDA_CHECK(cudaGetLastError());
    return MSK_OK;
}

} // namespace msk
