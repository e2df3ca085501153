// GPU BVH build for the closest-hit / any-hit kernels: replaces rtcCommitScene
// (reference src/librender/scene.cpp:201-212, Embree 3.12.2 SAH builder).
//
// Pipeline (all on the device, one stream):
//   1. k_tri_setup   gather each triangle's three positions (+prim, geom ids), reduce centroid bounds
//   2. k_morton      63-bit Morton code of the centroid
//   3. radix sort    (cub::DeviceRadixSort, build-time plumbing)
//   4. binary tree over the Morton-ordered triangles, one of
//        PLOC (MSK_BVH_BUILDER=ploc)  parallel locally-ordered clustering (Meister & Bittner 2018): every cluster finds the
//                        neighbour within +-kPlocRadius positions whose merged box has the smallest area, mutual
//                        pairs merge, the cluster list is compacted; repeat until one cluster is left.  The tree
//                        is SAH-driven (Embree's builder is an SAH builder too) instead of bit-prefix-driven.
//        LBVH (default)  k_karras (Karras 2012) + k_refit
//   6. k_collapse    level-synchronous top-down collapse of the binary tree into 8-wide nodes
//                    with 8-bit quantised child boxes (80 B per node, five 128-bit loads),
//                    inner children placed in octant-ordered slots (Ylitie, Karras, Laine 2017), leaves in the
//                    tightest run of free slots, their triangles at static slots 3j + k (msk_device.cuh)
#include "msk_device.cuh"
#include "msk_bvh.h"

#include <cub/cub.cuh>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace msk {

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n) { return (unsigned) ((n + kThreads - 1) / kThreads); }

// float <-> order-preserving uint for atomicMin/atomicMax
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

struct BuildState {
    uint32_t cmin[3], cmax[3];   // centroid bounds (ordered uints)
    uint32_t node_count;         // wide nodes allocated
    uint32_t tri_count;          // triangle SLOTS allocated (padded per-node blocks, msk_device.cuh)
    uint32_t tri_emitted;        // triangles written
    uint32_t queue_count;        // items produced for the next level
    uint32_t overflow;
};

__global__ void k_init_state(BuildState *st) {
    for (int a = 0; a < 3; ++a) { st->cmin[a] = 0xffffffffu; st->cmax[a] = 0u; }
    st->node_count = 1; st->tri_count = 0; st->tri_emitted = 0; st->queue_count = 0; st->overflow = 0;
}

// one launch per mesh: gathers vertices of its triangles into build order
__global__ void k_tri_setup(const float4 *__restrict__ verts, const uint32_t *__restrict__ indices, uint32_t vert_offset,
                            uint32_t tri_offset, uint32_t ntris, uint32_t geom, float4 *__restrict__ gathered,
                            BuildState *st) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cx = 0, cy = 0, cz = 0;
    bool valid = i < ntris;
    if (valid) {
        uint32_t g = tri_offset + i;
        uint32_t i0 = indices[3 * (size_t) g], i1 = indices[3 * (size_t) g + 1], i2 = indices[3 * (size_t) g + 2];
        float4 a = verts[2 * (size_t) (vert_offset + i0)], b = verts[2 * (size_t) (vert_offset + i1)],
               c = verts[2 * (size_t) (vert_offset + i2)];
        gathered[3 * (size_t) g + 0] = make_float4(a.x, a.y, a.z, __uint_as_float(i));
        gathered[3 * (size_t) g + 1] = make_float4(b.x, b.y, b.z, __uint_as_float(geom));
        gathered[3 * (size_t) g + 2] = make_float4(c.x, c.y, c.z, 0.f);
        cx = 0.5f * (fminf(a.x, fminf(b.x, c.x)) + fmaxf(a.x, fmaxf(b.x, c.x)));
        cy = 0.5f * (fminf(a.y, fminf(b.y, c.y)) + fmaxf(a.y, fmaxf(b.y, c.y)));
        cz = 0.5f * (fminf(a.z, fminf(b.z, c.z)) + fmaxf(a.z, fmaxf(b.z, c.z)));
    }
    // warp reduce, then one atomic per warp and axis
    float mn[3] = { valid ? cx : FLT_MAX, valid ? cy : FLT_MAX, valid ? cz : FLT_MAX };
    float mx[3] = { valid ? cx : -FLT_MAX, valid ? cy : -FLT_MAX, valid ? cz : -FLT_MAX };
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if ((threadIdx.x & 31) == 0 && mn[0] != FLT_MAX)
        for (int a = 0; a < 3; ++a) { atomicMin(&st->cmin[a], f2ord(mn[a])); atomicMax(&st->cmax[a], f2ord(mx[a])); }
}

__device__ __forceinline__ uint64_t expand21(uint32_t v) { // spread 21 bits to every third bit
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4 *__restrict__ gathered, uint32_t n, const BuildState *st, uint64_t *__restrict__ keys,
                         uint32_t *__restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = gathered[3 * (size_t) i], b = gathered[3 * (size_t) i + 1], c = gathered[3 * (size_t) i + 2];
    float ctr[3] = { 0.5f * (fminf(a.x, fminf(b.x, c.x)) + fmaxf(a.x, fmaxf(b.x, c.x))),
                     0.5f * (fminf(a.y, fminf(b.y, c.y)) + fmaxf(a.y, fmaxf(b.y, c.y))),
                     0.5f * (fminf(a.z, fminf(b.z, c.z)) + fmaxf(a.z, fmaxf(b.z, c.z))) };
    uint32_t q[3];
    for (int k = 0; k < 3; ++k) {
        float lo = ord2f(st->cmin[k]), hi = ord2f(st->cmax[k]);
        float ext = hi - lo;
        float t = ext > 0.f ? (ctr[k] - lo) / ext : 0.f;
        q[k] = (uint32_t) fminf(fmaxf(t * 2097152.f, 0.f), 2097151.f);
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[i] = i;
}

// ---- binary radix tree (Karras 2012).  Internal nodes 0..n-2, leaves n-1..2n-2. ----
struct Bvh2 {
    uint32_t *left, *right, *parent; // per internal node (parent: per node, 2n-1; LBVH only)
    uint32_t *count;                 // per internal node: triangles below it
    float4 *lo, *hi;                 // per node (2n-1)
    uint32_t *flags;                 // per internal node, refit arrival counter
};

__device__ __forceinline__ int delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t) i ^ (uint32_t) j);
    return __clzll((long long) (a ^ b));
}

__global__ void k_karras(const uint64_t *__restrict__ keys, int n, Bvh2 t) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d     = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin  = delta(keys, n, i, i - d);
    int lmax  = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s >= 1; s >>= 1)
        if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    int j     = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0, tt = l;
    do {
        tt = (tt + 1) >> 1;
        if (delta(keys, n, i, i + (s + tt) * d) > dnode) s += tt;
    } while (tt > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    uint32_t lc = (lo == gamma) ? (uint32_t) (n - 1 + gamma) : (uint32_t) gamma;
    uint32_t rc = (hi == gamma + 1) ? (uint32_t) (n - 1 + gamma + 1) : (uint32_t) (gamma + 1);
    t.left[i] = lc; t.right[i] = rc;
    t.parent[lc] = i; t.parent[rc] = i;
    t.count[i] = (uint32_t) (hi - lo + 1);
    if (i == 0) t.parent[0] = 0xffffffffu;
}

__global__ void k_refit(const float4 *__restrict__ gathered, const uint32_t *__restrict__ sorted, int n, Bvh2 t) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t tri = sorted[j];
    float4 a = gathered[3 * (size_t) tri], b = gathered[3 * (size_t) tri + 1], c = gathered[3 * (size_t) tri + 2];
    float4 lo = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.f);
    float4 hi = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.f);
    uint32_t node = n - 1 + j;
    t.lo[node] = lo; t.hi[node] = hi;
    if (n == 1) return;
    uint32_t p = t.parent[node];
    while (p != 0xffffffffu) {
        __threadfence();
        if (atomicAdd(&t.flags[p], 1u) == 0) return; // first arrival: the sibling will finish this node
        uint32_t l = t.left[p], r = t.right[p];
        // .cg loads: the sibling's stores were made visible by its fence + the atomic
        float4 llo = __ldcg(&t.lo[l]), lhi = __ldcg(&t.hi[l]), rlo = __ldcg(&t.lo[r]), rhi = __ldcg(&t.hi[r]);
        lo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
        hi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
        t.lo[p] = lo; t.hi[p] = hi;
        p = t.parent[p];
    }
}

__device__ __forceinline__ float half_area(float4 lo, float4 hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ uint32_t node_count(const Bvh2 &t, int n, uint32_t node) {
    return node >= (uint32_t) (n - 1) ? 1u : t.count[node];
}

// ---- PLOC (Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction",
// 2018).  Clusters live in Morton order in compact arrays (node id + box, ping-ponged by the compaction). ----
constexpr int kPlocRadius = 16;

__global__ void k_ploc_init(const float4 *__restrict__ gathered, const uint32_t *__restrict__ sorted, int n, Bvh2 t,
                            uint32_t *__restrict__ cid, float4 *__restrict__ clo, float4 *__restrict__ chi) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t tri = sorted[j];
    float4 a = gathered[3 * (size_t) tri], b = gathered[3 * (size_t) tri + 1], c = gathered[3 * (size_t) tri + 2];
    float4 lo = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.f);
    float4 hi = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.f);
    uint32_t node = (uint32_t) (n - 1 + j);
    t.lo[node] = lo; t.hi[node] = hi;
    cid[j] = node; clo[j] = lo; chi[j] = hi;
}

// nearest neighbour of every cluster within +-kPlocRadius positions: smallest merged half-area, ties to the lower
// position (so that the globally best pair is always mutual and every round merges at least one pair)
__global__ void __launch_bounds__(kThreads) k_ploc_nn(uint32_t c, const float4 *__restrict__ clo, const float4 *__restrict__ chi,
                                                       uint32_t *__restrict__ nn) {
    __shared__ float s_lo[3][kThreads + 2 * kPlocRadius], s_hi[3][kThreads + 2 * kPlocRadius];
    const long long base = (long long) blockIdx.x * kThreads - kPlocRadius;
    for (int e = threadIdx.x; e < kThreads + 2 * kPlocRadius; e += kThreads) {
        long long g = base + e;
        if (g >= 0 && g < (long long) c) {
            float4 l = clo[g], h = chi[g];
            s_lo[0][e] = l.x; s_lo[1][e] = l.y; s_lo[2][e] = l.z; s_hi[0][e] = h.x; s_hi[1][e] = h.y; s_hi[2][e] = h.z;
        }
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= c) return;
    const int me = threadIdx.x + kPlocRadius;
    const float lx = s_lo[0][me], ly = s_lo[1][me], lz = s_lo[2][me], hx = s_hi[0][me], hy = s_hi[1][me], hz = s_hi[2][me];
    float best = FLT_MAX;
    uint32_t bj = 0xffffffffu;
    for (int d = -kPlocRadius; d <= kPlocRadius; ++d) {
        if (d == 0) continue;
        const long long g = (long long) i + d;
        if (g < 0 || g >= (long long) c) continue;
        const int e = me + d;
        const float dx = fmaxf(hx, s_hi[0][e]) - fminf(lx, s_lo[0][e]), dy = fmaxf(hy, s_hi[1][e]) - fminf(ly, s_lo[1][e]),
                    dz = fmaxf(hz, s_hi[2][e]) - fminf(lz, s_lo[2][e]);
        const float a = dx * dy + dy * dz + dz * dx;
        if (a < best) { best = a; bj = (uint32_t) g; } // ascending g: the first minimum is the lowest position
    }
    nn[i] = bj;
}

// mutual nearest neighbours merge into a new binary node at the lower position; ids descend from n-2 so that the
// last merge (the root) is node 0
__global__ void k_ploc_merge(uint32_t c, const uint32_t *__restrict__ nn, uint32_t *__restrict__ cid, float4 *__restrict__ clo,
                             float4 *__restrict__ chi, uint32_t *__restrict__ flags, Bvh2 t, int n, uint32_t *counter) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const uint32_t j = nn[i];
    uint32_t keep = 1u;
    if (j != 0xffffffffu && nn[j] == i) {
        if (i < j) {
            const uint32_t node = (uint32_t) (n - 2) - atomicAdd(counter, 1u);
            const uint32_t l = cid[i], r = cid[j];
            const float4 llo = clo[i], lhi = chi[i], rlo = clo[j], rhi = chi[j];
            const float4 lo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
            const float4 hi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
            t.left[node] = l; t.right[node] = r; t.lo[node] = lo; t.hi[node] = hi;
            t.count[node] = node_count(t, n, l) + node_count(t, n, r);
            cid[i] = node; clo[i] = lo; chi[i] = hi;
        } else keep = 0u;
    }
    flags[i] = keep;
}

__global__ void k_ploc_compact(uint32_t c, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos,
                               const uint32_t *__restrict__ cid, const float4 *__restrict__ clo, const float4 *__restrict__ chi,
                               uint32_t *__restrict__ cid2, float4 *__restrict__ clo2, float4 *__restrict__ chi2,
                               uint32_t *__restrict__ count_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const uint32_t f = flags[i], p = pos[i];
    if (f) { cid2[p] = cid[i]; clo2[p] = clo[i]; chi2[p] = chi[i]; }
    if (i == c - 1) *count_out = p + f;
}

// exponent byte e such that 2^(e-127) >= extent / 255
__device__ __forceinline__ uint32_t quant_exp(float extent) {
    float s = extent / 255.f;
    uint32_t u = __float_as_uint(s);
    uint32_t e = (u >> 23) & 0xffu;
    if (u & 0x7fffffu) e += 1;
    return min(max(e, 1u), 254u);
}

// ---- SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 4.1) --------------------------------------------
// Which binary nodes become the (at most 8) children of a wide node is decided by dynamic programming over the binary
// tree instead of greedily opening the largest child.  C(n, i) = the cheapest way to represent the subtree of binary
// node n by at most i roots in its parent's child list, a root being either a LEAF SLOT (<= kMaxLeafTris triangles
// behind one child box: cost area(n) * triangles * c_prim -- the triangles are tested whenever the box is entered) or a
// wide INNER node (cost area(n) * c_node + the cheapest distribution of its 8 slots over n's two children):
//     D(n, j)  = min over 0 < k < j of C(left, k) + C(right, j - k)                      j = 2..8
//     C(n, 1)  = min(C_leaf(n), area(n) * c_node + D(n, 8))
//     C(n, i)  = min(D(n, i), C(n, i - 1))                                               i = 2..7
// (areas are half areas; only ratios matter).  Bottom-up pass k_plan stores C(n, 1..7) and the arg-mins, the top-down
// collapse reads them back (expand below).  The greedy rule -- open the largest-area candidate while slots remain -- on
// a uniformly tessellated mesh hands every slot a subtree of equal size, powers of two: 29 % of C2's wide nodes had
// exactly 4 children (46 % under PLOC, whose 3x lower binary-tree SAH the collapse then wasted;
// gpurun_out/r02p_bvh_shape.txt).
struct Plan {
    float   *cost;  // 7 per internal binary node: C(n, 1..7)
    uint8_t *split; // 8 per internal binary node: [0] = 1 when C(n, 1) is a leaf slot; [j - 1], j = 2..7 = roots handed to the
                    // left child in D(n, j), or 0 when j - 1 roots are as cheap; [7] = the left share of D(n, 8)
};

__global__ void k_set_parents(Bvh2 t, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    t.parent[t.left[i]] = (uint32_t) i; t.parent[t.right[i]] = (uint32_t) i;
    if (i == 0) t.parent[0] = 0xffffffffu;
}

__device__ __forceinline__ void plan_costs(const Bvh2 &t, int n, const Plan &pl, uint32_t node, float c_prim, float (&c)[7]) {
    if (node >= (uint32_t) (n - 1)) { // a single triangle: one leaf slot, however many roots are allowed
        const float a = half_area(__ldcg(&t.lo[node]), __ldcg(&t.hi[node])) * c_prim;
#pragma unroll
        for (int i = 0; i < 7; ++i) c[i] = a;
    } else {
#pragma unroll
        for (int i = 0; i < 7; ++i) c[i] = __ldcg(pl.cost + 7 * (size_t) node + i);
    }
}

__global__ void k_plan(Bvh2 t, int n, Plan pl, float c_node, float c_prim) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || n == 1) return;
    uint32_t p = t.parent[n - 1 + j];
    while (p != 0xffffffffu) {
        __threadfence();
        if (atomicAdd(&t.flags[p], 1u) == 0) return; // first arrival: the sibling will finish this node
        float cl[7], cr[7];
        plan_costs(t, n, pl, t.left[p], c_prim, cl);
        plan_costs(t, n, pl, t.right[p], c_prim, cr);
        const float area = half_area(t.lo[p], t.hi[p]);
        float d[9];
        uint8_t kk[9];
#pragma unroll
        for (int jj = 2; jj <= 8; ++jj) {
            float best = FLT_MAX;
            uint8_t bk = 1;
#pragma unroll
            for (int k = 1; k < jj; ++k) {
                const float v = cl[k - 1] + cr[jj - k - 1]; // (k and jj - k are both <= 7)
                if (v < best) { best = v; bk = (uint8_t) k; }
            }
            d[jj] = best; kk[jj] = bk;
        }
        const uint32_t cnt = t.count[p];
        const float c_leaf = cnt <= (uint32_t) kMaxLeafTris ? area * (float) cnt * c_prim : FLT_MAX;
        const float c_inner = area * c_node + d[8];
        float c[8];
        uint8_t sp[8];
        c[1] = fminf(c_leaf, c_inner);
        sp[0] = c_leaf <= c_inner ? 1 : 0;
#pragma unroll
        for (int i = 2; i <= 7; ++i) {
            if (d[i] < c[i - 1]) { c[i] = d[i]; sp[i - 1] = kk[i]; }
            else { c[i] = c[i - 1]; sp[i - 1] = 0; }
        }
        sp[7] = kk[8];
#pragma unroll
        for (int i = 0; i < 7; ++i) pl.cost[7 * (size_t) p + i] = c[i + 1];
#pragma unroll
        for (int i = 0; i < 8; ++i) pl.split[8 * (size_t) p + i] = sp[i];
        p = t.parent[p];
    }
}

struct WorkItem { uint32_t bvh2, wide; };

// One thread builds one wide node from the binary subtree rooted at item.bvh2.
__global__ void k_collapse(const WorkItem *__restrict__ in, uint32_t nin, WorkItem *__restrict__ out, uint32_t out_cap,
                           Bvh2 t, int n, const uint32_t *__restrict__ sorted, const float4 *__restrict__ gathered,
                           float4 *__restrict__ nodes, uint32_t node_cap, float4 *__restrict__ tris, uint32_t tri_cap,
                           BuildState *st, int root_is_leaf, Plan pl) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nin) return;
    WorkItem item = in[w];
    uint32_t cand[8];
    int nc = 0;
    if (root_is_leaf) {
        cand[nc++] = item.bvh2; // n == 1: the single triangle
    } else if (pl.split) {
        // the 8 slots of this node distributed as k_plan found cheapest: (node, roots) pairs, left before right
        uint32_t snode[8];
        uint8_t sroots[8];
        int sp = 0;
        const uint32_t r = item.bvh2;
        const uint8_t k8 = pl.split[8 * (size_t) r + 7];
        snode[sp] = t.right[r]; sroots[sp++] = (uint8_t) (8 - k8);
        snode[sp] = t.left[r];  sroots[sp++] = k8;
        while (sp) {
            const uint32_t u = snode[--sp];
            const uint8_t j = sroots[sp];
            if (u >= (uint32_t) (n - 1) || j == 1) { cand[nc++] = u; continue; }
            const uint8_t k = pl.split[8 * (size_t) u + (j - 1)];
            if (k == 0) { snode[sp] = u; sroots[sp++] = (uint8_t) (j - 1); } // one root fewer is as cheap
            else {
                snode[sp] = t.right[u]; sroots[sp++] = (uint8_t) (j - k);
                snode[sp] = t.left[u];  sroots[sp++] = k;
            }
        }
    } else {
        cand[nc++] = t.left[item.bvh2];
        cand[nc++] = t.right[item.bvh2];
        // greedily open the largest-area internal candidate while slots remain
        while (nc < 8) {
            int best = -1;
            float best_area = -1.f;
            for (int i = 0; i < nc; ++i) {
                if (cand[i] >= (uint32_t) (n - 1)) continue; // single triangle
                float a = half_area(t.lo[cand[i]], t.hi[cand[i]]);
                if (a > best_area) { best_area = a; best = i; }
            }
            if (best < 0) break;
            uint32_t o = cand[best];
            cand[best] = t.left[o];
            cand[nc++] = t.right[o];
        }
    }
    float4 plo = t.lo[item.bvh2], phi = t.hi[item.bvh2];
    float pc[3] = { 0.5f * (plo.x + phi.x), 0.5f * (plo.y + phi.y), 0.5f * (plo.z + phi.z) };
    // Inner children: octant-ordered slot assignment -- slot s prefers the child furthest along
    // d_s = (s&4 ? + : -, s&2 ? + : -, s&1 ? + : -), so that visiting slots in the order s ^ octinv is front to back.
    // Leaves are all tested before the ray descends, so their slots carry no order: they take the tightest run of
    // the slots the inner children left free (their triangle block is padded over that run).
    float cost[8][8];
    float4 clo[8], chi[8];
    uint32_t cnts[8];
    uint32_t inner_cand = 0, nleaf = 0;
    for (int c = 0; c < nc; ++c) {
        clo[c] = t.lo[cand[c]]; chi[c] = t.hi[cand[c]];
        cnts[c] = node_count(t, n, cand[c]);
        const bool inner = pl.split ? (cand[c] < (uint32_t) (n - 1) && pl.split[8 * (size_t) cand[c]] == 0) : cnts[c] > (uint32_t) kMaxLeafTris;
        if (inner) inner_cand |= 1u << c; else nleaf++;
        float dx = 0.5f * (clo[c].x + chi[c].x) - pc[0], dy = 0.5f * (clo[c].y + chi[c].y) - pc[1],
              dz = 0.5f * (clo[c].z + chi[c].z) - pc[2];
        for (int s = 0; s < 8; ++s)
            cost[c][s] = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
    }
    int slot_child[8];
    for (int s = 0; s < 8; ++s) slot_child[s] = -1;
    uint32_t child_done = 0, slot_done = 0, imask = 0, ninner = 0;
    for (int k = 0; k < nc; ++k) {
        float best = -FLT_MAX; int bc = -1, bs = -1;
        for (int c = 0; c < nc; ++c) {
            if ((child_done & (1u << c)) || !(inner_cand & (1u << c))) continue;
            for (int s = 0; s < 8; ++s) {
                if (slot_done & (1u << s)) continue;
                if (cost[c][s] > best) { best = cost[c][s]; bc = c; bs = s; }
            }
        }
        if (bc < 0) break;
        slot_child[bs] = bc; child_done |= 1u << bc; slot_done |= 1u << bs;
        imask |= 1u << bs; ninner++;
    }
    uint32_t span = 0, smin = 0;
    if (nleaf) {
        int freeslots[8], nf = 0;
        for (int s = 0; s < 8; ++s) if (!(slot_done & (1u << s))) freeslots[nf++] = s;
        int bi = 0, bspan = 99;
        for (int i = 0; i + (int) nleaf <= nf; ++i) {
            const int sp_ = freeslots[i + nleaf - 1] - freeslots[i] + 1;
            if (sp_ < bspan) { bspan = sp_; bi = i; }
        }
        int k = bi;
        for (int c = 0; c < nc; ++c)
            if (!(inner_cand & (1u << c))) slot_child[freeslots[k++]] = c;
        smin = (uint32_t) freeslots[bi];
        span = 3u * (uint32_t) bspan;
    }
    uint32_t child_base = ninner ? atomicAdd(&st->node_count, ninner) : 0u;
    uint32_t tri_alloc  = span ? atomicAdd(&st->tri_count, span) : 0u;
    uint32_t qbase      = ninner ? atomicAdd(&st->queue_count, ninner) : 0u;
    if (child_base + ninner > node_cap || qbase + ninner > out_cap || (uint64_t) tri_alloc + span > (uint64_t) tri_cap) { st->overflow = 1; return; }
    const uint32_t tri_base = tri_alloc - 3u * smin; // slot 3j + k of this node lives at tri_base + 3j + k (mod 2^32)

    uint32_t ex = quant_exp(phi.x - plo.x), ey = quant_exp(phi.y - plo.y), ez = quant_exp(phi.z - plo.z);
    float sx = __uint_as_float(ex << 23), sy = __uint_as_float(ey << 23), sz = __uint_as_float(ez << 23);
    uint32_t q[6][2] = {};
    uint32_t inner_rank = 0, trimask = 0, emitted = 0;
    for (int s = 0; s < 8; ++s) {
        int c = slot_child[s];
        if (c < 0) continue;
        uint32_t node = cand[c];
        uint32_t cnt  = cnts[c];
        if (imask & (1u << s)) {
            out[qbase + inner_rank] = WorkItem{ node, child_base + inner_rank };
            inner_rank++;
        } else {
            trimask |= ((1u << cnt) - 1u) << (3 * s);
            // the <= kMaxLeafTris triangles of the subtree, left to right
            uint32_t todo[kMaxLeafTris + 1];
            int sp = 0;
            uint32_t k = 0;
            todo[sp++] = node;
            while (sp) {
                const uint32_t u = todo[--sp];
                if (u >= (uint32_t) (n - 1)) {
                    const uint32_t tri = sorted[u - (uint32_t) (n - 1)];
                    const size_t dst = 3 * (size_t) (uint32_t) (tri_base + 3u * (uint32_t) s + k);
                    tris[dst + 0] = gathered[3 * (size_t) tri + 0];
                    tris[dst + 1] = gathered[3 * (size_t) tri + 1];
                    tris[dst + 2] = gathered[3 * (size_t) tri + 2];
                    ++k;
                } else { todo[sp++] = t.right[u]; todo[sp++] = t.left[u]; }
            }
            emitted += cnt;
        }
        // conservative 8-bit quantisation against the grid origin plo with cell size 2^e
        float lo3[3] = { clo[c].x, clo[c].y, clo[c].z }, hi3[3] = { chi[c].x, chi[c].y, chi[c].z };
        float p3[3] = { plo.x, plo.y, plo.z }, s3[3] = { sx, sy, sz };
        for (int a = 0; a < 3; ++a) {
            int ql = (int) floorf((lo3[a] - p3[a]) / s3[a]);
            int qh = (int) ceilf((hi3[a] - p3[a]) / s3[a]);
            ql = min(max(ql, 0), 255); qh = min(max(qh, 0), 255);
            while (ql > 0 && p3[a] + (float) ql * s3[a] > lo3[a]) --ql;
            while (qh < 255 && p3[a] + (float) qh * s3[a] < hi3[a]) ++qh;
            q[a][s >> 2] |= (uint32_t) ql << (8 * (s & 3));
            q[3 + a][s >> 2] |= (uint32_t) qh << (8 * (s & 3));
        }
    }
    if (emitted) atomicAdd(&st->tri_emitted, emitted);
    float4 *dst = nodes + (size_t) item.wide * kNodeFloat4s;
    dst[0] = make_float4(plo.x, plo.y, plo.z, __uint_as_float(ex | (ey << 8) | (ez << 16) | (imask << 24)));
    dst[1] = make_float4(__uint_as_float(child_base), __uint_as_float(tri_base), __uint_as_float((imask << 24) | trimask), 0.f);
    dst[2] = make_float4(__uint_as_float(q[0][0]), __uint_as_float(q[0][1]), __uint_as_float(q[1][0]), __uint_as_float(q[1][1]));
    dst[3] = make_float4(__uint_as_float(q[2][0]), __uint_as_float(q[2][1]), __uint_as_float(q[3][0]), __uint_as_float(q[3][1]));
    dst[4] = make_float4(__uint_as_float(q[4][0]), __uint_as_float(q[4][1]), __uint_as_float(q[5][0]), __uint_as_float(q[5][1]));
}

__global__ void k_empty_root(float4 *nodes) { // scene without triangles: every ray misses
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        nodes[0] = make_float4(0.f, 0.f, 0.f, __uint_as_float(1u | (1u << 8) | (1u << 16)));
        for (int i = 1; i < kNodeFloat4s; ++i) nodes[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// sum of (area(child)/area(root)) over wide nodes and leaves: an SAH-style quality figure
__global__ void k_sah(const Bvh2 t, int n, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (i < n - 1) v = half_area(t.lo[i], t.hi[i]);
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(out, v);
}

// MSK_DEBUG_SETUP: shape of the wide tree -- histogram of inner / leaf children per node and of triangles per leaf
// out[0..8] nodes with k inner children | out[9..17] nodes with k leaf children | out[18..26] nodes with k children in total |
// out[27..30] leaves with 0..3 triangles
__global__ void k_shape(const float4 *__restrict__ nodes, uint32_t nnodes, unsigned long long *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    const uint32_t valid = __float_as_uint(nodes[(size_t) i * kNodeFloat4s + 1].z);
    const uint32_t ninner = __popc(valid >> 24);
    uint32_t nleaf = 0;
    for (int j = 0; j < 8; ++j) {
        const uint32_t t = __popc((valid >> (3 * j)) & 7u);
        if (t) { nleaf++; atomicAdd(out + 27 + t, 1ull); }
    }
    atomicAdd(out + ninner, 1ull); atomicAdd(out + 9 + nleaf, 1ull); atomicAdd(out + 18 + ninner + nleaf, 1ull);
}

template <typename T> cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **) p, std::max<size_t>(n, 1) * sizeof(T)); }

} // namespace

void bvh_print_shape(cudaStream_t stream, const BvhResult &r) {
    unsigned long long *d = nullptr, h[31] = {};
    if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return;
    cudaMemsetAsync(d, 0, sizeof(h), stream);
    k_shape<<<blocks_for(r.nnodes), kThreads, 0, stream>>>(r.nodes, (uint32_t) r.nnodes, d);
    cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    cudaFree(d);
    fprintf(stderr, "[msk] wide BVH: %llu nodes, depth %u, sah %.2f | nodes by inner children 0..8:", (unsigned long long) r.nnodes, r.depth, r.sah_cost);
    for (int k = 0; k <= 8; ++k) fprintf(stderr, " %llu", h[k]);
    fprintf(stderr, " | by leaf children 0..8:");
    for (int k = 0; k <= 8; ++k) fprintf(stderr, " %llu", h[9 + k]);
    fprintf(stderr, " | by children 0..8:");
    double tot = 0, cnt = 0;
    for (int k = 0; k <= 8; ++k) { fprintf(stderr, " %llu", h[18 + k]); tot += (double) k * h[18 + k]; cnt += (double) h[18 + k]; }
    fprintf(stderr, " (mean %.2f) | leaves with 1..3 triangles: %llu %llu %llu\n", cnt ? tot / cnt : 0.0, h[28], h[29], h[30]);
}

static int bvh_build_impl(cudaStream_t stream, const float4 *d_verts, const uint32_t *d_indices, const std::vector<DMeshInfo> &meshes,
                          BvhResult *out, int builder, size_t tri_cap_override) {
    *out = BvhResult{};
    out->builder = builder;
    size_t n = 0;
    for (auto &m : meshes) n += m.ntris;
    if (n > 0x7ffffff0ull) return fail(MSK_ERR_UNSUPPORTED, "too many triangles (%zu)", n);
    cudaEvent_t e0, e1;
    MSK_CUDA_CHECK(cudaEventCreate(&e0));
    MSK_CUDA_CHECK(cudaEventCreate(&e1));
    MSK_CUDA_CHECK(cudaEventRecord(e0, stream));

    size_t node_cap = n ? (size_t) (0.66 * (double) n) + 16 : 1;
    // Triangle slots are padded per node (msk_device.cuh): at most 24 per node with leaves; measured ~2.5-3 per
    // triangle on the BASELINE meshes.  Build into a generous block, shrink to fit afterwards; the rare overflow
    // (reported by the collapse) is retried with the worst-case bound.
    size_t tri_cap = std::min<size_t>(6 * n + 1024, 24 * node_cap);
    if (tri_cap_override) tri_cap = tri_cap_override;
    if (tri_cap > 0xfffffff0ull) return fail(MSK_ERR_UNSUPPORTED, "too many triangles (%zu) for the padded leaf layout", n);
    MSK_CUDA_CHECK(dalloc(&out->nodes, node_cap * kNodeFloat4s));
    MSK_CUDA_CHECK(dalloc(&out->tris, tri_cap * kTriFloat4s));
    if (n == 0) {
        k_empty_root<<<1, 32, 0, stream>>>(out->nodes);
        out->nnodes = 1; out->ntris = 0; out->tri_slots = 0; out->depth = 1;
        MSK_CUDA_CHECK(cudaEventRecord(e1, stream));
        MSK_CUDA_CHECK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&out->ms_build, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return MSK_OK;
    }

    float4 *gathered = nullptr;
    uint64_t *keys = nullptr, *keys_sorted = nullptr;
    uint32_t *vals = nullptr, *sorted = nullptr;
    BuildState *st = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    Bvh2 t{};
    WorkItem *qa = nullptr, *qb = nullptr;
    float *sah = nullptr;
    Plan pl{ nullptr, nullptr };
    uint32_t *p_cid[2] = { nullptr, nullptr }, *p_nn = nullptr, *p_flags = nullptr, *p_pos = nullptr, *p_counters = nullptr;
    float4 *p_lo[2] = { nullptr, nullptr }, *p_hi[2] = { nullptr, nullptr };
    void *p_tmp = nullptr;
    int rc = MSK_OK;
    auto cleanup = [&]() {
        cudaFree(gathered); cudaFree(keys); cudaFree(keys_sorted); cudaFree(vals); cudaFree(sorted); cudaFree(st);
        cudaFree(cub_tmp); cudaFree(t.left); cudaFree(t.right); cudaFree(t.parent); cudaFree(t.count);
        cudaFree(t.lo); cudaFree(t.hi); cudaFree(t.flags); cudaFree(qa); cudaFree(qb); cudaFree(sah);
        cudaFree(pl.cost); cudaFree(pl.split);
        for (int k = 0; k < 2; ++k) { cudaFree(p_cid[k]); cudaFree(p_lo[k]); cudaFree(p_hi[k]); }
        cudaFree(p_nn); cudaFree(p_flags); cudaFree(p_pos); cudaFree(p_tmp); cudaFree(p_counters);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    };
#define BVH_CHECK(expr)                                                          \
    do {                                                                         \
        cudaError_t err__ = (expr);                                              \
        if (err__ != cudaSuccess) { rc = cuda_fail(err__, #expr, __FILE__, __LINE__); cleanup(); \
            cudaFree(out->nodes); cudaFree(out->tris); *out = BvhResult{}; return rc; } \
    } while (0)

    BVH_CHECK(dalloc(&gathered, n * 3));
    BVH_CHECK(dalloc(&keys, n)); BVH_CHECK(dalloc(&keys_sorted, n));
    BVH_CHECK(dalloc(&vals, n)); BVH_CHECK(dalloc(&sorted, n));
    BVH_CHECK(dalloc(&st, 1));
    BVH_CHECK(dalloc(&sah, 1));
    BVH_CHECK(cudaMemsetAsync(sah, 0, sizeof(float), stream));
    k_init_state<<<1, 1, 0, stream>>>(st);
    for (uint32_t g = 0; g < meshes.size(); ++g) {
        const DMeshInfo &m = meshes[g];
        if (!m.ntris) continue;
        k_tri_setup<<<blocks_for(m.ntris), kThreads, 0, stream>>>(d_verts, d_indices, m.vert_offset, m.tri_offset, m.ntris, g,
                                                                  gathered, st);
    }
    k_morton<<<blocks_for(n), kThreads, 0, stream>>>(gathered, (uint32_t) n, st, keys, vals);
    BVH_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys, keys_sorted, vals, sorted, (int) n, 0, 63, stream));
    BVH_CHECK(cudaMalloc(&cub_tmp, std::max<size_t>(cub_bytes, 16)));
    BVH_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys, keys_sorted, vals, sorted, (int) n, 0, 63, stream));

    size_t nint = n > 1 ? n - 1 : 1;
    BVH_CHECK(dalloc(&t.left, nint)); BVH_CHECK(dalloc(&t.right, nint));
    BVH_CHECK(dalloc(&t.count, nint));
    BVH_CHECK(dalloc(&t.lo, 2 * n)); BVH_CHECK(dalloc(&t.hi, 2 * n));
    if (builder == MSK_BVH_PLOC) {
        for (int k = 0; k < 2; ++k) { BVH_CHECK(dalloc(&p_cid[k], n)); BVH_CHECK(dalloc(&p_lo[k], n)); BVH_CHECK(dalloc(&p_hi[k], n)); }
        BVH_CHECK(dalloc(&p_nn, n)); BVH_CHECK(dalloc(&p_flags, n)); BVH_CHECK(dalloc(&p_pos, n));
        BVH_CHECK(dalloc(&p_counters, 2));
        BVH_CHECK(cudaMemsetAsync(p_counters, 0, 2 * sizeof(uint32_t), stream));
        size_t scan_bytes = 0;
        BVH_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, p_flags, p_pos, (int) n, stream));
        BVH_CHECK(cudaMalloc(&p_tmp, std::max<size_t>(scan_bytes, 16)));
        k_ploc_init<<<blocks_for(n), kThreads, 0, stream>>>(gathered, sorted, (int) n, t, p_cid[0], p_lo[0], p_hi[0]);
        uint32_t c = (uint32_t) n, rounds = 0;
        int cur = 0;
        while (c > 1) {
            k_ploc_nn<<<blocks_for(c), kThreads, 0, stream>>>(c, p_lo[cur], p_hi[cur], p_nn);
            k_ploc_merge<<<blocks_for(c), kThreads, 0, stream>>>(c, p_nn, p_cid[cur], p_lo[cur], p_hi[cur], p_flags, t, (int) n,
                                                                 p_counters);
            BVH_CHECK(cub::DeviceScan::ExclusiveSum(p_tmp, scan_bytes, p_flags, p_pos, (int) c, stream));
            k_ploc_compact<<<blocks_for(c), kThreads, 0, stream>>>(c, p_flags, p_pos, p_cid[cur], p_lo[cur], p_hi[cur], p_cid[cur ^ 1],
                                                                   p_lo[cur ^ 1], p_hi[cur ^ 1], p_counters + 1);
            uint32_t c2 = 0;
            BVH_CHECK(cudaMemcpyAsync(&c2, p_counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            BVH_CHECK(cudaStreamSynchronize(stream));
            if (c2 == 0 || c2 >= c) { cleanup(); cudaFree(out->nodes); cudaFree(out->tris); *out = BvhResult{};
                return fail(MSK_ERR_CUDA, "PLOC made no progress (%u -> %u clusters)", c, c2); }
            c = c2; cur ^= 1; rounds++;
        }
        out->build_rounds = rounds;
        BVH_CHECK(dalloc(&t.flags, nint));
        BVH_CHECK(dalloc(&t.parent, 2 * n));
        if (n > 1) k_set_parents<<<blocks_for(n - 1), kThreads, 0, stream>>>(t, (int) n);
    } else {
        BVH_CHECK(dalloc(&t.flags, nint));
        BVH_CHECK(dalloc(&t.parent, 2 * n));
        BVH_CHECK(cudaMemsetAsync(t.flags, 0, nint * sizeof(uint32_t), stream));
        if (n > 1) k_karras<<<blocks_for(n - 1), kThreads, 0, stream>>>(keys_sorted, (int) n, t);
        k_refit<<<blocks_for(n), kThreads, 0, stream>>>(gathered, sorted, (int) n, t);
    }
    if (n > 1) k_sah<<<blocks_for(n - 1), kThreads, 0, stream>>>(t, (int) n, sah);

    // SAH-optimal collapse plan (MSK_BVH_COLLAPSE=greedy: the round-1 rule, open the largest child while slots remain)
    {
        const char *csel = getenv("MSK_BVH_COLLAPSE");
        const bool sah_collapse = !(csel && csel[0] == 'g');
        const char *cn = getenv("MSK_BVH_CNODE"), *cp = getenv("MSK_BVH_CPRIM");
        // a node step and a triangle test cost the lockstep traversal about the same number of issue slots per lane
        // (212 warp instructions at 23 lanes against 91 at 9.7: DESIGN.md section 3)
        const float c_node = cn && *cn ? (float) atof(cn) : 1.0f, c_prim = cp && *cp ? (float) atof(cp) : 1.0f;
        if (sah_collapse && n > 1) {
            BVH_CHECK(dalloc(&pl.cost, 7 * nint));
            BVH_CHECK(dalloc(&pl.split, 8 * nint));
            BVH_CHECK(cudaMemsetAsync(t.flags, 0, nint * sizeof(uint32_t), stream));
            const bool dbg = getenv("MSK_DEBUG_SETUP") && atoi(getenv("MSK_DEBUG_SETUP"));
            if (dbg) cudaStreamSynchronize(stream);
            const auto t0 = std::chrono::steady_clock::now();
            k_plan<<<blocks_for(n), kThreads, 0, stream>>>(t, (int) n, pl, c_node, c_prim);
            if (dbg) {
                cudaStreamSynchronize(stream);
                fprintf(stderr, "[msk] k_plan: %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            }
        }
    }

    // level-synchronous collapse
    size_t qcap = node_cap;
    BVH_CHECK(dalloc(&qa, qcap)); BVH_CHECK(dalloc(&qb, qcap));
    WorkItem root{ 0u, 0u };
    BVH_CHECK(cudaMemcpyAsync(qa, &root, sizeof(root), cudaMemcpyHostToDevice, stream));
    uint32_t nin = 1, depth = 0;
    BuildState hst{};
    const bool dbg_collapse = getenv("MSK_DEBUG_SETUP") && atoi(getenv("MSK_DEBUG_SETUP"));
    while (nin) {
        const auto tc0 = std::chrono::steady_clock::now();
        k_collapse<<<(nin + 127) / 128, 128, 0, stream>>>(qa, nin, qb, (uint32_t) qcap, t, (int) n, sorted, gathered, out->nodes,
                                                        (uint32_t) node_cap, out->tris, (uint32_t) tri_cap, st, n == 1 ? 1 : 0, pl);
        BVH_CHECK(cudaMemcpyAsync(&hst, st, sizeof(hst), cudaMemcpyDeviceToHost, stream));
        BVH_CHECK(cudaStreamSynchronize(stream));
        if (hst.overflow) {
            cleanup(); cudaFree(out->nodes); cudaFree(out->tris); *out = BvhResult{};
            if (tri_cap < 24 * node_cap && hst.node_count <= node_cap) // the padded triangle block was too small: worst-case bound
                return bvh_build_impl(stream, d_verts, d_indices, meshes, out, builder, 24 * node_cap);
            return fail(MSK_ERR_OOM, "BVH node pool overflow (n=%zu)", n);
        }
        if (dbg_collapse) fprintf(stderr, "[msk] collapse level %u: %u nodes, %.2f ms\n", depth, nin, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
        nin = hst.queue_count;
        BVH_CHECK(cudaMemsetAsync(&st->queue_count, 0, sizeof(uint32_t), stream));
        std::swap(qa, qb);
        depth++;
    }
    if (hst.tri_emitted != n) { cleanup(); cudaFree(out->nodes); cudaFree(out->tris); *out = BvhResult{};
        return fail(MSK_ERR_CUDA, "BVH collapse emitted %u of %zu triangles", hst.tri_emitted, n); }
    if ((tri_cap - hst.tri_count) * sizeof(float4) * kTriFloat4s > (64u << 20)) { // shrink to fit
        float4 *fit = nullptr;
        BVH_CHECK(dalloc(&fit, (size_t) hst.tri_count * kTriFloat4s));
        BVH_CHECK(cudaMemcpyAsync(fit, out->tris, (size_t) hst.tri_count * kTriFloat4s * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        BVH_CHECK(cudaStreamSynchronize(stream));
        cudaFree(out->tris);
        out->tris = fit;
    }
    float hsah = 0.f;
    float4 rlo, rhi;
    BVH_CHECK(cudaMemcpyAsync(&hsah, sah, sizeof(float), cudaMemcpyDeviceToHost, stream));
    BVH_CHECK(cudaMemcpyAsync(&rlo, t.lo + (n > 1 ? 0 : 0), sizeof(float4), cudaMemcpyDeviceToHost, stream));
    BVH_CHECK(cudaMemcpyAsync(&rhi, t.hi + (n > 1 ? 0 : 0), sizeof(float4), cudaMemcpyDeviceToHost, stream));
    BVH_CHECK(cudaEventRecord(e1, stream));
    BVH_CHECK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&out->ms_build, e0, e1);
    out->nnodes = hst.node_count; out->ntris = n; out->tri_slots = hst.tri_count; out->depth = depth;
    float ra = (rhi.x - rlo.x) * (rhi.y - rlo.y) + (rhi.y - rlo.y) * (rhi.z - rlo.z) + (rhi.z - rlo.z) * (rhi.x - rlo.x);
    out->sah_cost = ra > 0.f ? hsah / ra : 0.f;
    out->lo[0] = rlo.x; out->lo[1] = rlo.y; out->lo[2] = rlo.z;
    out->hi[0] = rhi.x; out->hi[1] = rhi.y; out->hi[2] = rhi.z;
    cleanup();
#undef BVH_CHECK
    return MSK_OK;
}

int bvh_build(cudaStream_t stream, const float4 *d_verts, const uint32_t *d_indices, const std::vector<DMeshInfo> &meshes,
              BvhResult *out, int builder) {
    return bvh_build_impl(stream, d_verts, d_indices, meshes, out, builder, 0);
}

void bvh_free(BvhResult *r) {
    if (!r) return;
    cudaFree(r->nodes); cudaFree(r->tris);
    *r = BvhResult{};
}

} // namespace msk
