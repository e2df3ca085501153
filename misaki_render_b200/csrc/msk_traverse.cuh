// Closest-hit / any-hit traversal of the 8-wide compressed BVH plus the watertight
// ray/triangle test.  Replaces rtcIntersect1 / rtcOccluded1 as called from
// reference src/librender/scene.cpp:216-273.
//
// Node decoding and the octant-ordered bit-stack follow Ylitie, Karras, Laine,
// "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (2017);
// the triangle test is Woop, Benthin, Wald, "Watertight Ray/Triangle Intersection"
// (2013) with the double-precision fallback on zero edge functions.  The accepted
// t range is Embree's: tnear < t <= tfar, no back-face culling, barycentrics
// (u, v) weight vertices 1 and 2 (hit = (1-u-v) v0 + u v1 + v v2).
#pragma once
#include "msk_device.cuh"

namespace msk {

constexpr int kStackSize = 64; // node groups + postponed triangle groups: at most 2 per level of the wide tree

// Lockstep driver: idle lanes are refilled once at least this many lanes of the warp are idle (or none is busy).
// begin() -- reciprocal direction, octant, the nine coefficients of the watertight shear -- is ~100 instructions; run
// for two or three lanes at a time it cost a quarter of the warp's issue slots (round-1 profile: 17.6 of 32 lanes
// active per instruction).  With the shared-memory pool of prepared rays (MSK_RAY_POOL) a refill is five 128-bit
// shared loads and the threshold that measures best drops from 8 to 4-6 (profiles/r03b_ab_raypool_refill.txt).
#ifndef MSK_REFILL_MIN
#define MSK_REFILL_MIN 4
#endif
// Entries of the traversal stack kept in shared memory (per lane; the rest spills to local memory).  0: all local.
#ifndef MSK_SMEM_STACK
#define MSK_SMEM_STACK 8
#endif
// Octant permutation of the inner-child hit byte: 1 = table in shared memory (one LDS), 0 = three delta swaps on the ALU.
#ifndef MSK_PERM_LUT
#define MSK_PERM_LUT 1
#endif
// Lockstep driver: a lane left with nothing but triangles to test ("starved") forces a triangle phase only after it has
// waited this many iterations; meanwhile other lanes' triangles accumulate and the phase runs with more lanes.
#ifndef MSK_TRI_PATIENCE
#define MSK_TRI_PATIENCE 0
#endif
// Lockstep driver: node phases per iteration (the triangle phase, the pop and the refill vote are then paid once per
// MSK_NODE_STEPS node visits; lanes out of node work pop a pending node group in between).
#ifndef MSK_NODE_STEPS
#define MSK_NODE_STEPS 1
#endif
// Lockstep driver: triangles a lane may test per triangle phase.
#ifndef MSK_TRI_REPS
#define MSK_TRI_REPS 1
#endif
// Any-hit queries visit children in slot order instead of front to back (see node_step): C2 any-hit stage 2.48 -> 2.41 ms,
// C3 77.4 -> 72.0 ms, same occlusion results and bit-identical films (profiles/r02x_ab_any_unordered.txt).
#ifndef MSK_ANY_UNORDERED
#define MSK_ANY_UNORDERED 1
#endif
#ifndef MSK_STATIC_MIN_GROUP
#define MSK_STATIC_MIN_GROUP 4 /* 32 disables the adaptive group size of short static queues */
#endif
#ifndef MSK_TRI_THRESHOLD
#define MSK_TRI_THRESHOLD 12
#endif

// Packed FP32 (sm_100: fma / add / mul .f32x2 on a 64-bit register pair, one issue slot for two lanes' worth of IEEE
// operations): the two plane parameters of a child box per axis (near, far) and the two sheared coordinates of a
// triangle vertex (x, y) are computed pairwise.  Every element is the same round-to-nearest operation on the same
// operands as the scalar code, so hits and films are bit-identical; the node test drops from 48 FFMA to 24 FFMA2
// (243 -> 226 SASS instructions per node visit).  MEASURED SLOWER and therefore off: the 64-bit register pairs cost
// 8-50 B of spills under the 80-register cap of the traversal kernels and the FMA pipe was never the bound (ALU pipe:
// PRMT / FMNMX / FSETP); C2 closest hit 5.04 -> 5.07 ms, C3 145.0 -> 147.7 ms (profiles/r03a_ab_f32x2_raypool.txt).
#ifndef MSK_F32X2
#define MSK_F32X2 0
#endif

struct RayHit {
    float t, u, v;
    uint32_t prim, geom;
};

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// Per-ray constants of the watertight test.  Woop et al. permute the axes so that the dominant direction
// component becomes z and shear the translated vertices: Ax = A[kx] - sx A[kz], Ay = A[ky] - sy A[kz],
// Az = sz A[kz].  Here the permutation and the shear are ONE 3x3 matrix with rows
//   mx = e_kx - sx e_kz,   my = e_ky - sy e_kz,   mz = sz e_kz
// applied to (v - o) with plain FMAs.  The first version selected A[kx], A[ky], A[kz] with a dynamic index
// (nine 3-way selects per triangle); ncu attributed 15 % of the kernel's warp instructions to those selects at
// 7 of 32 lanes active (profiles/r01c_ncu_k_intersect.txt).  Watertightness needs the sheared coordinates to be a
// function of (ray, vertex) alone -- they are: the same matrix multiplies every vertex -- and the edge functions
// U, V, W to be exact negatives across a shared edge, which the explicit round-to-nearest mul/sub below keep.
struct WoopRay {
    float mxx, mxy, mxz, myx, myy, myz, mzx, mzy, mzz;
};

__device__ __forceinline__ float rcp_approx(float x) { // MUFU.RCP, 1 ulp; inf for +-0
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ WoopRay woop_setup(float dx, float dy, float dz) {
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    const int kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    const float dkz = kz == 0 ? dx : (kz == 1 ? dy : dz);
    int kx = kz == 2 ? 0 : kz + 1, ky = 3 - kz - kx;
    if (dkz < 0.f) { const int tmp = kx; kx = ky; ky = tmp; } // keep the winding
    const float dkx = kx == 0 ? dx : (kx == 1 ? dy : dz), dky = ky == 0 ? dx : (ky == 1 ? dy : dz);
    // d = 0 (a failed BSDF sample never gets here, but the C ABI accepts any ray): sz = inf, the matrix is NaN,
    // every comparison fails and the ray misses, as with the reference's 0/0
    const float sz = rcp_approx(dkz), sx = dkx * sz, sy = dky * sz;
    const float ez0 = kz == 0 ? 1.f : 0.f, ez1 = kz == 1 ? 1.f : 0.f, ez2 = kz == 2 ? 1.f : 0.f;
    WoopRay w;
    w.mxx = fmaf(-sx, ez0, kx == 0 ? 1.f : 0.f); w.mxy = fmaf(-sx, ez1, kx == 1 ? 1.f : 0.f); w.mxz = fmaf(-sx, ez2, kx == 2 ? 1.f : 0.f);
    w.myx = fmaf(-sy, ez0, ky == 0 ? 1.f : 0.f); w.myy = fmaf(-sy, ez1, ky == 1 ? 1.f : 0.f); w.myz = fmaf(-sy, ez2, ky == 2 ? 1.f : 0.f);
    w.mzx = sz * ez0; w.mzy = sz * ez1; w.mzz = sz * ez2;
    return w;
}

// Returns true and updates (t,u,v) when tnear < t <= tfar.
__device__ __forceinline__ bool woop_intersect(const WoopRay &w, float ox, float oy, float oz, float4 v0, float4 v1, float4 v2,
                                               float tnear, float tfar, float &t_out, float &u_out, float &v_out) {
    const float ax = v0.x - ox, ay = v0.y - oy, az = v0.z - oz;
    const float bx = v1.x - ox, by = v1.y - oy, bz = v1.z - oz;
    const float cx = v2.x - ox, cy = v2.y - oy, cz = v2.z - oz;
#if MSK_F32X2
    float Ax, Ay, Bx, By, Cx, Cy;
    {
        const f32x2 mx = pack2(w.mxx, w.myx), my = pack2(w.mxy, w.myy), mz = pack2(w.mxz, w.myz); // register pairs, no moves
        unpack2(fma2(mx, pack2(ax, ax), fma2(my, pack2(ay, ay), mul2(mz, pack2(az, az)))), Ax, Ay);
        unpack2(fma2(mx, pack2(bx, bx), fma2(my, pack2(by, by), mul2(mz, pack2(bz, bz)))), Bx, By);
        unpack2(fma2(mx, pack2(cx, cx), fma2(my, pack2(cy, cy), mul2(mz, pack2(cz, cz)))), Cx, Cy);
    }
#else
    const float Ax = fmaf(w.mxx, ax, fmaf(w.mxy, ay, w.mxz * az)), Ay = fmaf(w.myx, ax, fmaf(w.myy, ay, w.myz * az));
    const float Bx = fmaf(w.mxx, bx, fmaf(w.mxy, by, w.mxz * bz)), By = fmaf(w.myx, bx, fmaf(w.myy, by, w.myz * bz));
    const float Cx = fmaf(w.mxx, cx, fmaf(w.mxy, cy, w.mxz * cz)), Cy = fmaf(w.myx, cx, fmaf(w.myy, cy, w.myz * cz));
#endif
    // edge functions: explicit round-to-nearest mul/sub (never contracted to FMA) so that a shared edge gives
    // exact negatives in the two triangles
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.f || V == 0.f || W == 0.f) { // rare: fall back to double precision for the sign
        double CxBy = (double) Cx * (double) By, CyBx = (double) Cy * (double) Bx;
        U = (float) (CxBy - CyBx);
        double AxCy = (double) Ax * (double) Cy, AyCx = (double) Ay * (double) Cx;
        V = (float) (AxCy - AyCx);
        double BxAy = (double) Bx * (double) Ay, ByAx = (double) By * (double) Ax;
        W = (float) (BxAy - ByAx);
    }
    if ((U < 0.f || V < 0.f || W < 0.f) && (U > 0.f || V > 0.f || W > 0.f)) return false;
    const float det = U + V + W;
    if (det == 0.f) return false;
    const float Az = fmaf(w.mzx, ax, fmaf(w.mzy, ay, w.mzz * az)), Bz = fmaf(w.mzx, bx, fmaf(w.mzy, by, w.mzz * bz)),
                Cz = fmaf(w.mzx, cx, fmaf(w.mzy, cy, w.mzz * cz));
    const float T  = U * Az + V * Bz + W * Cz;
    const float Ts = det < 0.f ? -T : T, absdet = fabsf(det);
    if (!(Ts > tnear * absdet) || !(Ts <= tfar * absdet)) return false;
    const float rcp = 1.f / det;
    t_out = T * rcp;
    u_out = V * rcp;
    v_out = W * rcp;
    return true;
}

__device__ __forceinline__ float nz(float d) { // keep reciprocal directions finite
    return fabsf(d) < 1e-18f ? copysignf(1e-18f, d) : d;
}

// What a traversal kernel needs of the scene.
struct Accel {
    const float4 *nodes, *tris;
    uint32_t k47; // 0x47000000, read from a kernel parameter so that ptxas cannot fold it (see qfloat)
    uint32_t lut; // shared-memory address of the octant permutation table (perm_lut_init), opaque
};

// Byte J of q as the float 32768 + b, built with ONE byte-permute on the ALU pipe: the byte lands in bits
// 8..15 of 0x47000000 (= 2^15, whose ulp is 2^-8... i.e. mantissa bit 8 has weight 1).  An int->float
// convert (I2F) would go through the quarter-rate XU pipe, which ncu showed 87 % busy in the first version of
// this kernel (profiles/r01a_ncu_k_intersect.txt); 48 of them per node visit made XU the bound.
// PRMT takes ONE immediate.  With both the selector and 0x47000000 known, ptxas put the constant into the
// immediate and re-materialised the selector from a uniform register before every PRMT (48 IMAD.U32 per node
// visit, a fifth of the node test's issue slots in the round-1 SASS); with k47 an opaque kernel parameter the
// selector is the immediate and k47 stays in one register.
template <int J> __device__ __forceinline__ float qfloat(uint32_t q, uint32_t k47) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(q), "r"(k47), "n"(0x7504 | (J << 4)));
    return __uint_as_float(r);
}

// Octant permutation table: perm[o][h] moves bit s of the byte h to bit s ^ o.  The inner-child hit bits of a
// node are produced in slot order (static bit positions, see node_step) and must be visited in the order
// slot ^ octinv descending; one LDS replaces a dozen ALU instructions per node visit on the pipe that bounds
// the kernel.
constexpr int kPermLutBytes = 8 * 256;
// A shared-memory address is cheap to re-derive in ptxas's cost model (S2UR SR_CgaCtaId, UMOV, UIADD3, ULEA, LEA), so it
// re-materialised the table and stack addresses at every use instead of holding two registers: 13 warp instructions per
// iteration of the lockstep loop, 2.7 % of the kernel (profiles/r02d_ncu_k_intersect.txt).  Passing the value through
// a self-shuffle (once per kernel) makes it a value ptxas cannot re-derive.
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
    return __shfl_sync(0xffffffffu, v, threadIdx.x & 31u);
}
__device__ __forceinline__ uint32_t perm_byte(uint32_t h, uint32_t o) {
    if (o & 1u) h = ((h & 0x55u) << 1) | ((h >> 1) & 0x55u);
    if (o & 2u) h = ((h & 0x33u) << 2) | ((h >> 2) & 0x33u);
    if (o & 4u) h = ((h & 0x0fu) << 4) | (h >> 4);
    return h;
}
__device__ __forceinline__ uint32_t perm_lut_init(uint8_t *lut) { // call from every thread of the block
    for (uint32_t i = threadIdx.x; i < (uint32_t) kPermLutBytes; i += blockDim.x) lut[i] = (uint8_t) perm_byte(i & 0xffu, i >> 8);
    __syncthreads();
    return opaque((uint32_t) __cvta_generic_to_shared(lut));
}

// Traversal stack: the first MSK_SMEM_STACK entries of every lane live in shared memory (entry k of the block's
// lanes is contiguous: conflict-free 64-bit accesses), deeper entries in local memory.  The round-1 kernels kept
// all 64 entries (512 B per thread) in local memory, i.e. in the same L1 the node fetches go through.
constexpr int kSmemStack = MSK_SMEM_STACK;
constexpr int kTravThreads = 128; // block size of every kernel that traverses
constexpr int kLocalStack = kStackSize - kSmemStack;
struct TravStack {
    uint2 *local;  // this lane's local-memory entries (kLocalStack of them)
    uint32_t smem; // shared address of this lane's entry 0 (opaque: held in a register, see above)
    __device__ __forceinline__ TravStack(uint2 *local_entries, uint2 *shared_entries)
        : local(local_entries), smem(kSmemStack ? opaque((uint32_t) __cvta_generic_to_shared(shared_entries + threadIdx.x)) : 0u) {}
    __device__ __forceinline__ void store(int i, uint2 v) {
        if (kSmemStack && i < kSmemStack)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(smem + (uint32_t) i * (kTravThreads * 8u)), "r"(v.x), "r"(v.y) : "memory");
        else local[i - kSmemStack] = v;
    }
    __device__ __forceinline__ uint2 load(int i) const {
        uint2 v;
        if (kSmemStack && i < kSmemStack)
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem + (uint32_t) i * (kTravThreads * 8u)) : "memory");
        else v = local[i - kSmemStack];
        return v;
    }
};
// Lockstep driver: rays are prepared (begin(): reciprocal direction, octant, watertight shear -- ~100 instructions) by ALL
// 32 lanes of a warp at once, 32 rays at a time, into a per-warp pool in shared memory; an idle lane then takes a prepared
// ray with five 128-bit shared loads.  0: every lane prefetches its own next ray and runs begin() when it refills (with
// the 8-12 lanes that are idle at that moment).  The raw rays of the NEXT fill sit in registers (their loads are in
// flight while the pool is consumed).  Which lane traces which ray changes nothing in a ray's result: films are
// bit-identical (profiles/r03a_film_hash_bit_identity.txt).  C2 1494 -> 1538 Mpaths/s (closest hit 5.05 -> 4.85 ms, any
// hit 2.37 -> 2.23), C3 922 -> 938, C5 sweep 17.54 -> 16.52 ms (incoherent closest hit 2.94 -> 3.10 Grays/s), fog
// workload 360 -> 370 (profiles/r03b_ab_raypool_refill.txt).  10 KB of shared memory per CTA.
#ifndef MSK_RAY_POOL
#define MSK_RAY_POOL 1
#endif
constexpr int kPoolSlotFloat4s = 5; // 80 B per prepared ray: 128-bit accesses at this stride are bank-conflict-free
#define MSK_TRAV_LOCAL_STACK uint2 msk_local_stack[msk::kLocalStack > 0 ? msk::kLocalStack : 1]
#define MSK_TRAV_SHARED                                                        \
    __shared__ uint8_t msk_s_perm[msk::kPermLutBytes];                           \
    __shared__ uint2 msk_s_stack[msk::kSmemStack ? msk::kSmemStack * msk::kTravThreads : 1]; \
    __shared__ float4 msk_s_pool[MSK_RAY_POOL ? msk::kPoolSlotFloat4s * msk::kTravThreads : 1]
#define MSK_TRAV_SMEM msk_s_stack, msk_s_pool

// Per-lane traversal state.  A lane owns one ray at a time; the warp refills idle lanes from the queue
// (dynamic fetch, Aila & Laine 2009) instead of waiting for its slowest ray.
struct Traversal {
    float ox, oy, oz, tmin, tmax;
    float tfar0;   // the ray's own maxt
    float idx, idy, idz;
    uint32_t octinv;
    WoopRay wr;
    uint2 ngroup, tgroup;
    int sp;
    RayHit hit;
    bool found;
    uint32_t cnt_nodes, cnt_tris;
    uint32_t tag; // a word the IO adaptor attaches to the ray when it is loaded (IO::tag) and reads back in commit()

    __device__ __forceinline__ void begin(float4 ro, float4 rd) {
        ox = ro.x; oy = ro.y; oz = ro.z; tmin = ro.w;
        tmax = rd.w; tfar0 = rd.w;
        // MUFU.RCP (1 ulp) instead of three IEEE divisions: the slab test already carries float rounding of the
        // same size, and begin() runs with few lanes active, so every instruction here costs a whole issue slot
        idx = rcp_approx(nz(rd.x)); idy = rcp_approx(nz(rd.y)); idz = rcp_approx(nz(rd.z));
        // signs of the clamped direction (-0 counts as negative)
        const uint32_t oct = (idx < 0.f ? 4u : 0u) | (idy < 0.f ? 2u : 0u) | (idz < 0.f ? 1u : 0u);
        octinv = 7u - oct;
        wr = woop_setup(rd.x, rd.y, rd.z);
        reset();
    }
    __device__ __forceinline__ void reset() {
        ngroup = make_uint2(0u, 0x80000000u); // root: "child slot 7 of a virtual parent at base 0"
        tgroup = make_uint2(0u, 0u);
        sp = 0;
        found = false;
        hit.t = __int_as_float(0x7f800000); hit.u = 0.f; hit.v = 0.f; hit.prim = 0xffffffffu; hit.geom = 0xffffffffu;
        cnt_nodes = 0; cnt_tris = 0;
    }
    // the per-ray constants that begin() derives, as one 80-byte record (the lockstep driver's shared-memory ray pool)
    __device__ __forceinline__ void save(float4 *slot, uint32_t q) const {
        slot[0] = make_float4(ox, oy, oz, tmin);
        slot[1] = make_float4(idx, idy, idz, tfar0);
        slot[2] = make_float4(wr.mxx, wr.mxy, wr.mxz, wr.myx);
        slot[3] = make_float4(wr.myy, wr.myz, wr.mzx, wr.mzy);
        slot[4] = make_float4(wr.mzz, __uint_as_float(q), __uint_as_float(octinv), __uint_as_float(tag));
    }
    __device__ __forceinline__ uint32_t restore(const float4 *slot) {
        const float4 a = slot[0], b = slot[1], c = slot[2], d = slot[3], e = slot[4];
        ox = a.x; oy = a.y; oz = a.z; tmin = a.w;
        idx = b.x; idy = b.y; idz = b.z; tmax = b.w; tfar0 = b.w;
        wr.mxx = c.x; wr.mxy = c.y; wr.mxz = c.z; wr.myx = c.w;
        wr.myy = d.x; wr.myz = d.y; wr.mzx = d.z; wr.mzy = d.w;
        wr.mzz = e.x; octinv = __float_as_uint(e.z); tag = __float_as_uint(e.w);
        reset();
        return __float_as_uint(e.y);
    }
    // Scene::ray_intersect / ray_test report a hit iff Embree moved tfar (scene.cpp:234,272): a hit at
    // exactly t == maxt reads as a miss
    __device__ __forceinline__ bool is_hit() const { return found && hit.t != tfar0; }
};

constexpr int kTriThreshold = MSK_TRI_THRESHOLD; // run a triangle phase once this many lanes have triangles pending
constexpr int kRefillMin    = MSK_REFILL_MIN;
#ifndef MSK_REFILL_MIN_ANY
#define MSK_REFILL_MIN_ANY MSK_REFILL_MIN /* any-hit queries end early and refill more often: their own threshold */
#endif
constexpr int kRefillMinAny = MSK_REFILL_MIN_ANY;

// Visit the next pending inner node of s.ngroup: pushes what remains of the group, tests the node's 8
// quantised child boxes, leaves the children hit in s.ngroup and the triangles hit in s.tgroup.
//
// Hit bits have STATIC positions (node layout in msk_device.cuh): the child in slot j raises bit 24 + j and bits
// 3j..3j+2 with one predicated OR of an immediate, and one AND with the node's `valid` word keeps the inner bit of
// inner children and the bits of the triangles a leaf child really holds.  (Round 1 decoded a bit position and a
// triangle count per child from meta bytes: five ALU instructions per child on the pipe that bounds this kernel.)
// The inner byte is then permuted into traversal order slot ^ octinv.
// ORDERED = false (any-hit queries under MSK_ANY_UNORDERED): children are visited in slot order -- no octant permutation of
// the hit byte, no slot ^ octinv; an occlusion query does not care which occluder it finds.
template <bool STATS, bool ORDERED = true>
__device__ __forceinline__ void node_step(const Accel &ac, Traversal &s, TravStack &stack) {
    const bool negx = s.idx < 0.f, negy = s.idy < 0.f, negz = s.idz < 0.f;
    const uint32_t hits  = s.ngroup.y;
    const uint32_t imask = s.ngroup.y & 0xffu;
    const uint32_t bit   = 31u - __clz(hits);
    s.ngroup.y &= ~(1u << bit);
    if (s.ngroup.y > 0x00ffffffu) stack.store(s.sp++, s.ngroup);
    const uint32_t slot = ORDERED ? (bit - 24u) ^ s.octinv : bit - 24u;
    const uint32_t rel  = __popc(imask & ~(0xffffffffu << slot));
    const uint32_t node = s.ngroup.x + rel;
    const float4 *np = ac.nodes + (size_t) node * kNodeFloat4s;
    const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    if (STATS) s.cnt_nodes++;
    const uint32_t e = __float_as_uint(n0.w);
    const float adx = __uint_as_float((e & 0xffu) << 23) * s.idx, ady = __uint_as_float(((e >> 8) & 0xffu) << 23) * s.idy,
                adz = __uint_as_float(((e >> 16) & 0xffu) << 23) * s.idz;
    // plane j at parameter t = (32768 + q_j) * ad + (ao - 32768 * ad).  The product carries a rounding
    // error of up to 2^-9 cell, so the near planes move 2^-8 cell towards the ray origin and the far
    // planes 2^-8 cell away from it: the decoded boxes are supersets of the stored ones.
    const float aox = fmaf(-32768.f, adx, (n0.x - s.ox) * s.idx), aoy = fmaf(-32768.f, ady, (n0.y - s.oy) * s.idy),
                aoz = fmaf(-32768.f, adz, (n0.z - s.oz) * s.idz);
    const float ex = fabsf(adx) * 0.00390625f, ey = fabsf(ady) * 0.00390625f, ez = fabsf(adz) * 0.00390625f;
#if MSK_F32X2
    const f32x2 nfx = add2(pack2(aox, aox), pack2(-ex, ex)), nfy = add2(pack2(aoy, aoy), pack2(-ey, ey)), nfz = add2(pack2(aoz, aoz), pack2(-ez, ez));
    const f32x2 adx2 = pack2(adx, adx), ady2 = pack2(ady, ady), adz2 = pack2(adz, adz);
#else
    const float nox = aox - ex, fox = aox + ex, noy = aoy - ey, foy = aoy + ey, noz = aoz - ez, foz = aoz + ez;
#endif
    const uint32_t k47 = ac.k47;
    s.ngroup.x = __float_as_uint(n1.x);
    s.tgroup.x = __float_as_uint(n1.y);
    uint32_t hitmask = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t qlox = __float_as_uint(h ? n2.y : n2.x), qloy = __float_as_uint(h ? n2.w : n2.z),
                       qloz = __float_as_uint(h ? n3.y : n3.x), qhix = __float_as_uint(h ? n3.w : n3.z),
                       qhiy = __float_as_uint(h ? n4.y : n4.x), qhiz = __float_as_uint(h ? n4.w : n4.z);
        // near / far planes by ray direction sign
        const uint32_t nx = negx ? qhix : qlox, fx = negx ? qlox : qhix;
        const uint32_t ny = negy ? qhiy : qloy, fy = negy ? qloy : qhiy;
        const uint32_t nzq = negz ? qhiz : qloz, fz = negz ? qloz : qhiz;
#if MSK_F32X2
#define MSK_CHILD_PLANES(J)                                                                                       \
        float t0x, t1x, t0y, t1y, t0z, t1z;                                                                       \
        unpack2(fma2(pack2(qfloat<J>(nx, k47), qfloat<J>(fx, k47)), adx2, nfx), t0x, t1x);                        \
        unpack2(fma2(pack2(qfloat<J>(ny, k47), qfloat<J>(fy, k47)), ady2, nfy), t0y, t1y);                        \
        unpack2(fma2(pack2(qfloat<J>(nzq, k47), qfloat<J>(fz, k47)), adz2, nfz), t0z, t1z);
#else
#define MSK_CHILD_PLANES(J)                                                                                       \
        const float t0x = fmaf(qfloat<J>(nx, k47), adx, nox), t1x = fmaf(qfloat<J>(fx, k47), adx, fox);           \
        const float t0y = fmaf(qfloat<J>(ny, k47), ady, noy), t1y = fmaf(qfloat<J>(fy, k47), ady, foy);           \
        const float t0z = fmaf(qfloat<J>(nzq, k47), adz, noz), t1z = fmaf(qfloat<J>(fz, k47), adz, foz);
#endif
#define MSK_CHILD(J)                                                                                              \
    {                                                                                                             \
        MSK_CHILD_PLANES(J)                                                                                       \
        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, s.tmin));                                              \
        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, s.tmax));                                              \
        if (tn <= tf) hitmask |= (1u << (24 + 4 * h + J)) | (7u << (3 * (4 * h + J)));                            \
    }
        MSK_CHILD(0) MSK_CHILD(1) MSK_CHILD(2) MSK_CHILD(3)
#undef MSK_CHILD
#undef MSK_CHILD_PLANES
    }
    hitmask &= __float_as_uint(n1.z); // valid: imask << 24 | bit 3j + k for triangle k of the leaf in slot j
    uint32_t inner;
    if (!ORDERED) inner = hitmask >> 24;
    else
#if MSK_PERM_LUT
    asm("ld.shared.u8 %0, [%1];" : "=r"(inner) : "r"(ac.lut + (s.octinv << 8) + (hitmask >> 24)));
#else
    inner = perm_byte(hitmask >> 24, s.octinv);
#endif
    s.ngroup.y = (inner << 24) | (e >> 24);
    s.tgroup.y = hitmask & 0x00ffffffu;
}

// Test the highest pending triangle of s.tgroup.  Returns true when it is hit (tmax shrinks).
template <bool STATS>
__device__ __forceinline__ bool tri_step(const Accel &ac, Traversal &s) {
    const uint32_t k = 31u - __clz(s.tgroup.y);
    s.tgroup.y &= ~(1u << k);
    const float4 *tp = ac.tris + (size_t) (s.tgroup.x + k) * kTriFloat4s; // slot 3j + k of the node's triangle block (mod 2^32)
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    if (STATS) s.cnt_tris++;
    float t, u, v;
    if (woop_intersect(s.wr, s.ox, s.oy, s.oz, v0, v1, v2, s.tmin, s.tmax, t, u, v)) {
        s.tmax = t;
        s.hit.t = t; s.hit.u = u; s.hit.v = v;
        s.hit.prim = __float_as_uint(v0.w); s.hit.geom = __float_as_uint(v1.w);
        s.found = true;
        return true;
    }
    return false;
}

// One ray to completion, classic while-while order.
template <bool ANY, bool STATS>
__device__ __forceinline__ void traverse_one(const Accel &ac, Traversal &s, TravStack &stack) {
    for (;;) {
        if (s.ngroup.y > 0x00ffffffu) node_step<STATS, !(ANY && MSK_ANY_UNORDERED)>(ac, s, stack);
        else { s.tgroup = s.ngroup; s.ngroup = make_uint2(0u, 0u); }
        bool stop = false;
        while (s.tgroup.y) {
            if (tri_step<STATS>(ac, s) && ANY) { stop = true; break; }
        }
        if (stop) break;
        if (s.ngroup.y <= 0x00ffffffu) {
            if (s.sp == 0) break;
            s.ngroup = stack.load(--s.sp);
        }
    }
}

// ---- Packet traversal of a coherent warp (camera rays: 32 samples of one pixel share the origin and almost the direction).
// The warp walks the tree ONCE: node groups, triangle groups and the stack are warp-uniform (every lane keeps the same copy
// in its own Traversal / stack slots), child j of a node is tested against the bounds of the whole packet by lane j & 7 --
// interval arithmetic over the origins [olo, ohi] and reciprocal directions [rlo, rhi] of the 32 rays -- and every lane
// tests every triangle of the leaves the packet enters with its own ray.  The per-ray node test (220 warp instructions per
// visit, each lane finding what its neighbours find) becomes ~100, and nothing diverges.
// Exactness: a child is entered whenever ANY ray of the packet could enter it (bounds are outward-rounded and padded), children
// are visited in the same octant order as by a single ray (all rays of a packet share the octant, otherwise the caller falls
// back to per-ray traversal), and a ray accepts exactly the triangle hits it would accept alone (tnear < t <= its own tfar),
// in the same order: hits and films are bit-identical to per-ray traversal.
struct PacketBounds {
    // per axis: reciprocal directions [rlo, rhi] (one sign), the origin bound that minimises the entry parameter of the
    // near plane (on) and the one that maximises the exit parameter of the far plane (of), sn = -1 / +1 for positive /
    // negative directions (sign of the padding of the near side), opad = 1e-6 max |origin|, sg = -+ 1/64 cell
    float rlo[3], rhi[3], on[3], of[3], sn[3], opad[3], sg[3];
    float tmin, tmax;
};
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
__device__ __forceinline__ void packet_axis(PacketBounds &pb, int a, bool valid, float o, float r) {
    const float inf = __int_as_float(0x7f800000);
    const float olo = warp_min(valid ? o : inf), ohi = warp_max(valid ? o : -inf);
    pb.rlo[a] = warp_min(valid ? r : inf); pb.rhi[a] = warp_max(valid ? r : -inf);
    const bool neg = pb.rhi[a] < 0.f; // all rays of the packet share the sign (checked by the caller through the octant)
    pb.on[a] = neg ? olo : ohi; pb.of[a] = neg ? ohi : olo;
    pb.sn[a] = neg ? 1.f : -1.f;
    pb.sg[a] = neg ? 0.015625f : -0.015625f;
    pb.opad[a] = 1e-6f * fmaxf(fabsf(olo), fabsf(ohi));
}
// One axis of the packet-vs-child test.  near4 / far4: the two words holding the 8 quantised near / far planes of the node
// (chosen by the packet's direction sign), j: this lane's child.  Returns a lower bound of the parameter at which any ray of
// the packet can pass the child's near plane and an upper bound of the parameter at which any ray can still be inside
// its far plane; planes are moved 1/64 cell outwards (the per-ray test pads by 1/256 cell and rounds within 1/512) and
// plane - origin is padded by 8 ulp of the coordinates involved.
__device__ __forceinline__ void packet_child_axis(const PacketBounds &pb, int a, uint32_t near_lo4, uint32_t near_hi4, uint32_t far_lo4,
                                                  uint32_t far_hi4, uint32_t j, float p, float c, float &tn, float &tf) {
    const float qn = (float) (__byte_perm(near_lo4, near_hi4, j) & 0xffu), qf = (float) (__byte_perm(far_lo4, far_hi4, j) & 0xffu);
    const float pn = fmaf(qn, c, fmaf(c, pb.sg[a], p)), pf = fmaf(qf, c, fmaf(c, -pb.sg[a], p));
    const float pad = fmaf(c, 256e-6f, fmaf(fabsf(p), 1e-6f, pb.opad[a]));
    const float an = fmaf(pb.sn[a], pad, pn - pb.on[a]), af = fmaf(-pb.sn[a], pad, pf - pb.of[a]);
    tn = fminf(an * pb.rlo[a], an * pb.rhi[a]);
    tf = fmaxf(af * pb.rlo[a], af * pb.rhi[a]);
}

template <bool STATS, bool ORDERED>
__device__ __forceinline__ void node_step_packet(const Accel &ac, Traversal &s, TravStack &stack, const PacketBounds &pb) {
    const uint32_t hits  = s.ngroup.y;
    const uint32_t imask = s.ngroup.y & 0xffu;
    const uint32_t bit   = 31u - __clz(hits);
    s.ngroup.y &= ~(1u << bit);
    if (s.ngroup.y > 0x00ffffffu) stack.store(s.sp++, s.ngroup);
    const uint32_t slot = ORDERED ? (bit - 24u) ^ s.octinv : bit - 24u;
    const uint32_t rel  = __popc(imask & ~(0xffffffffu << slot));
    const uint32_t node = s.ngroup.x + rel;
    const float4 *np = ac.nodes + (size_t) node * kNodeFloat4s;
    const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    if (STATS) s.cnt_nodes++;
    const uint32_t e = __float_as_uint(n0.w);
    const float cx = __uint_as_float((e & 0xffu) << 23), cy = __uint_as_float(((e >> 8) & 0xffu) << 23), cz = __uint_as_float(((e >> 16) & 0xffu) << 23);
    const uint32_t j = threadIdx.x & 7u; // this lane's child slot (lanes 8..31 repeat the work of lanes 0..7)
    const bool negx = !(s.octinv & 4u), negy = !(s.octinv & 2u), negz = !(s.octinv & 1u);
    const uint32_t lox0 = __float_as_uint(n2.x), lox1 = __float_as_uint(n2.y), loy0 = __float_as_uint(n2.z), loy1 = __float_as_uint(n2.w),
                   loz0 = __float_as_uint(n3.x), loz1 = __float_as_uint(n3.y), hix0 = __float_as_uint(n3.z), hix1 = __float_as_uint(n3.w),
                   hiy0 = __float_as_uint(n4.x), hiy1 = __float_as_uint(n4.y), hiz0 = __float_as_uint(n4.z), hiz1 = __float_as_uint(n4.w);
    float tnx, tfx, tny, tfy, tnz, tfz;
    packet_child_axis(pb, 0, negx ? hix0 : lox0, negx ? hix1 : lox1, negx ? lox0 : hix0, negx ? lox1 : hix1, j, n0.x, cx, tnx, tfx);
    packet_child_axis(pb, 1, negy ? hiy0 : loy0, negy ? hiy1 : loy1, negy ? loy0 : hiy0, negy ? loy1 : hiy1, j, n0.y, cy, tny, tfy);
    packet_child_axis(pb, 2, negz ? hiz0 : loz0, negz ? hiz1 : loz1, negz ? loz0 : hiz0, negz ? loz1 : hiz1, j, n0.z, cz, tnz, tfz);
    // some ray may enter the box only if the earliest possible entry is not after the latest possible exit (both moved
    // outwards by 4e-6 relative for the rounding of the products; x * (1 +- e) keeps an infinity an infinity)
    float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, pb.tmin));
    float tf = fminf(fminf(tfx, tfy), fminf(tfz, pb.tmax));
    tn = fminf(tn * 0.999996f, tn * 1.000004f);
    tf = fmaxf(tf * 0.999996f, tf * 1.000004f);
    const uint32_t w = tn <= tf ? (1u << (24u + j)) | (7u << (3u * j)) : 0u;
    uint32_t hitmask = __reduce_or_sync(0xffffffffu, w);
    s.ngroup.x = __float_as_uint(n1.x);
    s.tgroup.x = __float_as_uint(n1.y);
    hitmask &= __float_as_uint(n1.z);
    uint32_t inner;
    if (!ORDERED) inner = hitmask >> 24;
    else
#if MSK_PERM_LUT
    asm("ld.shared.u8 %0, [%1];" : "=r"(inner) : "r"(ac.lut + (s.octinv << 8) + (hitmask >> 24)));
#else
    inner = perm_byte(hitmask >> 24, s.octinv);
#endif
    s.ngroup.y = (inner << 24) | (e >> 24);
    s.tgroup.y = hitmask & 0x00ffffffu;
}

// All 32 lanes call this converged; `valid` lanes hold a ray that begin() has set up.  Returns false -- having done nothing --
// when the rays do not share one direction octant (the caller then traverses them one by one).
template <bool ANY, bool STATS>
__device__ __forceinline__ bool traverse_packet(const Accel &ac, Traversal &s, TravStack &stack, bool valid) {
    const uint32_t vm = __ballot_sync(0xffffffffu, valid);
    if (vm == 0u) return true;
    const uint32_t oct0 = __shfl_sync(0xffffffffu, s.octinv, __ffs(vm) - 1);
    if (__any_sync(0xffffffffu, valid && s.octinv != oct0)) return false;
    const float inf = __int_as_float(0x7f800000);
    PacketBounds pb;
    packet_axis(pb, 0, valid, s.ox, s.idx);
    packet_axis(pb, 1, valid, s.oy, s.idy);
    packet_axis(pb, 2, valid, s.oz, s.idz);
    pb.tmin = warp_min(valid ? s.tmin : inf);
    pb.tmax = warp_max(valid ? s.tmax : -inf);
    if (!valid) { s.octinv = oct0; s.reset(); } // warp-uniform control state in every lane
    bool active = valid; // this lane still looks for a hit
    for (;;) {
        if (s.ngroup.y > 0x00ffffffu) node_step_packet<STATS, !(ANY && MSK_ANY_UNORDERED)>(ac, s, stack, pb);
        else { s.tgroup = s.ngroup; s.ngroup = make_uint2(0u, 0u); }
        bool hit_any = false;
        while (s.tgroup.y) { // warp-uniform
            const uint32_t k = 31u - __clz(s.tgroup.y);
            s.tgroup.y &= ~(1u << k);
            const float4 *tp = ac.tris + (size_t) (s.tgroup.x + k) * kTriFloat4s;
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            float t, u, v;
            if (active) {
                if (STATS) s.cnt_tris++;
                if (woop_intersect(s.wr, s.ox, s.oy, s.oz, v0, v1, v2, s.tmin, s.tmax, t, u, v)) {
                    s.tmax = t;
                    s.hit.t = t; s.hit.u = u; s.hit.v = v;
                    s.hit.prim = __float_as_uint(v0.w); s.hit.geom = __float_as_uint(v1.w);
                    s.found = true;
                    hit_any = true;
                    if (ANY) active = false;
                }
            }
        }
        if (ANY) { if (__ballot_sync(0xffffffffu, active) == 0u) break; }
        else if (__any_sync(0xffffffffu, hit_any)) pb.tmax = warp_max(valid ? s.tmax : -inf); // closest hits only shrink the packet
        if (s.ngroup.y <= 0x00ffffffu) {
            if (s.sp == 0) break;
            s.ngroup = stack.load(--s.sp);
        }
    }
    return true;
}

// (used where rays are not queued: the per-path tail kernel)
template <bool ANY, bool STATS>
__device__ __forceinline__ bool traverse(const Accel &ac, TravStack &stack, float ox, float oy, float oz, float dx, float dy, float dz,
                                         float tmin, float tmax, RayHit &hit, uint32_t *nnodes = nullptr, uint32_t *ntris = nullptr) {
    Traversal s;
    s.begin(make_float4(ox, oy, oz, tmin), make_float4(dx, dy, dz, tmax));
    traverse_one<ANY, STATS>(ac, s, stack);
    if (s.is_hit()) hit = s.hit;
    if (STATS) { *nnodes = s.cnt_nodes; *ntris = s.cnt_tris; }
    return s.is_hit();
}

// Persistent-warp driver: every lane pulls rays from the queue [0, n) until it is empty.
//   io.load(q, ro, rd)                 fetch ray q
//   io.tag(ro)                         called once per loaded ray before it is set up: may unpack a word the producer stored in
//                                      the ray record (and restore the field it borrowed); the word is Traversal::tag in commit()
//   io.commit(have, q, s)              called by ALL 32 lanes converged; lanes with `have` deliver the result
//                                      of their finished ray q (s.is_hit(), s.hit, s.cnt_*)
//
// The warp advances in LOCKSTEP phases, each entered by a vote, so that the two expensive code blocks run
// with as many lanes as possible (ncu on the first while-while version: 8 of 32 lanes active in the node
// test, 3 of 32 in the triangle test -- profiles/r01a_ncu_k_intersect.txt):
//   refill    once kRefillMin lanes are idle they commit their results and take the next rays (indices come from a
//             per-warp chunk reserved with one atomic; the ray itself was prefetched into registers while the
//             previous ray was in flight)
//   node      every lane with a pending inner node visits one (8 child boxes)
//   triangle  runs only when >= kTriThreshold lanes hold pending triangles or a lane has nothing else to do;
//             every lane with pending triangles tests one.  Triangles found meanwhile wait on the stack.
//   pop       lanes out of work pop the stack or finish
constexpr uint32_t kChunk = 128; // ray indices reserved per atomic

#ifndef MSK_TRAVERSAL_MODE
#define MSK_TRAVERSAL_MODE 2 /* 0 static, 1 lockstep, 2 hybrid */
#endif

// Static assignment: the warp takes 32 consecutive rays and every lane runs its ray to completion.
template <bool ANY, bool STATS, bool PACKETS = false, typename IO>
__device__ __forceinline__ void trace_queue_static(const Accel &ac, uint32_t n, uint32_t *cursor, IO &io, TravStack &stack) {
    Traversal s;
    const uint32_t lane = threadIdx.x & 31u;
    // A queue shorter than one ray per lane of the grid is latency-bound: the launch lasts as long as its slowest
    // warp, and a warp of 32 divergent rays serialises their node and triangle steps (the tail bounces of C2 took
    // a flat ~90 us each whatever their length, profiles/r01d_launches_c2.csv).  Spread such a queue over all
    // warps instead: each warp takes `group` consecutive rays (a power of two, >= MSK_STATIC_MIN_GROUP).
    uint32_t group = 32u;
#if MSK_STATIC_MIN_GROUP < 32
    {
        const uint32_t nwarps = gridDim.x * (blockDim.x / 32u);
        const uint32_t per = (n + nwarps - 1u) / nwarps;
        if (per < 32u) { group = MSK_STATIC_MIN_GROUP; while (group < per) group <<= 1; }
    }
#endif
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, group);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t q = base + lane;
        const bool valid = lane < group && q < n;
        if (valid) {
            float4 ro, rd;
            io.load(q, ro, rd);
            s.tag = io.tag(ro);
            s.begin(ro, rd);
        }
        __syncwarp();
        if (!PACKETS || !traverse_packet<ANY, STATS>(ac, s, stack, valid)) {
            if (valid) traverse_one<ANY, STATS>(ac, s, stack);
        }
        __syncwarp();
        io.commit(valid, q, s);
    }
}

// PACKETS: the kernel is compiled for packet traversal only (every warp of the queue is a packet: camera rays); its own
// kernel, so that the register allocation of the lockstep loop is not disturbed by the packet code (and vice versa).
template <bool ANY, bool STATS, bool PACKETS = false, typename IO>
__device__ __forceinline__ void trace_queue(const Accel &ac, uint2 *shared_stack, float4 *shared_pool, uint32_t n, uint32_t *cursor, IO &io, int coherent) {
    if (PACKETS) {
        MSK_TRAV_LOCAL_STACK;
        TravStack stack(msk_local_stack, shared_stack);
        trace_queue_static<ANY, STATS, true>(ac, n, cursor, io, stack);
        return;
    }
    // Coherent queues (camera rays: neighbouring lanes follow the same nodes, so a static warp of 32 runs
    // converged and its node fetches coalesce) and queues too small to fill the machine (latency-bound: more,
    // shorter warps win) use the static assignment; incoherent bulk queues use the lockstep phases.
#if MSK_TRAVERSAL_MODE == 0
    const bool use_static = true;
#elif MSK_TRAVERSAL_MODE == 1
    const bool use_static = false;
#else
    const bool use_static = coherent != 0 || n < gridDim.x * (blockDim.x / 32u) * 64u;
#endif
    MSK_TRAV_LOCAL_STACK;
    TravStack stack(msk_local_stack, shared_stack);
    if (use_static) {
        trace_queue_static<ANY, STATS>(ac, n, cursor, io, stack);
        return;
    }
    Traversal s;
    s.ngroup = make_uint2(0u, 0u); s.tgroup = make_uint2(0u, 0u); s.sp = 0;
    const uint32_t lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    bool busy = false, have = false;
    uint32_t q = 0;
    int waited = 0; // iterations this lane has been starved (MSK_TRI_PATIENCE)
    // prefetched next ray of this lane
    float4 nro = make_float4(0, 0, 0, 0), nrd = make_float4(0, 0, 0, 0);
    uint32_t nq = 0xffffffffu;             // 0xffffffff: none
#if MSK_RAY_POOL
    // The warp's pool of prepared rays [pool_head, pool_count) in shared memory, and in registers (nq, nro, nrd) the 32 raw
    // rays the next fill will prepare -- their loads are in flight while the pool is being consumed.
    float4 *const pool_w = shared_pool + (threadIdx.x & ~31u) * kPoolSlotFloat4s;
    uint32_t pool_head = 0, pool_count = 0; // warp-uniform
    bool more = true;                       // warp-uniform: the global queue may still hold rays
    auto fetch_raw = [&]() {                // all lanes
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) { more = false; nq = 0xffffffffu; return; }
        nq = n - base > lane ? base + lane : 0xffffffffu;
        if (nq != 0xffffffffu) io.load(nq, nro, nrd);
    };
    auto fill = [&]() { // all lanes, pool empty: prepare the raw rays (a prefix of the lanes holds one), fetch the next 32
        pool_head = 0;
        pool_count = (uint32_t) __popc(__ballot_sync(0xffffffffu, nq != 0xffffffffu));
        __syncwarp(); // every lane has finished reading its slot of the previous fill
        if (nq != 0xffffffffu) {
            Traversal p;
            p.tag = io.tag(nro);
            p.begin(nro, nrd);
            p.save(pool_w + lane * kPoolSlotFloat4s, nq);
        }
        __syncwarp();
        if (more) fetch_raw(); else nq = 0xffffffffu;
    };
    fetch_raw();
    for (;;) {
        // ---- refill
        const uint32_t idle = __ballot_sync(0xffffffffu, !busy);
        if (__popc(idle) >= (ANY ? kRefillMinAny : kRefillMin)) {
            io.commit(have, q, s);
            have = false;
            uint32_t want = idle;
            for (;;) {
                if (pool_head == pool_count) {
                    if (__ballot_sync(0xffffffffu, nq != 0xffffffffu) == 0u) break; // queue exhausted
                    fill();
                }
                const uint32_t avail = pool_count - pool_head;
                const uint32_t rank = (uint32_t) __popc(want & below);
                const bool take = ((want >> lane) & 1u) && rank < avail;
                if (take) { q = s.restore(pool_w + (pool_head + rank) * kPoolSlotFloat4s); busy = true; }
                pool_head += min((uint32_t) __popc(want), avail);
                want &= ~__ballot_sync(0xffffffffu, take);
                if (!want) break;
            }
            if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        }
#else
    uint32_t chunk_next = 0, chunk_end = 0; // warp-uniform
    bool more = true;                       // warp-uniform: the global queue may still hold rays

    // take indices for the lanes in `want` (warp-uniform mask) from the warp's chunk; returns this lane's index
    auto take = [&](uint32_t want) -> uint32_t {
        uint32_t need = (uint32_t) __popc(want);
        if (chunk_end - chunk_next < need && more) { // reserve a new chunk (the rest of the old one is handed out first)
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cursor, kChunk);
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t left = chunk_end - chunk_next;
            const uint32_t rank = (uint32_t) __popc(want & below);
            uint32_t idx;
            if (rank < left) idx = chunk_next + rank;
            else idx = base + (rank - left);
            if (base >= n) more = false;
            chunk_next = base + (need - min(need, left));
            chunk_end  = min(base + kChunk, n);
            if (chunk_next > chunk_end) chunk_next = chunk_end;
            if (!(want & (1u << lane))) return 0xffffffffu;
            return idx < n ? idx : 0xffffffffu;
        }
        const uint32_t rank = (uint32_t) __popc(want & below);
        const uint32_t idx = chunk_next + rank;
        const uint32_t avail = chunk_end - chunk_next;
        chunk_next += min(need, avail);
        if (!(want & (1u << lane))) return 0xffffffffu;
        return rank < avail ? idx : 0xffffffffu;
    };

    { // prime: every lane prefetches its first ray
        nq = take(0xffffffffu);
        if (nq != 0xffffffffu) io.load(nq, nro, nrd);
    }
    for (;;) {
        // ---- refill
        const uint32_t idle = __ballot_sync(0xffffffffu, !busy);
        if (__popc(idle) >= (ANY ? kRefillMinAny : kRefillMin)) {
            io.commit(have, q, s);
            have = false;
            const bool start = !busy && nq != 0xffffffffu;
            if (start) { q = nq; s.tag = io.tag(nro); s.begin(nro, nrd); busy = true; nq = 0xffffffffu; }
            const uint32_t want = __ballot_sync(0xffffffffu, start);
            if (want && (more || chunk_next < chunk_end)) {
                const uint32_t idx = take(want);
                if (idx != 0xffffffffu) { nq = idx; io.load(nq, nro, nrd); }
            }
            if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        }
#endif
        auto node_phase = [&]() {
#pragma unroll
            for (int rep = 0; rep < MSK_NODE_STEPS; ++rep) {
                if (rep > 0 && busy && s.ngroup.y <= 0x00ffffffu && s.tgroup.y == 0u && s.sp > 0) { // next node group, if that is what is on top
                    const uint2 g = stack.load(s.sp - 1);
                    if (g.y > 0x00ffffffu) s.ngroup = g; else s.tgroup = g;
                    --s.sp;
                }
                if (busy && s.ngroup.y > 0x00ffffffu) {
                    if (s.tgroup.y) stack.store(s.sp++, s.tgroup); // postponed triangles wait on the stack
                    node_step<STATS, !(ANY && MSK_ANY_UNORDERED)>(ac, s, stack);
                }
            }
        };
        auto tri_phase = [&]() {
            const bool has_t = busy && s.tgroup.y != 0u;
            const uint32_t mt = __ballot_sync(0xffffffffu, has_t);
            if (mt) {
                const bool starved = has_t && s.ngroup.y <= 0x00ffffffu;
                if (__popc(mt) >= kTriThreshold || __any_sync(0xffffffffu, starved && waited >= MSK_TRI_PATIENCE)) {
                    if (has_t && tri_step<STATS>(ac, s) && ANY) { busy = false; have = true; }
                    // A leaf hit usually leaves two or three triangles pending, and a lane with nothing but triangles
                    // left sits out the node phases in between: up to MSK_TRI_REPS triangles per lane and phase.
#pragma unroll 1
                    for (int rep = 1; rep < MSK_TRI_REPS; ++rep) {
                        const bool again = busy && s.tgroup.y != 0u;
                        if (!__any_sync(0xffffffffu, again)) break;
                        if (again && tri_step<STATS>(ac, s) && ANY) { busy = false; have = true; }
                    }
                    waited = 0;
                } else if (starved) ++waited;
            }
        };
        auto pop_phase = [&]() {
            if (busy && s.ngroup.y <= 0x00ffffffu) {
                if (s.sp > 0) {
                    const uint2 g = stack.load(s.sp - 1);
                    if (g.y > 0x00ffffffu) { s.ngroup = g; --s.sp; }
                    else if (s.tgroup.y == 0u) { s.tgroup = g; --s.sp; }
                } else if (s.tgroup.y == 0u) { busy = false; have = true; }
            }
        };
        // (Tried: triangle -> pop -> node, i.e. the refill vote and the loop head between a node step and the test of the
        // triangles it found, as lead time for an L1 prefetch of the next triangle -- the wait for the triangle record is the
        // largest single stall of the kernel, 8 % of the samples.  Per lane the sequence of steps is the same, but a lane that
        // finishes in the pop phase then sits out a node phase before it is refilled: C2 1558 -> 1500 Mpaths/s, with the
        // prefetch 1465, C3 948 -> 913 / 900 (profiles/r03e_ab_shadow_tag_rotated_loop.txt).)
        node_phase();
        tri_phase();
        pop_phase();
    }
}

} // namespace msk
