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

constexpr int kStackSize = 48;

struct RayHit {
    float t, u, v;
    uint32_t prim, geom;
};

struct WoopRay { // per-ray constants of the watertight test
    int kx, ky, kz;
    float sx, sy, sz;
};

__device__ __forceinline__ float pick(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ WoopRay woop_setup(float dx, float dy, float dz) {
    WoopRay w;
    float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    w.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    w.kx = w.kz + 1; if (w.kx == 3) w.kx = 0;
    w.ky = w.kx + 1; if (w.ky == 3) w.ky = 0;
    float dkz = pick(dx, dy, dz, w.kz);
    if (dkz < 0.f) { int tmp = w.kx; w.kx = w.ky; w.ky = tmp; }
    w.sx = pick(dx, dy, dz, w.kx) / dkz;
    w.sy = pick(dx, dy, dz, w.ky) / dkz;
    w.sz = 1.f / dkz;
    return w;
}

// Returns true and updates (t,u,v) when tnear < t <= tfar.
__device__ __forceinline__ bool woop_intersect(const WoopRay &w, float ox, float oy, float oz, float4 v0, float4 v1, float4 v2,
                                               float tnear, float tfar, float &t_out, float &u_out, float &v_out) {
    float ax = v0.x - ox, ay = v0.y - oy, az = v0.z - oz;
    float bx = v1.x - ox, by = v1.y - oy, bz = v1.z - oz;
    float cx = v2.x - ox, cy = v2.y - oy, cz = v2.z - oz;
    float Akx = pick(ax, ay, az, w.kx), Aky = pick(ax, ay, az, w.ky), Akz = pick(ax, ay, az, w.kz);
    float Bkx = pick(bx, by, bz, w.kx), Bky = pick(bx, by, bz, w.ky), Bkz = pick(bx, by, bz, w.kz);
    float Ckx = pick(cx, cy, cz, w.kx), Cky = pick(cx, cy, cz, w.ky), Ckz = pick(cx, cy, cz, w.kz);
    // shear: explicit round-to-nearest mul/sub (never contracted to FMA) so that the edge functions of a
    // shared edge are exact negatives of each other in the two triangles -- the watertightness property
    float Ax = __fsub_rn(Akx, __fmul_rn(w.sx, Akz)), Ay = __fsub_rn(Aky, __fmul_rn(w.sy, Akz));
    float Bx = __fsub_rn(Bkx, __fmul_rn(w.sx, Bkz)), By = __fsub_rn(Bky, __fmul_rn(w.sy, Bkz));
    float Cx = __fsub_rn(Ckx, __fmul_rn(w.sx, Ckz)), Cy = __fsub_rn(Cky, __fmul_rn(w.sy, Ckz));
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
    float det = U + V + W;
    if (det == 0.f) return false;
    float Az = w.sz * Akz, Bz = w.sz * Bkz, Cz = w.sz * Ckz;
    float T  = U * Az + V * Bz + W * Cz;
    float sgn = det < 0.f ? -1.f : 1.f;
    float Ts = T * sgn, absdet = fabsf(det);
    if (!(Ts > tnear * absdet) || !(Ts <= tfar * absdet)) return false;
    float rcp = 1.f / det;
    t_out = T * rcp;
    u_out = V * rcp;
    v_out = W * rcp;
    return true;
}

__device__ __forceinline__ float nz(float d) { // keep reciprocal directions finite
    return fabsf(d) < 1e-18f ? copysignf(1e-18f, d) : d;
}

__device__ __forceinline__ uint32_t extract_byte(uint32_t x, int i) { return (x >> (8 * i)) & 0xffu; }

// ANY: stop at the first accepted hit.  STATS: count visited nodes / tested triangles.
template <bool ANY, bool STATS>
__device__ __forceinline__ bool traverse(const float4 *__restrict__ nodes, const float4 *__restrict__ tris, float ox, float oy,
                                         float oz, float dx, float dy, float dz, float tmin, float tmax, RayHit &hit,
                                         uint32_t *nnodes = nullptr, uint32_t *ntris = nullptr) {
    uint2 stack[kStackSize];
    int sp = 0;
    const float idx = 1.f / nz(dx), idy = 1.f / nz(dy), idz = 1.f / nz(dz);
    const bool negx = idx < 0.f, negy = idy < 0.f, negz = idz < 0.f; // signs of the clamped direction (-0 counts as negative)
    const uint32_t oct    = (negx ? 4u : 0u) | (negy ? 2u : 0u) | (negz ? 1u : 0u);
    const uint32_t octinv = 7u - oct;
    const uint32_t octinv4 = octinv * 0x01010101u;
    const WoopRay wr = woop_setup(dx, dy, dz);
    bool found = false;
    uint32_t cnt_nodes = 0, cnt_tris = 0;

    uint2 ngroup = make_uint2(0u, 0x80000000u); // root: "child slot 7 of a virtual parent at base 0"
    uint2 tgroup = make_uint2(0u, 0u);
    for (;;) {
        // invariant: ngroup.y > 0x00ffffff here (only such groups are pushed, and the loop
        // pops or exits as soon as the current group has no pending internal hits)
        {
            const uint32_t hits  = ngroup.y;
            const uint32_t imask = ngroup.y & 0xffu;
            const uint32_t bit   = 31u - __clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rel  = __popc(imask & ~(0xffffffffu << slot));
            const uint32_t node = ngroup.x + rel;
            const float4 *np = nodes + (size_t) node * kNodeFloat4s;
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (STATS) cnt_nodes++;
            const uint32_t e = __float_as_uint(n0.w);
            const float adx = __uint_as_float((e & 0xffu) << 23) * idx, ady = __uint_as_float(((e >> 8) & 0xffu) << 23) * idy,
                        adz = __uint_as_float(((e >> 16) & 0xffu) << 23) * idz;
            const float aox = (n0.x - ox) * idx, aoy = (n0.y - oy) * idy, aoz = (n0.z - oz) * idz;
            ngroup.x = __float_as_uint(n1.x);
            tgroup.x = __float_as_uint(n1.y);
            uint32_t hitmask = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t meta4 = __float_as_uint(h ? n1.w : n1.z);
                const uint32_t is_inner4   = (meta4 & (meta4 << 1)) & 0x10101010u;
                const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xffu;
                const uint32_t bit_index4  = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
                const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
                const uint32_t qlox = __float_as_uint(h ? n2.y : n2.x), qloy = __float_as_uint(h ? n2.w : n2.z),
                               qloz = __float_as_uint(h ? n3.y : n3.x), qhix = __float_as_uint(h ? n3.w : n3.z),
                               qhiy = __float_as_uint(h ? n4.y : n4.x), qhiz = __float_as_uint(h ? n4.w : n4.z);
                // near / far planes by ray direction sign
                const uint32_t nx = negx ? qhix : qlox, fx = negx ? qlox : qhix;
                const uint32_t ny = negy ? qhiy : qloy, fy = negy ? qloy : qhiy;
                const uint32_t nzq = negz ? qhiz : qloz, fz = negz ? qloz : qhiz;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float t0x = fmaf((float) extract_byte(nx, j), adx, aox), t1x = fmaf((float) extract_byte(fx, j), adx, aox);
                    const float t0y = fmaf((float) extract_byte(ny, j), ady, aoy), t1y = fmaf((float) extract_byte(fy, j), ady, aoy);
                    const float t0z = fmaf((float) extract_byte(nzq, j), adz, aoz), t1z = fmaf((float) extract_byte(fz, j), adz, aoz);
                    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
                    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
                    if (tn <= tf) hitmask |= extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
                }
            }
            ngroup.y = (hitmask & 0xff000000u) | (e >> 24);
            tgroup.y = hitmask & 0x00ffffffu;
        }
        while (tgroup.y) {
            const uint32_t k = 31u - __clz(tgroup.y);
            tgroup.y &= ~(1u << k);
            const float4 *tp = tris + (size_t) (tgroup.x + k) * kTriFloat4s;
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            if (STATS) cnt_tris++;
            float t, u, v;
            if (woop_intersect(wr, ox, oy, oz, v0, v1, v2, tmin, tmax, t, u, v)) {
                tmax = t;
                hit.t = t; hit.u = u; hit.v = v;
                hit.prim = __float_as_uint(v0.w); hit.geom = __float_as_uint(v1.w);
                found = true;
                if (ANY) { tgroup.y = 0; ngroup.y = 0; sp = 0; break; }
            }
        }
        if (ngroup.y <= 0x00ffffffu) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    if (STATS) { *nnodes = cnt_nodes; *ntris = cnt_tris; }
    return found;
}

} // namespace msk
