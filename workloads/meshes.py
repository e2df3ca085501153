"""Seeded procedural meshes (numpy).  All return (verts[N,8] float32, tris[M,3] uint32) with the
reference's vertex layout [px py pz nx ny nz u v] (shapes/obj.cpp:139-142)."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def read_obj(path, to_world=None):
    """OBJ reader with the vertex indexing of the reference loader (shapes/obj.cpp:58-181):
    positions transformed by to_world, quads split as (v1 v2 v3)(v4 v1 v3), one output vertex per
    distinct (p, uv, n) key in order of first appearance.  Returns verts, tris, has_normals, has_uvs."""
    P, N, T, keys, order, tris = [], [], [], {}, [], []
    M = np.eye(4, dtype=f32) if to_world is None else np.asarray(to_world, dtype=f32)
    Minv_T = np.linalg.inv(M.astype(np.float64)).astype(f32)[:3, :3].T
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            p = np.array([t[1], t[2], t[3], 1.0], dtype=f32)
            r = (M @ p).astype(f32)
            P.append((r[:3] / r[3]).astype(f32))
        elif t[0] == "vt":
            T.append(np.array([float(t[1]), 1.0 - float(t[2])], dtype=f32))  # filp_tex_coords default true
        elif t[0] == "vn":
            n = (Minv_T @ np.array(t[1:4], dtype=f32)).astype(f32)
            N.append((n / f32(np.sqrt(np.dot(n, n)))).astype(f32))
        elif t[0] == "f":
            vs = t[1:5]
            idx = [0, 1, 2] if len(vs) == 3 else [0, 1, 2, 3, 0, 2]
            for k in idx:
                parts = (vs[k].split("/") + ["", ""])[:3]
                key = (int(parts[0]), int(parts[1]) if parts[1] else -1, int(parts[2]) if parts[2] else -1)
                if key not in keys:
                    keys[key] = len(order)
                    order.append(key)
                tris.append(keys[key])
    verts = np.zeros((len(order), 8), dtype=f32)
    for i, (p, uv, n) in enumerate(order):
        verts[i, :3] = P[p - 1]
        if n != -1:
            verts[i, 3:6] = N[n - 1]
        if uv != -1:
            verts[i, 6:8] = T[uv - 1]
    return verts, np.asarray(tris, dtype=np.uint32).reshape(-1, 3), bool(N), bool(T)


def quad(p0, p1, p2, p3):
    """Two triangles (0,1,2),(3,0,2) like the OBJ loader's quad split."""
    v = np.zeros((4, 8), dtype=f32)
    v[:, :3] = np.array([p0, p1, p2, p3], dtype=f32)
    return v, np.array([[0, 1, 2], [3, 0, 2]], dtype=np.uint32)


def _value_noise(p, seed, octaves, base_freq=2.0):
    """Smooth periodic-free value noise on R^3 (hash lattice + quintic interpolation), float64."""
    def hash3(ix, iy, iz, s):
        h = (ix.astype(np.int64) * 73856093) ^ (iy.astype(np.int64) * 19349663) ^ (iz.astype(np.int64) * 83492791) ^ (s * 2654435761)
        h = (h ^ (h >> 13)) * 1274126177
        h = h ^ (h >> 16)
        return (h & 0xFFFFFF).astype(np.float64) / float(0xFFFFFF) * 2.0 - 1.0
    out = np.zeros(p.shape[0])
    amp, freq = 1.0, base_freq
    for o in range(octaves):
        q = p * freq
        i = np.floor(q)
        f = q - i
        u = f * f * f * (f * (f * 6 - 15) + 10)
        ix, iy, iz = i[:, 0], i[:, 1], i[:, 2]
        acc = 0.0
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (u[:, 0] if dx else 1 - u[:, 0]) * (u[:, 1] if dy else 1 - u[:, 1]) * (u[:, 2] if dz else 1 - u[:, 2])
                    acc = acc + w * hash3(ix + dx, iy + dy, iz + dz, seed + 101 * o)
        out += amp * acc
        amp *= 0.5
        freq *= 2.0
    return out


def _vertex_normals(pos, tris):
    fn = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]])
    vn = np.zeros_like(pos)
    for k in range(3):
        np.add.at(vn, tris[:, k], fn)
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    return vn / np.where(ln > 0, ln, 1.0)


def cube_sphere(n, seed, octaves=3, amplitude=0.12, radius=1.0, center=(0, 0, 0), stretch=(1, 1, 1), normals=True, uvs=False):
    """Closed, welded cube-sphere with 12 n^2 triangles, displaced radially by value noise."""
    faces = []
    lin = np.linspace(-1.0, 1.0, n + 1)
    a, b = np.meshgrid(lin, lin, indexing="ij")
    one = np.ones_like(a)
    for axis, sign in [(0, 1), (0, -1), (1, 1), (1, -1), (2, 1), (2, -1)]:
        c = [None, None, None]
        c[axis] = sign * one
        c[(axis + 1) % 3] = a if sign > 0 else b
        c[(axis + 2) % 3] = b if sign > 0 else a
        faces.append(np.stack(c, axis=-1).reshape(-1, 3))
    cube = np.concatenate(faces)
    # weld shared edge/corner vertices on an integer lattice
    lattice = np.rint((cube + 1.0) * 0.5 * n).astype(np.int64)
    key = (lattice[:, 0] * (n + 1) + lattice[:, 1]) * (n + 1) + lattice[:, 2]
    uniq, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    cube_u = lattice[first].astype(np.float64) / n * 2.0 - 1.0
    d = cube_u / np.linalg.norm(cube_u, axis=1, keepdims=True)
    r = radius * (1.0 + amplitude * _value_noise(d, seed, octaves))
    pos = d * r[:, None] * np.asarray(stretch, dtype=np.float64) + np.asarray(center, dtype=np.float64)
    tris = []
    stride = (n + 1) * (n + 1)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    for f in range(6):
        v00 = f * stride + ii * (n + 1) + jj
        v10, v01, v11 = v00 + (n + 1), v00 + 1, v00 + (n + 2)
        tris.append(np.stack([v00, v10, v11], axis=-1).reshape(-1, 3))
        tris.append(np.stack([v00, v11, v01], axis=-1).reshape(-1, 3))
    tris = inverse[np.concatenate(tris)]
    # outward orientation
    fn = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]])
    fc = pos[tris].mean(axis=1) - np.asarray(center, dtype=np.float64)
    flip = np.einsum("ij,ij->i", fn, fc) < 0
    tris[flip] = tris[flip][:, ::-1]
    verts = np.zeros((pos.shape[0], 8), dtype=f32)
    verts[:, :3] = pos.astype(f32)
    if normals:
        verts[:, 3:6] = _vertex_normals(verts[:, :3].astype(np.float64), tris).astype(f32)
    if uvs:
        verts[:, 6] = (np.arctan2(d[:, 1], d[:, 0]) / (2 * np.pi) + 0.5).astype(f32)
        verts[:, 7] = (np.arccos(np.clip(d[:, 2], -1, 1)) / np.pi).astype(f32)
    return verts, np.ascontiguousarray(tris, dtype=np.uint32)


def sphere_grid(nu=3163, nv=1582, seed=3, octaves=5, amplitude=0.08):
    """The C5 mesh: an nu x nv vertex grid on a unit sphere displaced by fBm; 2 (nu-1)(nv-1) triangles
    (9 998 244 for the default size), single geomID, no normals/uvs.  float32 throughout for speed."""
    u = np.linspace(0.0, 2.0 * np.pi, nu, dtype=np.float64)
    v = np.linspace(0.02, np.pi - 0.02, nv, dtype=np.float64)  # open at the poles: no degenerate triangles
    uu, vv = np.meshgrid(u, v, indexing="ij")
    d = np.stack([np.sin(vv) * np.cos(uu), np.sin(vv) * np.sin(uu), np.cos(vv)], axis=-1).reshape(-1, 3)
    r = 1.0 + amplitude * _value_noise(d, seed, octaves, base_freq=3.0)
    verts = np.zeros((d.shape[0], 8), dtype=f32)
    verts[:, :3] = (d * r[:, None]).astype(f32)
    i, j = np.meshgrid(np.arange(nu - 1, dtype=np.int64), np.arange(nv - 1, dtype=np.int64), indexing="ij")
    v00 = (i * nv + j).reshape(-1)
    v10, v01, v11 = v00 + nv, v00 + 1, v00 + nv + 1
    tris = np.concatenate([np.stack([v00, v01, v11], axis=-1), np.stack([v00, v11, v10], axis=-1)])
    return verts, np.ascontiguousarray(tris, dtype=np.uint32)
