#!/usr/bin/env python3
"""Golden vectors from the reference's OWN math headers, compiled here (oracle/Makefile.ref ->
oracle/_ref/libmisaki_ref_math.so: include/misaki/core/{mathutils,warp,frame,spectrum,distribution}.h,
include/misaki/render/{fresnel,microfacet,srgb}.h and src/librender/spectrum.cpp against the Eigen stand-in under
oracle/ref_shim/).  Writes tests/golden/ref_math.json; tests/test_oracle_ref_math.py replays the inputs through the
oracle.  Floats are stored as their uint32 bit patterns, so the comparison is exact.

    make -C oracle -f Makefile.ref && python tools/gen_golden_ref_math.py
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libmisaki_ref_math.so"
OUT = ROOT / "tests" / "golden" / "ref_math.json"
f32 = np.float32


def bits(a):
    return [int(x) for x in np.ascontiguousarray(a, dtype=f32).reshape(-1).view(np.uint32)]


def fp(a):
    return np.ascontiguousarray(a, dtype=f32).ctypes.data_as(C.c_void_p)


def unit(rng, n, upper=False):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if upper:
        v[:, 2] = np.abs(v[:, 2])
    return v.astype(f32)


def main():
    L = C.CDLL(str(LIB))
    rng = np.random.default_rng(20241017)
    g = {"source": "reference headers compiled by oracle/Makefile.ref (Eigen stand-in: oracle/ref_shim); float32 values as uint32 bit patterns"}

    # PCG32 (mathutils.h:89-143)
    cases = []
    for state, seq in [(42, 54), (0, 0xda3e39cb94b95bdb), (123456789, 0xda3e39cb94b95bdb), (2**40 + 7, 1)]:
        u = np.empty(16, np.uint32); f = np.empty(16, f32)
        L.ref_pcg32_uints(C.c_uint64(state), C.c_uint64(seq), u.ctypes.data_as(C.c_void_p), C.c_size_t(16))
        L.ref_pcg32_floats(C.c_uint64(state), C.c_uint64(seq), f.ctypes.data_as(C.c_void_p), C.c_size_t(16))
        cases.append({"state": state, "seq": seq, "uints": [int(x) for x in u], "floats": bits(f)})
    g["pcg32"] = cases

    # warps (warp.h:11-53), including the corners / axes of the concentric map
    uv = np.concatenate([rng.random((40, 2)), [[0.5, 0.5], [0, 0], [1, 1], [0.5, 0.25], [0.25, 0.5], [0.75, 0.5], [0.5, 0.75], [1, 0]]]).astype(f32)
    out = np.empty(3, f32)
    g["warp"] = []
    for which in range(4):
        for u, v in uv:
            L.ref_warp(which, C.c_float(u), C.c_float(v), fp(out))
            g["warp"].append({"which": which, "uv": bits([u, v]), "out": bits(out)})

    # coordinate_system / Frame (mathutils.h:196-203, frame.h:16-26)
    g["frame"] = []
    normals = np.concatenate([unit(rng, 20), [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, -1, 0]]]).astype(f32)
    vs = unit(rng, len(normals))
    for n, v in zip(normals, vs):
        s, t, lo, wo = (np.empty(3, f32) for _ in range(4))
        L.ref_coordinate_system(fp(n), fp(s), fp(t))
        L.ref_frame_roundtrip(fp(n), fp(v), fp(lo), fp(wo))
        g["frame"].append({"n": bits(n), "v": bits(v), "s": bits(s), "t": bits(t), "local": bits(lo), "world": bits(wo)})

    # Fresnel (fresnel.h:37-88), reflect / refract (:16-34)
    g["fresnel"] = []
    o4 = np.empty(4, f32)
    for eta in (1.5, 1.0 / 1.5, 1.33, 1.0, 2.4):
        for c in list(rng.uniform(-1, 1, 8)) + [0.0, 1.0, -1.0, 1e-4]:
            L.ref_fresnel(C.c_float(c), C.c_float(eta), fp(o4))
            g["fresnel"].append({"cos": bits([c]), "eta": bits([eta]), "out": bits(o4)})
    g["fresnel_conductor"] = []
    o3 = np.empty(3, f32)
    for _ in range(24):
        c = f32(rng.uniform(0, 1)); eta = rng.uniform(0.1, 3.0, 3).astype(f32); k = rng.uniform(0.0, 6.0, 3).astype(f32)
        L.ref_fresnel_conductor(C.c_float(c), fp(eta), fp(k), fp(o3))
        g["fresnel_conductor"].append({"cos": bits([c]), "eta": bits(eta), "k": bits(k), "out": bits(o3)})
    g["reflect_refract"] = []
    for wi, m in zip(unit(rng, 16, True), unit(rng, 16, True)):
        ct, ti = f32(rng.uniform(-1, 1)), f32(rng.uniform(0.5, 2.0))
        r, t = np.empty(3, f32), np.empty(3, f32)
        L.ref_reflect_refract(fp(wi), fp(m), C.c_float(ct), C.c_float(ti), fp(r), fp(t))
        g["reflect_refract"].append({"wi": bits(wi), "m": bits(m), "ct": bits([ct]), "ti": bits([ti]), "reflect": bits(r), "refract": bits(t)})

    # GGX (microfacet.h:11-43,108-175)
    g["ggx"] = []
    zero = np.zeros(3, f32)
    for au, av in [(0.1, 0.1), (0.3, 0.05), (0.5, 0.5), (1e-5, 1e-5)]:
        for a, b, c in zip(unit(rng, 10, True), unit(rng, 10, True), unit(rng, 10, True)):
            smp = rng.random(2).astype(f32)
            for which, (x, y, z) in enumerate([(a, zero, zero), (a, b, zero), (a, np.array([smp[0], smp[1], 0], f32), zero), (a, b, c), (a, b, zero)]):
                L.ref_ggx(which, C.c_float(au), C.c_float(av), fp(x), fp(y), fp(z), fp(o4))
                g["ggx"].append({"which": which, "au": bits([au]), "av": bits([av]), "a": bits(x), "b": bits(y), "c": bits(z), "out": bits(o4)})

    # spectral sampling and colour (spectrum.h:83-181, srgb.h:8-19; CIE table from src/librender/spectrum.cpp)
    g["sample_wavelength"] = []
    for u in list(rng.random(24)) + [0.0, 0.25, 0.5, 0.999999]:
        wl, w = np.empty(4, f32), np.empty(4, f32)
        L.ref_sample_wavelength(C.c_float(u), fp(wl), fp(w))
        g["sample_wavelength"].append({"u": bits([u]), "wl": bits(wl), "weight": bits(w)})
    g["spectrum_to_xyz"] = []
    for _ in range(24):
        wl = rng.uniform(360, 830, 4).astype(f32); val = rng.uniform(0, 2, 4).astype(f32)
        L.ref_spectrum_to_xyz(fp(val), fp(wl), fp(o3))
        rgb = np.empty(3, f32)
        L.ref_xyz_to_srgb(fp(o3.copy()), fp(rgb))
        g["spectrum_to_xyz"].append({"value": bits(val), "wl": bits(wl), "xyz": bits(o3), "rgb": bits(rgb)})
    g["srgb_model_eval"] = []
    for _ in range(24):
        c = np.array([rng.uniform(-1e-4, 1e-4), rng.uniform(-0.1, 0.1), rng.uniform(-30, 30)], f32)
        wl = rng.uniform(360, 830, 4).astype(f32)
        L.ref_srgb_model_eval(fp(c), fp(wl), fp(o4))
        g["srgb_model_eval"].append({"c": bits(c), "wl": bits(wl), "out": bits(o4)})
    for z in (np.inf, -np.inf):
        c = np.array([0, 0, z], f32); wl = np.array([400, 500, 600, 700], f32)
        L.ref_srgb_model_eval(fp(c), fp(wl), fp(o4))
        g["srgb_model_eval"].append({"c": bits(c), "wl": bits(wl), "out": bits(o4)})

    # Distribution1D (distribution.h:84-123) as Mesh::sample_position uses it
    g["distribution"] = []
    for n in (1, 2, 7, 64):
        pdf = rng.uniform(0.01, 3.0, n).astype(f32)
        u = np.concatenate([rng.random(12), [0.0, 0.5, 0.99999994]]).astype(f32)
        idx = np.empty(len(u), np.uint32); re = np.empty(len(u), f32); cdf = np.empty(n + 1, f32)
        L.ref_distribution_sample_reuse(fp(pdf), C.c_size_t(n), fp(u), C.c_size_t(len(u)), idx.ctypes.data_as(C.c_void_p), fp(re), fp(cdf))
        g["distribution"].append({"pdf": bits(pdf), "u": bits(u), "index": [int(i) for i in idx], "reused": bits(re), "cdf": bits(cdf)})

    # SmoothDiffuse::sample / eval / pdf from the reference's own src/librender/bsdfs/diffuse.cpp (constant reflectance)
    g["bsdf_diffuse"] = []
    osamp, oev, opdf = np.empty(10, f32), np.empty(4, f32), C.c_float()
    wl = np.array([420, 510, 600, 690], f32)
    for refl in (0.6, 0.05, 1.0):
        params = np.zeros(10, f32); params[0] = refl
        for wi, wo in zip(unit(rng, 12), unit(rng, 12)):  # both hemispheres: the one-sided BRDF returns zero below
            smp = rng.random(3).astype(f32)
            assert L.ref_bsdf(0, fp(params), fp(wi), fp(wl), fp(smp), fp(wo), fp(osamp), fp(oev), C.byref(opdf)) == 0
            g["bsdf_diffuse"].append({"reflectance": bits([refl]), "wi": bits(wi), "wl": bits(wl), "smp": bits(smp), "wo": bits(wo),
                                      "sample": bits(osamp), "eval": bits(oev), "pdf": bits([opdf.value])})

    # plugin sources the reference's build compiles (oracle/ref_plugins_wrap.cpp)
    g["gaussian_filter"] = []
    o35 = np.empty(35, f32)
    for stddev in (0.5, 0.25, 1.0, 0.8):
        L.ref_gaussian_filter(C.c_float(stddev), fp(o35))
        g["gaussian_filter"].append({"stddev": bits([stddev]), "radius": bits(o35[:1]), "border": int(o35[1]), "table": bits(o35[2:])})
    g["independent_sampler"] = []
    for base, seed in [(0, 0), (0, 12345), (9, 7), (1000, 2**33 + 5)]:
        o = np.empty(6 + 2 * 5, f32)
        L.ref_independent_sampler(C.c_uint64(base), C.c_uint64(seed), 6, 5, fp(o))
        g["independent_sampler"].append({"base_seed": base, "seed": seed, "n1": 6, "n2": 5, "out": bits(o)})
    g["regular_spectrum"] = []
    mean = C.c_float()
    for size, (lo, hi) in [(2, (400, 700)), (4, (400, 700)), (95, (360, 830)), (16, (300, 900))]:
        vals = rng.uniform(0.0, 2.0, size).astype(f32)
        for _ in range(6):
            wl = rng.uniform(lo - 30, hi + 30, 4).astype(f32)  # also outside the range: the index is clamped, the weights are not
            L.ref_regular_spectrum(C.c_float(lo), C.c_float(hi), fp(vals), C.c_size_t(size), fp(wl), fp(o4), C.byref(mean))
            g["regular_spectrum"].append({"range": bits([lo, hi]), "values": bits(vals), "wl": bits(wl), "out": bits(o4)})
    g["uniform_spectrum"] = []
    for wl in ([400, 500, 600, 700], [360, 500, 600, 830], [359.9, 500, 600, 700], [400, 500, 600, 830.1]):
        L.ref_uniform_spectrum(C.c_float(0.37), fp(np.array(wl, f32)), fp(o4))
        g["uniform_spectrum"].append({"value": bits([0.37]), "wl": bits(wl), "out": bits(o4)})

    # src/librender/{mesh,shape,records,interaction}.cpp (oracle/ref_mesh_wrap.cpp): hit reconstruction and mesh sampling
    import sys
    sys.path.insert(0, str(ROOT))
    from workloads import meshes as wm
    g["mesh"] = []
    cases = []
    for normals, uvs in [(False, False), (True, False), (True, True), (False, True)]:
        v, t = wm.cube_sphere(3, seed=11, octaves=2, amplitude=0.1, radius=1.0, center=(0.2, -0.1, 0.3), normals=True, uvs=True)
        v = np.ascontiguousarray(v, f32).copy()
        if not normals: v[:, 3:6] = 0
        if not uvs: v[:, 6:8] = 0
        cases.append((v, np.ascontiguousarray(t, np.uint32), normals, uvs))
    qv = np.zeros((4, 8), f32); qv[:, :3] = [(-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1)]
    cases.append((qv, np.array([[0, 1, 2], [0, 2, 3]], np.uint32), False, False))  # the C2 light quad: no normals, no uvs
    for v, t, normals, uvs in cases:
        entry = {"verts": bits(v), "tris": [int(x) for x in t.reshape(-1)], "normals": int(normals), "uvs": int(uvs), "hits": [], "samples": []}
        nv, nt = v.shape[0], t.shape[0]
        tp = t.ctypes.data_as(C.c_void_p)
        for _ in range(10):
            prim = int(rng.integers(nt)); b = rng.random(2); b = b if b.sum() < 1 else 1 - b
            u_, v_ = f32(b[0]), f32(b[1]); tt = f32(rng.uniform(0.5, 5.0))
            o = rng.uniform(-3, 3, 3).astype(f32); d = unit(rng, 1)[0]
            out = np.empty(36, f32)
            assert L.ref_mesh_interaction(fp(v), nv, tp, nt, int(normals), int(uvs), prim, C.c_float(u_), C.c_float(v_), C.c_float(tt), fp(o), fp(d), fp(out)) == 0
            entry["hits"].append({"prim": prim, "uvt": bits([u_, v_, tt]), "o": bits(o), "d": bits(d), "out": bits(out[:27])})
        cdf = np.empty(nt + 1, f32)
        for _ in range(10):
            smp = rng.random(2).astype(f32); ref_p = rng.uniform(-3, 3, 3).astype(f32)
            out = np.empty(22, f32)
            assert L.ref_mesh_sampling(fp(v), nv, tp, nt, int(normals), int(uvs), fp(smp), fp(ref_p), fp(out), fp(cdf)) == 0
            entry["samples"].append({"sample": bits(smp), "ref_p": bits(ref_p), "out": bits(out)})
        entry["cdf"] = bits(cdf)
        g["mesh"].append(entry)

    # PathTracer::sample itself (oracle/ref_path_wrap.cpp) on the Cornell box with uniform spectra
    from oracle import pyoracle
    from workloads import scenes
    L.ref_path_scene_create.restype = C.c_void_p

    def ref_scene(sd, params, env):
        nm = len(sd.meshes)
        vv = [np.ascontiguousarray(m["verts"], f32) for m in sd.meshes]; tt = [np.ascontiguousarray(m["tris"], np.uint32) for m in sd.meshes]
        vptr = (C.c_void_p * nm)(*[a.ctypes.data for a in vv]); tptr = (C.c_void_p * nm)(*[a.ctypes.data for a in tt])
        nv = (C.c_uint32 * nm)(*[a.shape[0] for a in vv]); nt = (C.c_uint32 * nm)(*[a.shape[0] for a in tt])
        hn = (C.c_int * nm)(*[int(m["has_normals"]) for m in sd.meshes]); hu = (C.c_int * nm)(*[int(m["has_uvs"]) for m in sd.meshes])
        refl = (C.c_float * nm)(*[r for r, _ in params]); rad = (C.c_float * nm)(*[-1.0 if e is None else e for _, e in params])
        handle = L.ref_path_scene_create(nm, vptr, nv, tptr, nt, hn, hu, refl, rad, C.c_float(env))
        assert handle
        return handle

    def ref_scene_rgb(sd, refl_rgb, rad_rgb):
        nm = len(sd.meshes)
        vv = [np.ascontiguousarray(m["verts"], f32) for m in sd.meshes]; tt = [np.ascontiguousarray(m["tris"], np.uint32) for m in sd.meshes]
        vptr = (C.c_void_p * nm)(*[a.ctypes.data for a in vv]); tptr = (C.c_void_p * nm)(*[a.ctypes.data for a in tt])
        nv = (C.c_uint32 * nm)(*[a.shape[0] for a in vv]); nt = (C.c_uint32 * nm)(*[a.shape[0] for a in tt])
        hn = (C.c_int * nm)(*[int(m["has_normals"]) for m in sd.meshes]); hu = (C.c_int * nm)(*[int(m["has_uvs"]) for m in sd.meshes])
        refl = np.ascontiguousarray(refl_rgb, f32); rad = np.ascontiguousarray(rad_rgb, f32)
        L.ref_path_scene_create_rgb.restype = C.c_void_p
        handle = L.ref_path_scene_create_rgb(nm, vptr, nv, tptr, nt, hn, hu, fp(refl), fp(rad))
        assert handle
        return handle

    def path_vectors(sd, params, env, n, seed0, aov=False):
        nm = len(sd.meshes)
        vv = [np.ascontiguousarray(m["verts"], f32) for m in sd.meshes]; tt = [np.ascontiguousarray(m["tris"], np.uint32) for m in sd.meshes]
        vptr = (C.c_void_p * nm)(*[a.ctypes.data for a in vv]); tptr = (C.c_void_p * nm)(*[a.ctypes.data for a in tt])
        nv = (C.c_uint32 * nm)(*[a.shape[0] for a in vv]); nt = (C.c_uint32 * nm)(*[a.shape[0] for a in tt])
        hn = (C.c_int * nm)(*[int(m["has_normals"]) for m in sd.meshes]); hu = (C.c_int * nm)(*[int(m["has_uvs"]) for m in sd.meshes])
        refl = (C.c_float * nm)(*[r for r, _ in params]); rad = (C.c_float * nm)(*[-1.0 if e is None else e for _, e in params])
        handle = L.ref_path_scene_create(nm, vptr, nv, tptr, nt, hn, hu, refl, rad, C.c_float(env))
        assert handle
        smp = np.stack([rng.uniform(0, sd.width, n), rng.uniform(0, sd.height, n), rng.random(n)], axis=1).astype(f32)
        rays = pyoracle.OracleScene(sd).camera_rays(smp)  # camera rays are INPUT here (perspective.cpp is not part of this build)
        out_list = []
        for i in range(n):
            wl, _ = pyoracle.sample_wavelength(float(smp[i, 2]))
            out = np.empty(20 if aov else 4, f32)
            r = rays[i]
            fn = L.ref_aov_sample if aov else L.ref_path_sample
            assert fn(C.c_void_p(handle), C.c_uint64(seed0 + i), fp(r["o"]), fp(r["d"]), C.c_float(r["tmin"]), C.c_float(r["tmax"]), fp(wl), fp(out)) == 0
            out_list.append({"seed": seed0 + i, "o": bits(r["o"]), "d": bits(r["d"]), "t": bits([r["tmin"], r["tmax"]]), "wl": bits(wl), "out": bits(out)})
        nz = sum(1 for c in out_list if np.array(c["out"], np.uint32).view(f32).any())
        print(f"path_sample: {nz} of {n} paths returned radiance")
        return out_list

    g["path_sample"] = path_vectors(scenes.cbox_uniform(64, 64), [(r, e) for _, r, e in scenes.CBOX_UNIFORM], -1.0, 400, 1000)
    sd2, params2 = scenes.open_uniform(64, 64)
    g["path_sample_env"] = path_vectors(sd2, params2, scenes.OPEN_UNIFORM_ENV, 400, 5000)
    # AOVIntegrator::sample (integrators/aov.cpp) with the path tracer nested, same open scene
    g["aov_sample"] = path_vectors(sd2, params2, scenes.OPEN_UNIFORM_ENV, 200, 9000, aov=True)

    # src/librender/imageblock.cpp (oracle/ref_film_wrap.cpp): filtered splat, block merge, spiral order.  The film is
    # stored as a SHA-256 of its bytes plus a few probe pixels (a 70 x 45 x 5 film would be 60 KB of JSON).
    import hashlib
    g["film"] = []
    for k, (W, H, bs, stddev, n) in enumerate([(70, 45, 32, 0.5, 3000), (33, 64, 32, 0.5, 2000), (40, 40, 16, 0.8, 2000)]):
        nch = 5
        smp = scenes.film_samples(W, H, n, 777 + k)
        nb = int(np.ceil(W / bs) * np.ceil(H / bs))
        film = np.empty((H, W, nch), f32); order = np.empty((nb, 4), np.int32)
        assert L.ref_film_accumulate(C.c_float(stddev), W, H, nch, bs, fp(smp), C.c_size_t(n), fp(film), order.ctypes.data_as(C.c_void_p)) == 0
        probes = [[int(y), int(x)] for y, x in zip(rng.integers(0, H, 12), rng.integers(0, W, 12))] + [[0, 0], [H - 1, W - 1], [31, 31], [32, 32 if W > 32 else W - 1]]
        g["film"].append({"W": W, "H": H, "block_size": bs, "stddev": bits([stddev]), "sample_seed": 777 + k, "samples_sha256": hashlib.sha256(smp.tobytes()).hexdigest(), "n": n, "sha256": hashlib.sha256(film.tobytes()).hexdigest(),
                          "order": [int(v) for v in order.reshape(-1)], "probes": probes, "probe_values": bits(np.stack([film[y, x] for y, x in probes]))})

    # src/librender/shapes/obj.cpp (oracle/ref_obj_wrap.cpp): what the reference's own loader makes of OBJ files -- the
    # Cornell-box meshes of this repository and a synthetic file with quads, shared / split vertices, normals, texcoords
    import tempfile
    L.ref_obj_load.restype = C.c_void_p
    synthetic = ("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 0.25\nvn 0 0 1\nvn 0 1 0\nvn 1 2 3\n"
                 "f 1/1/1 2/2/1 3/3/1 4/4/1\nf 1/1/2 2/2/2 5/5/3\nf 2/2/1 3/3/1 5/5/3\nf 3/3 4/4 5/5\nf 4//2 1//2 5//3\n")
    g["obj"] = []
    files = [("assets/cbox/meshes/cbox_%s.obj" % n, None) for n, _, _ in scenes.CBOX_UNIFORM] + [("synthetic.obj", synthetic)]
    for rel, text in files:
        for flip in (1, 0):
            if text is None:
                path = str(ROOT / rel)
            else:
                tf = tempfile.NamedTemporaryFile("w", suffix=".obj", delete=False); tf.write(text); tf.close(); path = tf.name
            h = L.ref_obj_load(path.encode(), flip)
            assert h, rel
            counts = (C.c_uint32 * 4)()
            L.ref_obj_get(C.c_void_p(h), counts, None, None)
            v = np.zeros((counts[0], 8), f32); t = np.zeros((counts[1], 3), np.uint32)
            L.ref_obj_get(C.c_void_p(h), counts, fp(v), t.ctypes.data_as(C.c_void_p))
            if not counts[2]: v[:, 3:6] = 0  # absent attributes are uninitialised memory in the reference
            if not counts[3]: v[:, 6:8] = 0
            g["obj"].append({"file": rel, "text": text, "flip": flip, "counts": [int(c) for c in counts], "verts": bits(v), "faces": [int(x) for x in t.reshape(-1)]})

    # what an <rgb> tag becomes (oracle/ref_spectra_wrap.cpp): srgb.cpp + spectra/{srgb,srgb_d65,d65}.cpp with the table
    # data/srgb.coeff generated by the reference's own optimiser (oracle/_ref/srgb.coeff == misaki_render_b200/data/srgb.coeff)
    import os
    os.environ["MSK_REF_DATA_ROOT"] = str(ROOT / "misaki_render_b200")
    g["colour_spectrum"] = []
    for kind in (0, 1, 2):
        for _ in range(16):
            rgb = (rng.uniform(0, 1, 3) if kind == 0 else rng.uniform(0, 40, 3)).astype(f32)
            if _ == 0: rgb = np.array([0.5, 0.5, 0.5], f32)
            if _ == 1 and kind == 0: rgb = np.array([0, 0, 0], f32)  # (a black <rgb> inside an emitter throws in the reference: empty distribution)
            scale = f32(1.0 if _ % 2 == 0 else rng.uniform(0.2, 3.0))
            wl = rng.uniform(360, 830, 4).astype(f32)
            assert L.ref_colour_spectrum(kind, fp(rgb), C.c_float(scale), fp(wl), fp(o4)) == 0
            g["colour_spectrum"].append({"kind": kind, "rgb": bits(rgb), "scale": bits([scale]), "wl": bits(wl), "out": bits(o4)})

    # SamplingIntegrator::render itself (oracle/ref_render_wrap.cpp): whole films of the reference's own tile loop
    g["render"] = []
    CB = C.CFUNCTYPE(None, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float))
    for (W, H, spp, which) in [(48, 40, 2, "cbox"), (40, 33, 2, "open"), (44, 36, 3, "cbox_rgb")]:
        if which == "cbox":
            sd, params, env = scenes.cbox_uniform(W, H), [(r, e) for _, r, e in scenes.CBOX_UNIFORM], -1.0
        elif which == "cbox_rgb":  # config C1 itself: the <rgb> reflectances and the (40, 40, 40) luminaire of assets/cbox/scene.xml
            sd = scenes.cbox(W, H)
        else:
            (sd, params), env = scenes.open_uniform(W, H), scenes.OPEN_UNIFORM_ENV
        osc = pyoracle.OracleScene(sd)

        def camera(ws, px, py, out, osc=osc):  # the camera ray of a sample is an input (perspective.cpp is not in the pinned build)
            r = osc.camera_rays(np.array([[px, py, ws]], f32))[0]
            wl, w = pyoracle.sample_wavelength(float(ws))
            vals = list(r["o"]) + list(r["d"]) + [r["tmin"], r["tmax"]] + list(wl) + list(w)
            for i, v in enumerate(vals):
                out[i] = float(v)

        film = np.empty((H, W, 5), f32)
        if which == "cbox_rgb":
            handle = ref_scene_rgb(sd, [r for _, r in scenes.CBOX_SHAPES], [(40, 40, 40) if n == "luminaire" else (-1, -1, -1) for n, _ in scenes.CBOX_SHAPES])
        else:
            handle = ref_scene(sd, params, env)
        assert L.ref_render(C.c_void_p(handle), W, H, spp, C.c_float(0.5), CB(camera), fp(film)) == 0
        probes = [[int(y), int(x)] for y, x in zip(rng.integers(0, H, 10), rng.integers(0, W, 10))] + [[0, 0], [H - 1, W - 1], [31, 31], [32, 32]]
        g["render"].append({"scene": which, "W": W, "H": H, "spp": spp, "sha256": hashlib.sha256(film.tobytes()).hexdigest(), "probes": probes,
                            "probe_values": bits(np.stack([film[y, x] for y, x in probes])), "mean": bits(film.mean(axis=(0, 1)))})
        print(f"render {which} {W}x{H}x{spp}: mean XYZAW = {film.mean(axis=(0, 1))}")
        if which == "cbox":  # the same loop with AOVIntegrator (every camera ray of this scene hits: no uninitialised reads)
            afilm = np.empty((H, W, 21), f32)
            assert L.ref_render_aov(C.c_void_p(handle), W, H, spp, C.c_float(0.5), CB(camera), fp(afilm)) == 0
            assert np.isfinite(afilm).all()
            g["render_aov"] = {"scene": which, "W": W, "H": H, "spp": spp, "sha256": hashlib.sha256(afilm.tobytes()).hexdigest(), "probes": probes,
                               "probe_values": bits(np.stack([afilm[y, x] for y, x in probes]))}

    # HDRFilm::image (films/hdrfilm.cpp:48-90 over film.cpp and the reference's own ImageBlock storage): the develop step
    g["hdrfilm_image"] = []
    for (W, H, nch) in [(16, 12, 5), (9, 7, 8)]:
        film = np.empty((H, W, nch), f32)
        film[..., :3] = rng.random((H, W, 3)) * rng.choice([0.01, 1.0, 60.0, 4000.0], (H, W, 1))
        film[..., 4] = rng.random((H, W)) * 4 + 0.25
        film[..., 3] = film[..., 4] * (rng.random((H, W)) > 0.2)
        film[..., 5:] = rng.normal(size=(H, W, nch - 5)) * 100
        film[0, 0] = 0.0  # a pixel no sample reached: weight 0 -> 0, not NaN (hdrfilm.cpp:73-74)
        film[1, 2, 4] = 0.0  # radiance without weight
        out = np.empty((H, W, nch - 1), f32)
        assert L.ref_hdrfilm_image(fp(film), W, H, nch, fp(out)) == 0
        assert np.isfinite(out).all()
        g["hdrfilm_image"].append({"W": W, "H": H, "nch": nch, "film": bits(film), "image": bits(out)})

    OUT.write_text(json.dumps(g, separators=(",", ":")))
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes): " + ", ".join(f"{k}={len(v)}" for k, v in g.items() if isinstance(v, list)))


if __name__ == "__main__":
    main()
