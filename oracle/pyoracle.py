"""ctypes bindings of oracle/_build/liboracle.so (see oracle/oracle.h).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from misaki_render_b200 import capi

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"
REF_LIB = HERE / "_ref" / "librgb2spec_ref.so"
REF_COEFF = HERE / "_ref" / "srgb.coeff"


class OrcStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("seconds", C.c_double),
                ("threads", C.c_int32), ("pad_", C.c_int32)]


def build():
    subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)


_lib = None
_flavour = "checker"


def select_fast() -> bool:
    """bench.py's CPU arm only: use the speed build of the same restatement (oracle/Makefile `fast`: -O3 -march=native,
    FMA on), compiled here and now because -march=native binds it to this host.  Must be called before the first use
    of the library in the process; returns False (and keeps the checker build) when the build fails."""
    global _flavour
    assert _lib is None, "select_fast() must precede the first use of the oracle"
    fast = HERE / "_build" / "liboracle_fast.so"
    try:
        if fast.exists():
            fast.unlink()  # a copy built on another host may use instructions this one lacks
        subprocess.run(["make", "-C", str(HERE), "fast"], check=True, capture_output=True, timeout=300)
    except (subprocess.SubprocessError, OSError):
        return False
    _flavour = "fast"
    return True


def flavour() -> str:
    return _flavour


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = HERE / "_build" / "liboracle_fast.so" if _flavour == "fast" else LIB
        if not path.exists():
            build()
        L = C.CDLL(str(path))
        L.orc_last_error.restype = C.c_char_p
        L.orc_scene_create.argtypes = [C.POINTER(capi.MskSceneDesc), C.POINTER(C.c_void_p)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_destroy.restype = None
        L.orc_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_intersect_margin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_camera_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(capi.MskRenderDesc), C.c_void_p, C.c_int, C.POINTER(OrcStats)]
        L.orc_aov_channels.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_render_aov.argtypes = [C.c_void_p, C.POINTER(capi.MskRenderDesc), C.c_void_p, C.c_uint32, C.c_void_p, C.c_int,
                                     C.POINTER(OrcStats)]
        L.orc_trace_samples.argtypes = [C.c_void_p, C.POINTER(capi.MskRenderDesc), C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_develop.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_develop.restype = None
        L.orc_gaussian_filter.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_void_p]
        L.orc_gaussian_filter.restype = None
        L.orc_pcg32_floats.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]
        L.orc_pcg32_floats.restype = None
        L.orc_pcg32_uints.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]
        L.orc_pcg32_uints.restype = None
        for name in ("orc_sample_wavelength", "orc_warp", "orc_fresnel", "orc_fresnel_conductor", "orc_srgb_model_eval",
                     "orc_spectrum_to_xyz", "orc_ggx"):
            getattr(L, name).restype = None
        L.orc_sample_wavelength.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        L.orc_warp.argtypes = [C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.orc_fresnel.argtypes = [C.c_float, C.c_float, C.c_void_p]
        L.orc_fresnel_conductor.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_srgb_model_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_spectrum_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_texture_eval.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_spectrum_to_xyz.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ggx.argtypes = [C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bsdf.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.orc_rgb2spec_load.argtypes = [C.c_char_p]
        L.orc_rgb2spec_load.restype = C.c_void_p
        L.orc_rgb2spec_free.argtypes = [C.c_void_p]
        L.orc_rgb2spec_free.restype = None
        L.orc_rgb2spec_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_rgb2spec_fetch.restype = None
        _lib = L
    return _lib


def _f(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert n is None or a.size == n
    return a


class OracleScene:
    def __init__(self, desc):
        self.L = lib()
        self.desc = desc
        self.h = C.c_void_p()
        if self.L.orc_scene_create(C.byref(desc.c_desc()), C.byref(self.h)) != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        self.width, self.height = desc.width, desc.height

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def intersect(self, rays, brute_force=False):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=capi.HIT_DTYPE)
        assert self.L.orc_intersect(self.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0], int(brute_force)) == 0
        return hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        assert self.L.orc_occluded(self.h, rays.ctypes.data, occ.ctypes.data, rays.shape[0]) == 0
        return occ

    def margin(self, rays):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        t2 = np.empty(rays.shape[0], dtype=np.float32)
        mb = np.empty(rays.shape[0], dtype=np.float32)
        assert self.L.orc_intersect_margin(self.h, rays.ctypes.data, t2.ctypes.data, mb.ctypes.data, rays.shape[0]) == 0
        return t2, mb

    def camera_rays(self, samples):
        s = _f(samples).reshape(-1, 3)
        rays = np.empty(s.shape[0], dtype=capi.RAY_DTYPE)
        assert self.L.orc_camera_rays(self.h, s.ctypes.data, rays.ctypes.data, s.shape[0]) == 0
        return rays

    def render(self, rd, nthreads=0, film=None):
        if film is None:
            film = np.zeros((self.height, self.width, 5), dtype=np.float32)
        st = OrcStats()
        if self.L.orc_render(self.h, C.byref(rd), film.ctypes.data, nthreads, C.byref(st)) != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return film, st

    def render_aov(self, rd, types, nthreads=0):
        """AOVIntegrator restatement: film H x W x (5 + channels)."""
        ids = np.asarray([capi.AOV_NAMES[t] if isinstance(t, str) else int(t) for t in types], dtype=np.int32)
        nch = self.L.orc_aov_channels(ids.ctypes.data, len(ids))
        if nch < 0:
            raise RuntimeError("invalid AOV type")
        film = np.zeros((self.height, self.width, 5 + nch), dtype=np.float32)
        st = OrcStats()
        if self.L.orc_render_aov(self.h, C.byref(rd), ids.ctypes.data, len(ids), film.ctypes.data, nthreads, C.byref(st)) != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return film, st

    def trace_samples(self, rd, pixel_sample):
        ps = np.ascontiguousarray(pixel_sample, dtype=np.uint32).reshape(-1, 2)
        out = np.empty((ps.shape[0], 9), dtype=np.float32)
        assert self.L.orc_trace_samples(self.h, C.byref(rd), ps.ctypes.data, out.ctypes.data, ps.shape[0]) == 0
        return out

    def sample_ray(self, rd, seed, o, d, mint, maxt, wl, bsdf_draws_right_to_left=False):
        """MonteCarloIntegrator::sample for one camera ray, sampler seeded with `seed` (orc_sample_ray).
        bsdf_draws_right_to_left: draw next2d() before next1d() at path.cpp:71-72, as GCC orders the two arguments."""
        o, d, wl = _f(o, 3), _f(d, 3), _f(wl, 4)
        out = np.empty(4, np.float32)
        assert self.L.orc_sample_ray(self.h, C.byref(rd), C.c_uint64(seed), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                     C.c_float(mint), C.c_float(maxt), wl.ctypes.data_as(C.c_void_p), int(bsdf_draws_right_to_left),
                                     out.ctypes.data_as(C.c_void_p)) == 0
        return out

    def texture_eval(self, sid, u, v, wl):
        wl = np.ascontiguousarray(wl, dtype=np.float32)
        out = np.empty(4, np.float32)
        assert self.L.orc_texture_eval(self.h, sid, float(u), float(v), wl.ctypes.data, out.ctypes.data) == 0
        return out

    def spectrum_eval(self, sid, wl):
        wl = _f(wl, 4)
        out = np.empty(4, dtype=np.float32)
        assert self.L.orc_spectrum_eval(self.h, sid, wl.ctypes.data, out.ctypes.data) == 0
        return out

    def bsdf(self, bsdf_id, wi, wl, smp, wo):
        wi, wl, smp, wo = _f(wi, 3), _f(wl, 4), _f(smp, 3), _f(wo, 3)
        s = np.empty(10, dtype=np.float32)
        e = np.empty(4, dtype=np.float32)
        p = C.c_float()
        assert self.L.orc_bsdf(self.h, bsdf_id, wi.ctypes.data, wl.ctypes.data, smp.ctypes.data, wo.ctypes.data, s.ctypes.data,
                               e.ctypes.data, C.addressof(p)) == 0
        return dict(wo=s[:3].copy(), pdf=float(s[3]), eta=float(s[4]), type=int(s[5]), weight=s[6:10].copy(), eval=e, eval_pdf=p.value)


def develop(film):
    film = np.ascontiguousarray(film, dtype=np.float32)
    rgba = np.empty(film.shape[:2] + (4,), dtype=np.float32)
    lib().orc_develop(film.ctypes.data, rgba.ctypes.data, film.shape[0] * film.shape[1])
    return rgba


def gaussian_filter(stddev=0.5):
    r = C.c_float()
    t = np.empty(33, dtype=np.float32)
    lib().orc_gaussian_filter(stddev, C.byref(r), t.ctypes.data)
    return r.value, t


def pcg32_floats(seed, n, base_seed=0):
    out = np.empty(n, dtype=np.float32)
    lib().orc_pcg32_floats(seed, base_seed, out.ctypes.data, n)
    return out


def pcg32_uints(initstate, initseq, n):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_pcg32_uints(initstate, initseq, out.ctypes.data, n)
    return out


def sample_wavelength(u):
    wl, w = np.empty(4, np.float32), np.empty(4, np.float32)
    lib().orc_sample_wavelength(u, wl.ctypes.data, w.ctypes.data)
    return wl, w


def warp(which, u, v):
    out = np.empty(3, np.float32)
    lib().orc_warp(which, u, v, out.ctypes.data)
    return out


def fresnel(cos_theta_i, eta):
    out = np.empty(4, np.float32)
    lib().orc_fresnel(cos_theta_i, eta, out.ctypes.data)
    return out


def fresnel_conductor(cos_theta_i, eta, k):
    eta, k = _f(eta, 4), _f(k, 4)
    out = np.empty(4, np.float32)
    lib().orc_fresnel_conductor(cos_theta_i, eta.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def srgb_model_eval(c, wl):
    c, wl = _f(c, 3), _f(wl, 4)
    out = np.empty(4, np.float32)
    lib().orc_srgb_model_eval(c.ctypes.data, wl.ctypes.data, out.ctypes.data)
    return out


def spectrum_to_xyz(value, wl):
    value, wl = _f(value, 4), _f(wl, 4)
    out = np.empty(3, np.float32)
    lib().orc_spectrum_to_xyz(value.ctypes.data, wl.ctypes.data, out.ctypes.data)
    return out


def ggx(which, au, av, a, b=(0, 0, 0)):
    a, b = _f(a, 3), _f(b, 3)
    out = np.zeros(4, np.float32)
    lib().orc_ggx(which, au, av, a.ctypes.data, b.ctypes.data, out.ctypes.data)
    return out


class Rgb2Spec:
    """The oracle's restatement of rgb2spec_load/fetch."""

    def __init__(self, path=REF_COEFF):
        self.h = lib().orc_rgb2spec_load(str(path).encode())
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())

    def fetch(self, rgb):
        rgb = _f(rgb, 3)
        out = np.empty(3, np.float32)
        lib().orc_rgb2spec_fetch(self.h, rgb.ctypes.data, out.ctypes.data)
        return out


class RefRgb2Spec:
    """The UNMODIFIED reference ext/rgb2spec/rgb2spec.c compiled into oracle/_ref (oracle/Makefile.ref)."""

    def __init__(self, path=REF_COEFF):
        self.L = C.CDLL(str(REF_LIB))
        self.L.rgb2spec_load.restype = C.c_void_p
        self.L.rgb2spec_load.argtypes = [C.c_char_p]
        self.L.rgb2spec_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.rgb2spec_eval_precise.restype = C.c_float
        self.L.rgb2spec_eval_precise.argtypes = [C.c_void_p, C.c_float]
        self.h = self.L.rgb2spec_load(str(path).encode())
        assert self.h

    def fetch(self, rgb):
        rgb = _f(rgb, 3).copy()
        out = np.empty(3, np.float32)
        self.L.rgb2spec_fetch(self.h, rgb.ctypes.data, out.ctypes.data)
        return out

    def eval_precise(self, coeff, lam):
        coeff = _f(coeff, 3).copy()
        return self.L.rgb2spec_eval_precise(coeff.ctypes.data, lam)


REF_CODE_LIB = HERE / "_ref" / "libmisaki_ref_math.so"


class ReferenceCamera:
    """The reference's OWN compiled camera (oracle/ref_camera_wrap.cpp: include/misaki/core/transform.h, sensor.cpp,
    sensors/perspective.cpp over the Eigen stand-in).  Used by tools/gen_golden_ref_camera.py and, as the ray callback of
    ReferenceLoop, by tools/gen_golden_ref_converged.py."""

    def __init__(self, width, height, fov, near_clip, far_clip, to_world=None):
        self.L = C.CDLL(str(REF_CODE_LIB))
        self.L.ref_camera_create.restype = C.c_void_p
        self.L.ref_camera_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
        self.L.ref_camera_destroy.argtypes = [C.c_void_p]
        self.L.ref_camera_sample_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        self.L.ref_camera_callback.restype = C.c_void_p
        self.L.ref_camera_callback.argtypes = [C.c_void_p]
        tw = None if to_world is None else _f(np.asarray(to_world, np.float32).reshape(-1), 16).copy()
        self.h = self.L.ref_camera_create(int(width), int(height), float(fov), float(near_clip), float(far_clip), None if tw is None else tw.ctypes.data)
        if not self.h:
            raise RuntimeError("ref_camera_create failed")

    @classmethod
    def of(cls, sd):
        return cls(sd.width, sd.height, sd.fov, sd.near_clip, sd.far_clip, sd.to_world)

    def sample_rays(self, samples):
        """samples: n x 3 (wavelength sample, px, py).  Returns n x 16: o[3] d[3] mint maxt | wavelengths[4] | weight[4]."""
        s = _f(samples).reshape(-1, 3)
        out = np.empty((s.shape[0], 16), np.float32)
        if self.L.ref_camera_sample_rays(self.h, s.ctypes.data, s.shape[0], out.ctypes.data) != 0:
            raise RuntimeError("ref_camera_sample_rays failed")
        return out

    def callback(self):
        return C.c_void_p(self.L.ref_camera_callback(self.h))

    def matrices(self, width, height, fov, near_clip, far_clip):
        self.L.ref_camera_matrices.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        a, b = np.empty(16, np.float32), np.empty(16, np.float32)
        assert self.L.ref_camera_matrices(int(width), int(height), float(fov), float(near_clip), float(far_clip), a.ctypes.data, b.ctypes.data) == 0
        return a.reshape(4, 4), b.reshape(4, 4)

    def close(self):
        if self.h:
            self.L.ref_camera_destroy(self.h)
            self.h = None


def ref_lookat(origin, target, up):
    L = C.CDLL(str(REF_CODE_LIB))
    L.ref_transform_lookat.argtypes = [C.c_void_p] * 4
    o, t, u, out = _f(origin, 3).copy(), _f(target, 3).copy(), _f(up, 3).copy(), np.empty(16, np.float32)
    L.ref_transform_lookat(o.ctypes.data, t.ctypes.data, u.ctypes.data, out.ctypes.data)
    return out.reshape(4, 4)


def ref_transform(kind, v, angle=0.0):
    """kind: 'translate' | 'scale' | 'rotate' (axis v, angle as Transform4f::rotate receives it).  Returns (matrix, carried inverse)."""
    L = C.CDLL(str(REF_CODE_LIB))
    L.ref_transform_make.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    vv, a, b = _f(v, 3).copy(), np.empty(16, np.float32), np.empty(16, np.float32)
    L.ref_transform_make({"translate": 0, "scale": 1, "rotate": 2}[kind], vv.ctypes.data, float(angle), a.ctypes.data, b.ctypes.data)
    return a.reshape(4, 4), b.reshape(4, 4)


class ReferenceLoop:
    """The reference's OWN compiled render loop (oracle/ref_render_wrap.cpp: integrator.cpp, integrators/path.cpp,
    scene.cpp's emitter sampling, mesh / interaction, imageblock.cpp, films/hdrfilm.cpp, its srgb / srgb_d65 textures)
    on a scene of diffuse meshes with <rgb> reflectances and area lights -- what assets/cbox/scene.xml describes.
    Stand-ins: Embree's two calls are a brute-force Moeller-Trumbore loop, tbb::parallel_for is `threads` std::threads
    over the same tasks, Eigen is oracle/ref_shim, the camera ray of a sample comes from the oracle's camera.  Used by
    tools/gen_golden_ref_math.py and by bench.py --impl reference --ref-kind reference."""

    def __init__(self, sd, reflectance_rgb, radiance_rgb, camera="oracle"):
        """camera: "oracle" -- the oracle's restatement supplies the ray of each sample; "reference" -- the reference's own
        PerspectiveCamera does (ReferenceCamera)."""
        import os
        os.environ.setdefault("MSK_REF_DATA_ROOT", str(HERE.parent / "misaki_render_b200"))  # data/srgb.coeff (srgb.cpp:14-18)
        self.L = C.CDLL(str(REF_CODE_LIB))
        self.sd = sd
        nm = len(sd.meshes)
        self._keep = [[np.ascontiguousarray(m["verts"], np.float32) for m in sd.meshes],
                      [np.ascontiguousarray(m["tris"], np.uint32) for m in sd.meshes]]
        vptr = (C.c_void_p * nm)(*[a.ctypes.data for a in self._keep[0]])
        tptr = (C.c_void_p * nm)(*[a.ctypes.data for a in self._keep[1]])
        nv = (C.c_uint32 * nm)(*[a.shape[0] for a in self._keep[0]])
        nt = (C.c_uint32 * nm)(*[a.shape[0] for a in self._keep[1]])
        hn = (C.c_int * nm)(*[int(m["has_normals"]) for m in sd.meshes])
        hu = (C.c_int * nm)(*[int(m["has_uvs"]) for m in sd.meshes])
        refl, rad = _f(reflectance_rgb, nm * 3).copy(), _f(radiance_rgb, nm * 3).copy()
        self.L.ref_path_scene_create_rgb.restype = C.c_void_p
        self.h = self.L.ref_path_scene_create_rgb(nm, vptr, nv, tptr, nt, hn, hu, C.c_void_p(refl.ctypes.data), C.c_void_p(rad.ctypes.data))
        if not self.h:
            raise RuntimeError("ref_path_scene_create_rgb failed")
        self.osc = OracleScene(sd)
        L = lib()
        L.orc_ref_camera_bind.argtypes = [C.c_void_p]
        L.orc_ref_camera_callback.restype = C.c_void_p
        if L.orc_ref_camera_bind(self.osc.h) != 0:
            raise RuntimeError("orc_ref_camera_bind failed")
        self.cb = C.c_void_p(L.orc_ref_camera_callback())
        self.ref_camera = None
        if camera == "reference":
            self.ref_camera = ReferenceCamera.of(sd)
            self.cb = self.ref_camera.callback()
        self.L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]

    def render(self, spp: int, threads: int = 1, stddev: float = 0.5):
        """Returns (film H x W x 5, seconds).  threads == 1 is the deterministic single-task mode of the golden vectors."""
        import os
        import sys
        import time
        film = np.empty((self.sd.height, self.sd.width, 5), np.float32)
        os.environ["MSK_REF_TBB_THREADS"] = str(int(threads))
        sys.stdout.flush()
        saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)  # the reference prints a progress bar per tile (utils.h:15-39)
        os.dup2(devnull, 1)
        try:
            t = time.perf_counter()
            rc = self.L.ref_render(self.h, self.sd.width, self.sd.height, int(spp), float(stddev), self.cb, film.ctypes.data)
            dt = time.perf_counter() - t
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
            os.environ["MSK_REF_TBB_THREADS"] = "1"
        if rc != 0:
            raise RuntimeError(f"ref_render failed ({rc})")
        return film, dt
