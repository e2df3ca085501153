"""ctypes mirror of include/misaki_b200.h (the drop-in C ABI).

Only plain pointers and sizes cross this boundary.  Every wrapper raises
``MskError`` with ``msk_gpu_last_error()`` when the library reports a failure --
mirroring the reference's ``Throw(...)`` -> ``std::runtime_error``
(include/misaki/core/logger.h:81-85).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
# MSK_B200_LIB: development override used by tools/sweep_variants.sh to compare builds of the same library
LIB_PATH = Path(os.environ["MSK_B200_LIB"]) if os.environ.get("MSK_B200_LIB") else ROOT / "lib" / "libmisaki_b200.so"

EXPORTED_SYMBOLS = [
    "msk_gpu_abi_version", "msk_gpu_last_error", "msk_gpu_init", "msk_gpu_shutdown", "msk_gpu_stream",
    "msk_gpu_scene_create", "msk_gpu_scene_destroy", "msk_gpu_accel_info", "msk_gpu_intersect",
    "msk_gpu_occluded", "msk_gpu_intersect_dev", "msk_gpu_occluded_dev", "msk_gpu_intersect_stats",
    "msk_gpu_render", "msk_gpu_render_dev", "msk_gpu_render_reserve", "msk_gpu_develop", "msk_gpu_develop_dev",
    "msk_gpu_aov_channels", "msk_gpu_render_aov", "msk_gpu_render_aov_dev",
    "msk_gpu_film_share_create", "msk_gpu_film_share_ptr", "msk_gpu_film_share_export", "msk_gpu_film_share_open",
    "msk_gpu_reduce_film", "msk_gpu_film_share_check", "msk_gpu_film_share_destroy",
    "msk_gpu_device_count", "msk_gpu_film_share_attach", "msk_gpu_render_multi",
]

# enums
SPEC_UNIFORM, SPEC_SRGB, SPEC_SRGB_D65, SPEC_REGULAR, SPEC_SRGB_UNBOUNDED, SPEC_CHECKERBOARD = range(6)
BSDF_DIFFUSE, BSDF_CONDUCTOR, BSDF_ROUGHCONDUCTOR, BSDF_ROUGHDIELECTRIC, BSDF_DIELECTRIC = range(5)
EMITTER_AREA, EMITTER_CONSTANT = range(2)
RENDER_STAGE_TIMERS = 1
RENDER_TRAVERSAL_STATS = 2
ABI_VERSION = 5
INTEGRATOR_PATH, INTEGRATOR_VOLPATH = range(2)
AOV_DEPTH, AOV_POSITION, AOV_UV, AOV_GEO_NORMAL, AOV_SH_NORMAL, AOV_INTEGRATOR_RGBA = range(6)
AOV_NAMES = {"depth": AOV_DEPTH, "position": AOV_POSITION, "uv": AOV_UV, "geo_normal": AOV_GEO_NORMAL, "sh_normal": AOV_SH_NORMAL,
             "integrator": AOV_INTEGRATOR_RGBA}


class MskError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[msk {code}] {message}")
        self.code = code


class MskSpectrum(C.Structure):
    _fields_ = [("kind", C.c_int32), ("c", C.c_float * 3), ("value", C.c_float), ("table_offset", C.c_uint32),
                ("table_size", C.c_uint32), ("lambda_min", C.c_float), ("lambda_max", C.c_float),
                ("child0", C.c_int32), ("child1", C.c_int32), ("to_uv", C.c_float * 6)]


class MskBsdf(C.Structure):
    _fields_ = [("type", C.c_int32), ("reflectance", C.c_int32), ("transmittance", C.c_int32), ("eta", C.c_int32),
                ("k", C.c_int32), ("alpha_u", C.c_float), ("alpha_v", C.c_float), ("int_ior", C.c_float),
                ("ext_ior", C.c_float), ("distribution", C.c_int32), ("sample_visible", C.c_int32), ("twosided", C.c_int32)]


class MskEmitter(C.Structure):
    _fields_ = [("type", C.c_int32), ("radiance", C.c_int32), ("shape", C.c_int32)]


class MskMesh(C.Structure):
    _fields_ = [("verts", C.POINTER(C.c_float)), ("tris", C.POINTER(C.c_uint32)), ("nverts", C.c_uint32), ("ntris", C.c_uint32),
                ("bsdf", C.c_int32), ("emitter", C.c_int32), ("has_normals", C.c_uint8), ("has_uvs", C.c_uint8),
                ("pad_", C.c_uint8 * 2), ("interior_medium", C.c_int32), ("exterior_medium", C.c_int32)]


class MskMedium(C.Structure):
    _fields_ = [("sigma_a", C.c_int32), ("sigma_s", C.c_int32), ("phase", C.c_int32), ("scale", C.c_float)]


class MskCamera(C.Structure):
    _fields_ = [("sample_to_camera", C.c_float * 16), ("to_world", C.c_float * 16), ("near_clip", C.c_float),
                ("far_clip", C.c_float), ("width", C.c_uint32), ("height", C.c_uint32), ("filter_radius", C.c_float),
                ("filter_table", C.c_float * 33)]


class MskSceneDesc(C.Structure):
    _fields_ = [("meshes", C.POINTER(MskMesh)), ("nmeshes", C.c_uint32), ("bsdfs", C.POINTER(MskBsdf)), ("nbsdfs", C.c_uint32),
                ("emitters", C.POINTER(MskEmitter)), ("nemitters", C.c_uint32), ("spectra", C.POINTER(MskSpectrum)),
                ("nspectra", C.c_uint32), ("spectrum_tables", C.POINTER(C.c_float)), ("ntable_floats", C.c_uint32),
                ("environment", C.c_int32), ("camera", MskCamera), ("media", C.POINTER(MskMedium)), ("nmedia", C.c_uint32),
                ("sensor_medium", C.c_int32)]


class MskRenderDesc(C.Structure):
    _fields_ = [("spp", C.c_uint32), ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32), ("max_depth", C.c_int32),
                ("rr_depth", C.c_int32), ("hide_emitters", C.c_int32), ("base_seed", C.c_uint64), ("clear_film", C.c_uint32),
                ("paths_per_batch", C.c_uint32), ("flags", C.c_uint32), ("integrator", C.c_uint32)]


class MskStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_render", C.c_float), ("ms_intersect", C.c_float), ("ms_shadow", C.c_float), ("ms_shade", C.c_float),
                ("ms_raygen", C.c_float), ("ms_film", C.c_float), ("bounces", C.c_uint32), ("batches", C.c_uint32),
                ("n_intersect_launches", C.c_uint32), ("n_shade_launches", C.c_uint32), ("n_shadow_launches", C.c_uint32),
                ("pad_", C.c_uint32), ("shaded_vertices", C.c_uint64), ("nodes_closest", C.c_uint64),
                ("tris_closest", C.c_uint64), ("nodes_shadow", C.c_uint64), ("tris_shadow", C.c_uint64),
                ("ms_sort", C.c_float), ("pad2_", C.c_uint32), ("tail_rays_closest", C.c_uint64), ("tail_rays_shadow", C.c_uint64),
                ("ms_tail", C.c_float), ("n_tail_launches", C.c_uint32)]


class MskAovDesc(C.Structure):
    _fields_ = [("types", C.POINTER(C.c_int32)), ("ntypes", C.c_uint32), ("pad_", C.c_uint32)]


def aov_desc(types) -> MskAovDesc:
    """`types`: MskAovType values or their names ("depth", "position", "uv", "geo_normal", "sh_normal", "integrator")."""
    ids = [AOV_NAMES[t] if isinstance(t, str) else int(t) for t in types]
    arr = (C.c_int32 * max(len(ids), 1))(*ids)
    d = MskAovDesc(C.cast(arr, C.POINTER(C.c_int32)), len(ids), 0)
    d._keep = arr
    return d


class MskAccelInfo(C.Structure):
    _fields_ = [("ntris", C.c_uint64), ("nnodes", C.c_uint64), ("node_bytes", C.c_uint64), ("tri_bytes", C.c_uint64),
                ("ms_build", C.c_float), ("sah_cost", C.c_float), ("max_depth", C.c_uint32), ("pad_", C.c_uint32)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32), ("geom", np.uint32)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 20

_lib = None


def load(path: os.PathLike | None = None) -> C.CDLL:
    """Load libmisaki_b200.so.  Raises if it is missing: the product path never falls back to CPU code."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise MskError(-3, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(str(p))
    lib.msk_gpu_last_error.restype = C.c_char_p
    lib.msk_gpu_stream.restype = C.c_void_p
    lib.msk_gpu_stream.argtypes = [C.c_void_p]
    lib.msk_gpu_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.msk_gpu_shutdown.argtypes = [C.c_void_p]
    lib.msk_gpu_shutdown.restype = None
    lib.msk_gpu_scene_create.argtypes = [C.c_void_p, C.POINTER(MskSceneDesc), C.POINTER(C.c_void_p)]
    lib.msk_gpu_scene_destroy.argtypes = [C.c_void_p]
    lib.msk_gpu_scene_destroy.restype = None
    lib.msk_gpu_accel_info.argtypes = [C.c_void_p, C.POINTER(MskAccelInfo)]
    lib.msk_gpu_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.msk_gpu_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.msk_gpu_intersect_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.msk_gpu_occluded_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.msk_gpu_intersect_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.msk_gpu_render.argtypes = [C.c_void_p, C.POINTER(MskRenderDesc), C.c_void_p, C.POINTER(MskStats)]
    lib.msk_gpu_render_dev.argtypes = [C.c_void_p, C.POINTER(MskRenderDesc), C.c_void_p, C.POINTER(MskStats)]
    lib.msk_gpu_render_reserve.argtypes = [C.c_void_p, C.POINTER(MskRenderDesc)]
    lib.msk_gpu_develop_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.msk_gpu_develop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.msk_gpu_aov_channels.argtypes = [C.POINTER(MskAovDesc)]
    lib.msk_gpu_render_aov.argtypes = [C.c_void_p, C.POINTER(MskRenderDesc), C.POINTER(MskAovDesc), C.c_void_p, C.POINTER(MskStats)]
    lib.msk_gpu_render_aov_dev.argtypes = [C.c_void_p, C.POINTER(MskRenderDesc), C.POINTER(MskAovDesc), C.c_void_p, C.POINTER(MskStats)]
    lib.msk_gpu_film_share_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.msk_gpu_film_share_ptr.argtypes = [C.c_void_p]
    lib.msk_gpu_film_share_ptr.restype = C.c_void_p
    lib.msk_gpu_film_share_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.msk_gpu_film_share_open.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.msk_gpu_reduce_film.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    lib.msk_gpu_film_share_check.argtypes = [C.c_void_p]
    lib.msk_gpu_film_share_destroy.argtypes = [C.c_void_p]
    lib.msk_gpu_film_share_destroy.restype = None
    lib.msk_gpu_device_count.argtypes = []
    lib.msk_gpu_film_share_attach.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32]
    lib.msk_gpu_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(MskRenderDesc), C.c_void_p, C.POINTER(MskStats)]
    if path is None:
        _lib = lib
    return lib


def check(lib, rc):
    if rc != 0:
        raise MskError(rc, (lib.msk_gpu_last_error() or b"").decode(errors="replace"))


def render_desc(spp, max_depth=-1, rr_depth=5, hide_emitters=False, base_seed=0, sample_begin=0, sample_end=None,
                clear_film=True, paths_per_batch=0, stage_timers=False, traversal_stats=False, integrator=0) -> MskRenderDesc:
    rd = MskRenderDesc()
    rd.spp = spp
    rd.sample_begin = sample_begin
    rd.sample_end = spp if sample_end is None else sample_end
    rd.max_depth = max_depth
    rd.rr_depth = rr_depth
    rd.hide_emitters = int(hide_emitters)
    rd.base_seed = base_seed
    rd.clear_film = int(clear_film)
    rd.paths_per_batch = paths_per_batch
    rd.flags = (RENDER_STAGE_TIMERS if stage_timers else 0) | (RENDER_TRAVERSAL_STATS if traversal_stats else 0)
    rd.integrator = {"path": INTEGRATOR_PATH, "volpath": INTEGRATOR_VOLPATH}.get(integrator, integrator)
    return rd


class Context:
    """One per GPU (msk_gpu_init / msk_gpu_shutdown)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.handle = C.c_void_p()
        check(self.lib, self.lib.msk_gpu_init(device, C.byref(self.handle)))
        self.device = device

    @property
    def stream(self) -> int:
        return self.lib.msk_gpu_stream(self.handle) or 0

    def close(self):
        if self.handle:
            self.lib.msk_gpu_shutdown(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Scene:
    """Device-resident scene + BVH (msk_gpu_scene_create)."""

    def __init__(self, ctx: Context, desc):
        self.ctx = ctx
        self.lib = ctx.lib
        self.desc = desc  # a scene.SceneDescription (keeps host arrays alive)
        self.handle = C.c_void_p()
        check(self.lib, self.lib.msk_gpu_scene_create(ctx.handle, C.byref(desc.c_desc()), C.byref(self.handle)))
        self.width, self.height = desc.width, desc.height

    def close(self):
        if self.handle:
            self.lib.msk_gpu_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def accel_info(self) -> MskAccelInfo:
        info = MskAccelInfo()
        check(self.lib, self.lib.msk_gpu_accel_info(self.handle, C.byref(info)))
        return info

    def intersect(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        check(self.lib, self.lib.msk_gpu_intersect(self.handle, rays.ctypes.data, hits.ctypes.data, rays.shape[0]))
        return hits

    def occluded(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        check(self.lib, self.lib.msk_gpu_occluded(self.handle, rays.ctypes.data, occ.ctypes.data, rays.shape[0]))
        return occ

    def intersect_stats(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        nodes = np.empty(rays.shape[0], dtype=np.uint32)
        tris = np.empty(rays.shape[0], dtype=np.uint32)
        check(self.lib, self.lib.msk_gpu_intersect_stats(self.handle, rays.ctypes.data, rays.shape[0], nodes.ctypes.data,
                                                         tris.ctypes.data))
        return nodes, tris

    def intersect_dev(self, d_rays: int, d_hits: int, n: int):
        check(self.lib, self.lib.msk_gpu_intersect_dev(self.handle, d_rays, d_hits, n))

    def occluded_dev(self, d_rays: int, d_occ: int, n: int):
        check(self.lib, self.lib.msk_gpu_occluded_dev(self.handle, d_rays, d_occ, n))

    def render(self, rd: MskRenderDesc, film: np.ndarray | None = None):
        """Host film in / out (H x W x 5 float32).  Returns (film, stats)."""
        if film is None:
            film = np.zeros((self.height, self.width, 5), dtype=np.float32)
        assert film.dtype == np.float32 and film.flags.c_contiguous and film.shape == (self.height, self.width, 5)
        stats = MskStats()
        check(self.lib, self.lib.msk_gpu_render(self.handle, C.byref(rd), film.ctypes.data, C.byref(stats)))
        return film, stats

    def reserve(self, rd: MskRenderDesc):
        """Allocate the path pools a render of `rd` will use now (msk_gpu_render_reserve)."""
        check(self.lib, self.lib.msk_gpu_render_reserve(self.handle, C.byref(rd)))

    def develop_dev(self, d_film: int, d_rgba: int):
        """HDRFilm::image on device buffers (msk_gpu_develop_dev), asynchronous on the context's stream."""
        check(self.lib, self.lib.msk_gpu_develop_dev(self.handle, d_film, d_rgba))

    def render_dev(self, rd: MskRenderDesc, d_film: int, want_stats: bool = True):
        stats = MskStats()
        check(self.lib, self.lib.msk_gpu_render_dev(self.handle, C.byref(rd), d_film, C.byref(stats) if want_stats else None))
        return stats

    def render_aov(self, rd: MskRenderDesc, types, film: np.ndarray | None = None):
        """The AOV integrator (reference integrators/aov.cpp).  Host film in / out, H x W x (5 + channels) float32:
        X,Y,Z,A,W then the AOV channels in the order of `types`.  Returns (film, stats)."""
        ad = aov_desc(types)
        nch = self.lib.msk_gpu_aov_channels(C.byref(ad))
        if nch < 0:
            check(self.lib, nch)
        if film is None:
            film = np.zeros((self.height, self.width, 5 + nch), dtype=np.float32)
        assert film.dtype == np.float32 and film.flags.c_contiguous and film.shape == (self.height, self.width, 5 + nch)
        stats = MskStats()
        check(self.lib, self.lib.msk_gpu_render_aov(self.handle, C.byref(rd), C.byref(ad), film.ctypes.data, C.byref(stats)))
        return film, stats

    def develop(self, film: np.ndarray) -> np.ndarray:
        film = np.ascontiguousarray(film[..., :5], dtype=np.float32)
        rgba = np.empty((self.height, self.width, 4), dtype=np.float32)
        check(self.lib, self.lib.msk_gpu_develop(self.handle, film.ctypes.data, rgba.ctypes.data))
        return rgba


def device_count() -> int:
    return int(load().msk_gpu_device_count())


def render_multi(scenes, rd: MskRenderDesc, film: np.ndarray | None = None):
    """msk_gpu_render_multi: `scenes` are capi.Scene objects of the same description, one per context / GPU; GPU i
    renders the i-th sub-range of rd's samples, the first scene's GPU sums the films over peer memory.  Host film out
    (H x W x 5 float32).  Returns (film, stats)."""
    s0 = scenes[0]
    if film is None:
        film = np.zeros((s0.height, s0.width, 5), dtype=np.float32)
    assert film.dtype == np.float32 and film.flags.c_contiguous and film.shape == (s0.height, s0.width, 5)
    arr = (C.c_void_p * len(scenes))(*[s.handle for s in scenes])
    stats = MskStats()
    check(s0.lib, s0.lib.msk_gpu_render_multi(arr, len(scenes), C.byref(rd), film.ctypes.data, C.byref(stats)))
    return film, stats


class FilmShare:
    """A film in exportable device memory + the NVLink peer reduction (msk_gpu_film_share_*, msk_gpu_reduce_film)."""
    IPC_HANDLE_BYTES = 64

    def __init__(self, ctx: "Context", shape):
        self.ctx, self.lib, self.shape = ctx, ctx.lib, tuple(int(x) for x in shape)
        n = 1
        for x in self.shape:
            n *= x
        self.nfloats = n = (n + 3) // 4 * 4  # the reduction moves float4s; the padding stays zero
        self.handle = C.c_void_p()
        check(self.lib, self.lib.msk_gpu_film_share_create(ctx.handle, n, C.byref(self.handle)))
        self.ptr = int(self.lib.msk_gpu_film_share_ptr(self.handle))
        self._epoch = 0

    @property
    def __cuda_array_interface__(self):  # lets torch.as_tensor(share, device="cuda") wrap the film without a copy
        return {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False), "version": 2, "strides": None}

    def export(self) -> bytes:
        buf = C.create_string_buffer(self.IPC_HANDLE_BYTES)
        check(self.lib, self.lib.msk_gpu_film_share_export(self.handle, buf))
        return buf.raw

    def open_peers(self, handles):
        blob = b"".join(handles)
        assert len(blob) == self.IPC_HANDLE_BYTES * len(handles)
        buf = C.create_string_buffer(blob, len(blob)) if blob else None
        check(self.lib, self.lib.msk_gpu_film_share_open(self.handle, buf, len(handles)))

    def reduce(self, is_root: bool, epoch: int | None = None):
        """Asynchronous on the context's stream; every rank calls it once per step with the same epoch."""
        if epoch is None:
            self._epoch = self._epoch % 0x7fffffff + 1
            epoch = self._epoch
        check(self.lib, self.lib.msk_gpu_reduce_film(self.handle, int(bool(is_root)), int(epoch)))

    def check(self):
        check(self.lib, self.lib.msk_gpu_film_share_check(self.handle))

    def close(self):
        if self.handle:
            self.lib.msk_gpu_film_share_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
