"""ctypes bindings of the C++ host front-end (lib/libmisaki_host.so, host/host_capi.h): load a misaki XML
scene file through the plugin system, inspect the flattened device description, render it.

The C++ side is the product's host layer (reference-facing: same XML format, plugin names, parameters and
error behaviour as misaki-render); this module only exists so that tests, bench.py and scripts can drive it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi

ROOT = Path(__file__).resolve().parent
LIB_PATH = ROOT / "lib" / "libmisaki_host.so"
_lib = None


class HostError(RuntimeError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise HostError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
        capi.load()  # libmisaki_b200.so first (same directory, also found through the rpath)
        L = C.CDLL(str(LIB_PATH))
        L.mskh_last_error.restype = C.c_char_p
        L.mskh_set_log_level.argtypes = [C.c_int]
        L.mskh_add_search_path.argtypes = [C.c_char_p]
        L.mskh_load_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.mskh_load_file_params.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(C.c_void_p)]
        L.mskh_load_string.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.mskh_free.argtypes = [C.c_void_p]
        L.mskh_free.restype = None
        L.mskh_scene_desc.argtypes = [C.c_void_p]
        L.mskh_scene_desc.restype = C.POINTER(capi.MskSceneDesc)
        L.mskh_render_desc.argtypes = [C.c_void_p, C.POINTER(capi.MskRenderDesc)]
        L.mskh_render.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(capi.MskStats)]
        L.mskh_registered_plugins.argtypes = [C.c_char_p, C.c_size_t]
        L.mskh_develop.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.mskh_develop_channels.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.mskh_write_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.mskh_write_pfm.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.mskh_srgb_model_fetch.argtypes = [C.c_void_p, C.c_void_p]
        L.mskh_add_search_path(str(ROOT).encode())  # data/srgb.coeff lives in the package
        L.mskh_set_log_level(3)
        _lib = L
    return _lib


def registered_plugins() -> list[str]:
    L = load()
    buf = C.create_string_buffer(4096)
    L.mskh_registered_plugins(buf, 4096)
    return buf.value.decode().split()


class HostScene:
    """A scene loaded by the C++ front-end (xml::load_file -> plugin instances)."""

    def __init__(self, path=None, xml: str | None = None, base_dir: str | None = None, params: dict | None = None):
        self.L = load()
        self.h = C.c_void_p()
        if path is not None:
            items = list((params or {}).items())
            names = (C.c_char_p * max(len(items), 1))(*[str(k).encode() for k, _ in items])
            values = (C.c_char_p * max(len(items), 1))(*[str(v).encode() for _, v in items])
            rc = self.L.mskh_load_file_params(str(path).encode(), names, values, len(items), C.byref(self.h))
        else:
            rc = self.L.mskh_load_string(xml.encode(), (base_dir or "").encode(), C.byref(self.h))
        if rc != 0:
            raise HostError(self.L.mskh_last_error().decode(errors="replace"))

    def close(self):
        if self.h:
            self.L.mskh_free(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def desc(self) -> capi.MskSceneDesc:
        p = self.L.mskh_scene_desc(self.h)
        if not p:
            raise HostError(self.L.mskh_last_error().decode(errors="replace"))
        return p.contents

    def render_desc(self) -> capi.MskRenderDesc:
        rd = capi.MskRenderDesc()
        if self.L.mskh_render_desc(self.h, C.byref(rd)) != 0:
            raise HostError(self.L.mskh_last_error().decode(errors="replace"))
        return rd

    def render(self, output: str | None = None) -> capi.MskStats:
        st = capi.MskStats()
        if self.L.mskh_render(self.h, output.encode() if output else None, C.byref(st)) != 0:
            raise HostError(self.L.mskh_last_error().decode(errors="replace"))
        return st

    # ---- views of the flattened description as numpy arrays (copies)
    def meshes(self):
        d = self.desc()
        out = []
        for i in range(d.nmeshes):
            m = d.meshes[i]
            v = np.ctypeslib.as_array(m.verts, shape=(m.nverts, 8)).copy() if m.nverts else np.zeros((0, 8), np.float32)
            t = np.ctypeslib.as_array(m.tris, shape=(m.ntris, 3)).copy() if m.ntris else np.zeros((0, 3), np.uint32)
            out.append(dict(verts=v, tris=t, bsdf=m.bsdf, emitter=m.emitter, has_normals=bool(m.has_normals), has_uvs=bool(m.has_uvs)))
        return out


def _check(rc):
    if rc != 0:
        raise HostError(load().mskh_last_error().decode(errors="replace"))


def develop(film: np.ndarray) -> np.ndarray:
    film = np.ascontiguousarray(film, dtype=np.float32)
    rgba = np.empty(film.shape[:2] + (4,), dtype=np.float32)
    _check(load().mskh_develop(film.ctypes.data, film.shape[0] * film.shape[1], rgba.ctypes.data))
    return rgba


def develop_channels(film: np.ndarray) -> np.ndarray:
    """HDRFilm::image of a film with AOV channels after X, Y, Z, A, W: RGBA, then every further channel / W."""
    film = np.ascontiguousarray(film, dtype=np.float32)
    out = np.empty(film.shape[:2] + (film.shape[2] - 1,), dtype=np.float32)
    _check(load().mskh_develop_channels(film.ctypes.data, film.shape[0] * film.shape[1], film.shape[2], out.ctypes.data))
    return out


def write_exr(path, rgba: np.ndarray):
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    _check(load().mskh_write_exr(str(path).encode(), rgba.ctypes.data, rgba.shape[1], rgba.shape[0]))


def write_pfm(path, rgba: np.ndarray):
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    _check(load().mskh_write_pfm(str(path).encode(), rgba.ctypes.data, rgba.shape[1], rgba.shape[0]))


def srgb_model_fetch(rgb) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.empty(3, dtype=np.float32)
    _check(load().mskh_srgb_model_fetch(rgb.ctypes.data, out.ctypes.data))
    return out


def read_exr_channels(path):
    """Reader for the uncompressed scanline float EXR files host/imageio.cpp writes (tests only).
    Returns (channel names in file order, H x W x C float32)."""
    import struct
    raw = Path(path).read_bytes()
    assert struct.unpack_from("<i", raw, 0)[0] == 20000630
    pos, attrs = 8, {}
    while raw[pos] != 0:
        e = raw.index(b"\0", pos); name = raw[pos:e].decode(); pos = e + 1
        e = raw.index(b"\0", pos); typ = raw[pos:e].decode(); pos = e + 1
        (size,) = struct.unpack_from("<i", raw, pos); pos += 4
        attrs[name] = (typ, raw[pos:pos + size]); pos += size
    pos += 1
    names, cl, cp = [], attrs["channels"][1], 0
    while cl[cp] != 0:
        e = cl.index(b"\0", cp); names.append(cl[cp:e].decode()); cp = e + 1
        assert struct.unpack_from("<i", cl, cp)[0] == 2  # FLOAT
        cp += 16
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    assert attrs["compression"][1] == b"\0"
    offsets = struct.unpack_from(f"<{h}Q", raw, pos)
    n = len(names)
    img = np.zeros((h, w, n), np.float32)
    for off in offsets:
        yy, nbytes = struct.unpack_from("<ii", raw, off)
        assert nbytes == w * n * 4
        img[yy - y0] = np.frombuffer(raw, dtype="<f4", count=w * n, offset=off + 8).reshape(n, w).T
    return names, img


def read_exr_rgba(path):
    names, img = read_exr_channels(path)
    return np.stack([img[..., names.index(c)] for c in "RGBA"], axis=-1)
