"""Programmatic scene descriptions for the C ABI (the Python face of MskSceneDesc).

Mirrors, in miniature, what the reference's plugin constructors compute at load time:
  * PerspectiveCamera   src/librender/sensors/perspective.cpp:9-20
  * Transform4f::lookat / perspective   include/misaki/core/transform.h:170-188
  * GaussianFilter + init_discretization   filters/gaussian.cpp:9-20, rfilter.cpp:12-27
  * srgb / srgb_d65 / d65 / uniform spectra   src/librender/spectra/*.cpp
The XML front-end (host/ C++) produces the same structures from the reference's scene files.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi
from .rgb2spec import model as rgb2spec_model

f32 = np.float32

# CIE D65, 360..830 nm @ 5 nm (public CIE data; same rows as csrc/spectral_tables.h)
def _load_d65():
    import re
    from pathlib import Path
    txt = (Path(__file__).resolve().parent / "csrc" / "spectral_tables.h").read_text()
    rows = re.findall(r"\{\s*([0-9.eE+-]+)f,\s*([0-9.eE+-]+)f,\s*([0-9.eE+-]+)f,\s*([0-9.eE+-]+)f\s*\}", txt)
    assert len(rows) == 95
    return np.array([r[3] for r in rows], dtype=f32), np.array([[r[0], r[1], r[2]] for r in rows], dtype=f32)


D65_TABLE, CIE_XYZ = _load_d65()


def lookat(origin, target, up):
    """Transform4f::lookat, transform.h:170-179 (float32)."""
    o, t, u = (np.asarray(v, dtype=f32) for v in (origin, target, up))
    def nrm(v):
        return (v / f32(np.sqrt(np.dot(v, v), dtype=f32))).astype(f32)
    d = nrm(t - o)
    left = nrm(np.cross(nrm(u), d).astype(f32))
    new_up = nrm(np.cross(d, left).astype(f32))
    m = np.eye(4, dtype=f32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, o
    return m


def translate(v):
    m = np.eye(4, dtype=f32)
    m[:3, 3] = np.asarray(v, dtype=f32)
    return m


def scale(v):
    return np.diag(np.array([v[0], v[1], v[2], 1.0], dtype=f32))


def perspective(fov, near, far):
    """Transform4f::perspective, transform.h:181-188."""
    recip = f32(1.0) / (f32(far) - f32(near))
    cot = f32(1.0) / f32(math.tan(float(f32(fov) / f32(2.0) * f32(math.pi / 180))))
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0] = cot; m[1, 1] = cot; m[2, 2] = f32(far) * recip; m[2, 3] = -f32(near) * f32(far) * recip; m[3, 2] = 1
    return m


def _inv(m):
    return np.linalg.inv(m.astype(np.float64)).astype(f32)


def gaussian_filter(stddev=0.5):
    """GaussianFilter ctor + ReconstructionFilter::init_discretization (float32)."""
    stddev = f32(stddev)
    radius = f32(4) * stddev
    alpha = f32(-1.0) / (f32(2.0) * stddev * stddev)
    bias = f32(math.exp(float(alpha * radius * radius)))
    res = 32
    vals = np.zeros(res + 1, dtype=f32)
    s = f32(0)
    for i in range(res):
        x = f32(f32(radius * f32(i)) / f32(res))
        vals[i] = max(f32(0), f32(f32(math.exp(float(f32(alpha * x) * x))) - bias))
        s = f32(s + vals[i])
    s = f32(s * f32(f32(2) * radius / f32(res)))
    vals[:res] = (vals[:res] * f32(f32(1.0) / s)).astype(f32)
    return float(radius), vals


class SceneDescription:
    def __init__(self, width, height, fov=30.0, to_world=None, near_clip=1e-2, far_clip=1e4, filter_stddev=0.5):
        self.width, self.height = int(width), int(height)
        self.fov, self.near_clip, self.far_clip = float(fov), float(near_clip), float(far_clip)
        self.to_world = np.eye(4, dtype=f32) if to_world is None else np.asarray(to_world, dtype=f32)
        self.filter_radius, self.filter_table = gaussian_filter(filter_stddev)
        self.spectra: list[capi.MskSpectrum] = []
        self.tables: list[np.ndarray] = []
        self.bsdfs: list[capi.MskBsdf] = []
        self.emitters: list[capi.MskEmitter] = []
        self.meshes: list[dict] = []
        self.environment = -1
        self.media: list[capi.MskMedium] = []
        self.sensor_medium = -1
        self._keep = []
        self._cdesc = None

    # ---- spectra -----------------------------------------------------------------
    def _add_spec(self, kind, c=(0, 0, 0), value=0.0, table=None, lmin=0.0, lmax=0.0):
        s = capi.MskSpectrum()
        s.kind = kind
        s.c[:] = [float(x) for x in c]
        s.value = float(value)
        if table is not None:
            s.table_offset = int(sum(t.size for t in self.tables))
            s.table_size = int(table.size)
            self.tables.append(np.asarray(table, dtype=f32))
            s.lambda_min, s.lambda_max = float(lmin), float(lmax)
        self.spectra.append(s)
        self._cdesc = None
        return len(self.spectra) - 1

    def spectrum_checkerboard(self, color0, color1, to_uv=None):
        """"checkerboard" texture (textures/checkerboard.cpp:11-31) over two already declared spectra; `to_uv` is the
        4x4 "to_uv" transform, of which Transform4f::extract keeps the top-left 3x3 (transform.h:142-148)."""
        m = np.eye(4, dtype=f32) if to_uv is None else np.asarray(to_uv, dtype=f32).reshape(4, 4)
        sid = self._add_spec(capi.SPEC_CHECKERBOARD)
        s = self.spectra[sid]
        s.child0, s.child1 = int(color0), int(color1)
        s.to_uv[:] = [float(x) for x in m[:2, :3].reshape(-1)]
        return sid

    def spectrum_uniform(self, value):
        return self._add_spec(capi.SPEC_UNIFORM, value=value)

    def spectrum_srgb(self, rgb):
        """<rgb> outside an emitter -> "srgb" texture (xml.cpp:269-277, spectra/srgb.cpp:14-23)."""
        return self._add_spec(capi.SPEC_SRGB, c=rgb2spec_model().fetch(rgb))

    def spectrum_d65(self, scale=1.0):
        """Texture::D65 / "d65" expanded into "regular" (spectra/d65.cpp:29-50)."""
        m_scale = f32(f32(scale) * f32(f32(1.0) / f32(10568.0)))
        return self._add_spec(capi.SPEC_REGULAR, table=(D65_TABLE * m_scale).astype(f32), lmin=360.0, lmax=830.0)

    def spectrum_srgb_d65(self, rgb, scale=1.0):
        """<rgb> inside an emitter -> "srgb_d65" (spectra/srgb_d65.cpp:14-36)."""
        color = np.asarray(rgb, dtype=f32)
        s = f32(color.max() * f32(2.0))
        if s != 0:
            color = (color / s).astype(f32)
        coeff = rgb2spec_model().fetch(color)
        m_scale = f32(f32(f32(scale) * s) * f32(f32(1.0) / f32(10568.0)))
        return self._add_spec(capi.SPEC_SRGB_D65, c=coeff, table=(D65_TABLE * m_scale).astype(f32), lmin=360.0, lmax=830.0)

    def spectrum_unbounded(self, rgb):
        """Builder decision for conductor eta / k given as <rgb> (values may exceed 1; the reference's
        eval_3 path is unimplemented): scale = 2 max(rgb) as srgb_d65.cpp:18-23 does,
        value(lambda) = scale * srgb_model_eval(fetch(rgb / scale))."""
        color = np.asarray(rgb, dtype=f32)
        s = f32(color.max() * f32(2.0))
        if s != 0:
            color = (color / s).astype(f32)
        return self._add_spec(capi.SPEC_SRGB_UNBOUNDED, c=rgb2spec_model().fetch(color), value=s)

    def _as_spectrum(self, v, unbounded=False):
        if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
            return int(v)
        if isinstance(v, float):
            return self.spectrum_uniform(v)
        return self.spectrum_unbounded(v) if unbounded else self.spectrum_srgb(v)

    # ---- bsdfs ---------------------------------------------------------------------
    def _add_bsdf(self, type_, reflectance, transmittance=-1, eta=-1, k=-1, alpha=(0.1, 0.1), ior=(1.5046, 1.00028),
                  distribution="ggx", sample_visible=False, twosided=False):
        b = capi.MskBsdf()
        b.type = type_
        b.reflectance, b.transmittance, b.eta, b.k = reflectance, transmittance, eta, k
        b.alpha_u, b.alpha_v = float(alpha[0]), float(alpha[1])
        b.int_ior, b.ext_ior = float(ior[0]), float(ior[1])
        b.distribution = {"beckmann": 0, "ggx": 1}[distribution]
        b.sample_visible = int(sample_visible)
        b.twosided = int(twosided)
        self.bsdfs.append(b)
        self._cdesc = None
        return len(self.bsdfs) - 1

    def bsdf_diffuse(self, reflectance=(0.5, 0.5, 0.5), twosided=False):
        return self._add_bsdf(capi.BSDF_DIFFUSE, self._as_spectrum(reflectance), twosided=twosided)

    def bsdf_conductor(self, eta, k, specular_reflectance=(1.0, 1.0, 1.0), twosided=False):
        return self._add_bsdf(capi.BSDF_CONDUCTOR, self._as_spectrum(specular_reflectance), eta=self._as_spectrum(eta, True),
                              k=self._as_spectrum(k, True), twosided=twosided)

    def bsdf_roughconductor(self, eta, k, alpha=0.1, specular_reflectance=(1.0, 1.0, 1.0), distribution="ggx",
                            sample_visible=False, twosided=False):
        a = (alpha, alpha) if np.isscalar(alpha) else alpha
        return self._add_bsdf(capi.BSDF_ROUGHCONDUCTOR, self._as_spectrum(specular_reflectance),
                              eta=self._as_spectrum(eta, True), k=self._as_spectrum(k, True), alpha=a,
                              distribution=distribution, sample_visible=sample_visible, twosided=twosided)

    def bsdf_roughdielectric(self, int_ior=1.5046, ext_ior=1.00028, alpha=0.1, specular_reflectance=(1.0, 1.0, 1.0),
                             specular_transmittance=(1.0, 1.0, 1.0), distribution="ggx", sample_visible=False):
        a = (alpha, alpha) if np.isscalar(alpha) else alpha
        return self._add_bsdf(capi.BSDF_ROUGHDIELECTRIC, self._as_spectrum(specular_reflectance),
                              transmittance=self._as_spectrum(specular_transmittance), alpha=a, ior=(int_ior, ext_ior),
                              distribution=distribution, sample_visible=sample_visible)

    def bsdf_dielectric(self, int_ior=1.49, ext_ior=1.00028, specular_reflectance=(1.0, 1.0, 1.0),
                        specular_transmittance=(1.0, 1.0, 1.0)):
        return self._add_bsdf(capi.BSDF_DIELECTRIC, self._as_spectrum(specular_reflectance),
                              transmittance=self._as_spectrum(specular_transmittance), ior=(int_ior, ext_ior))

    # ---- shapes / emitters ------------------------------------------------------------
    def add_medium(self, sigma_a, sigma_s, scale=1.0):
        """"homogeneous" medium (media/homogeneous.cpp:12-19) with the default isotropic phase function
        (medium.cpp:24-29).  sigma_a / sigma_s: rgb tuples (-> unbounded spectra: coefficients are not reflectances),
        floats (uniform) or spectrum ids."""
        m = capi.MskMedium()
        m.sigma_a, m.sigma_s = self._as_spectrum(sigma_a, True), self._as_spectrum(sigma_s, True)
        m.phase, m.scale = 0, float(scale)
        self.media.append(m)
        self._cdesc = None
        return len(self.media) - 1

    def add_mesh(self, verts, tris, bsdf, radiance=None, has_normals=False, has_uvs=False, interior_medium=-1, exterior_medium=-1):
        """verts: (N, 8) or (N, 3) float32 world-space; tris: (M, 3) uint32.  radiance: rgb tuple or a
        spectrum id -> attaches an area emitter (emitters/area.cpp).  interior/exterior_medium: shape.cpp:28-39."""
        verts = np.asarray(verts, dtype=f32)
        if verts.shape[1] == 3:
            v8 = np.zeros((verts.shape[0], 8), dtype=f32)
            v8[:, :3] = verts
            verts = v8
        verts = np.ascontiguousarray(verts, dtype=f32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        emitter = -1
        if radiance is not None:
            spec = radiance if isinstance(radiance, (int, np.integer)) else self.spectrum_srgb_d65(radiance)
            e = capi.MskEmitter()
            e.type, e.radiance, e.shape = capi.EMITTER_AREA, spec, len(self.meshes)
            self.emitters.append(e)
            emitter = len(self.emitters) - 1
        self.meshes.append(dict(verts=verts, tris=tris, bsdf=bsdf, emitter=emitter, has_normals=has_normals, has_uvs=has_uvs,
                                interior_medium=int(interior_medium), exterior_medium=int(exterior_medium)))
        self._cdesc = None
        return len(self.meshes) - 1

    def add_constant_environment(self, radiance):
        spec = radiance if isinstance(radiance, (int, np.integer)) else self.spectrum_srgb_d65(radiance)
        e = capi.MskEmitter()
        e.type, e.radiance, e.shape = capi.EMITTER_CONSTANT, spec, -1
        self.emitters.append(e)
        self.environment = len(self.emitters) - 1
        self._cdesc = None
        return self.environment

    # ---- flattening -------------------------------------------------------------------
    def camera(self) -> capi.MskCamera:
        cam = capi.MskCamera()
        aspect = f32(self.width) / f32(self.height)
        # perspective.cpp:12-19.  Transform4f products carry the inverse as the product of the
        # factors' inverses (transform.h:100-103), which is what sample_to_camera is.
        factors = [scale((self.width, self.height, 1.0)), scale((-0.5, float(f32(-0.5) * aspect), 1.0)),
                   translate((-1.0, float(f32(-1.0) / aspect), 0.0)), perspective(self.fov, self.near_clip, self.far_clip)]
        inv = np.eye(4, dtype=f32)
        for fct in factors:  # (A B C D)^-1 = D^-1 C^-1 B^-1 A^-1, accumulated as t.inv * this.inv
            inv = (_inv(fct) @ inv).astype(f32)
        cam.sample_to_camera[:] = [float(x) for x in inv.reshape(-1)]
        cam.to_world[:] = [float(x) for x in self.to_world.reshape(-1)]
        cam.near_clip, cam.far_clip = self.near_clip, self.far_clip
        cam.width, cam.height = self.width, self.height
        cam.filter_radius = self.filter_radius
        cam.filter_table[:] = [float(x) for x in self.filter_table]
        return cam

    def c_desc(self) -> capi.MskSceneDesc:
        if self._cdesc is not None:
            return self._cdesc
        d = capi.MskSceneDesc()
        meshes = (capi.MskMesh * max(len(self.meshes), 1))()
        for i, m in enumerate(self.meshes):
            meshes[i].verts = m["verts"].ctypes.data_as(C.POINTER(C.c_float))
            meshes[i].tris = m["tris"].ctypes.data_as(C.POINTER(C.c_uint32))
            meshes[i].nverts, meshes[i].ntris = m["verts"].shape[0], m["tris"].shape[0]
            meshes[i].bsdf, meshes[i].emitter = m["bsdf"], m["emitter"]
            meshes[i].has_normals, meshes[i].has_uvs = int(m["has_normals"]), int(m["has_uvs"])
            meshes[i].interior_medium, meshes[i].exterior_medium = m.get("interior_medium", -1), m.get("exterior_medium", -1)
        bsdfs = (capi.MskBsdf * max(len(self.bsdfs), 1))(*self.bsdfs)
        emitters = (capi.MskEmitter * max(len(self.emitters), 1))(*self.emitters)
        spectra = (capi.MskSpectrum * max(len(self.spectra), 1))(*self.spectra)
        tables = np.concatenate(self.tables).astype(f32) if self.tables else np.zeros(1, dtype=f32)
        d.meshes, d.nmeshes = meshes, len(self.meshes)
        d.bsdfs, d.nbsdfs = bsdfs, len(self.bsdfs)
        d.emitters, d.nemitters = emitters, len(self.emitters)
        d.spectra, d.nspectra = spectra, len(self.spectra)
        d.spectrum_tables = tables.ctypes.data_as(C.POINTER(C.c_float))
        d.ntable_floats = int(tables.size) if self.tables else 0
        d.environment = self.environment
        d.camera = self.camera()
        media = (capi.MskMedium * max(len(self.media), 1))(*self.media)
        d.media, d.nmedia, d.sensor_medium = media, len(self.media), int(self.sensor_medium)
        self._keep = [meshes, bsdfs, emitters, spectra, tables, media]
        self._cdesc = d
        return d
