"""Host-side RGB -> spectral-coefficient lookup (setup time, once per texture).

Restates ``rgb2spec_load`` / ``rgb2spec_fetch`` of the reference's vendored
ext/rgb2spec/rgb2spec.c:12-47,59-119 (Jakob & Hanika 2019) in numpy float32.  The
table file is the one the reference generates at build time with ``rgb2spec_opt 64``
(ext/rgb2spec/CMakeLists.txt:41-46) and loads as data/srgb.coeff (srgb.cpp:14-18).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

DATA_DIR = Path(__file__).resolve().parent / "data"
DEFAULT_COEFF = DATA_DIR / "srgb.coeff"
f32 = np.float32


class RGB2Spec:
    def __init__(self, path=None):
        path = Path(path) if path else DEFAULT_COEFF
        if not path.exists():
            raise FileNotFoundError(f"Could not load sRGB-to-spectrum upsampling model ('{path}'); run build() first")
        raw = path.read_bytes()
        if raw[:4] != b"SPEC":
            raise ValueError("malformed coefficient file")
        (self.res,) = struct.unpack_from("<I", raw, 4)
        res = self.res
        self.scale = np.frombuffer(raw, dtype="<f4", count=res, offset=8)
        self.data = np.frombuffer(raw, dtype="<f4", count=3 * res ** 3 * 3, offset=8 + 4 * res)

    def _find_interval(self, x):
        left, last, size = 0, self.res - 2, self.res - 2
        while size > 0:
            half = size >> 1
            middle = left + half + 1
            if self.scale[middle] <= x:
                left = middle
                size -= half + 1
            else:
                size = half
        return min(left, last)

    def fetch(self, rgb):
        """rgb2spec_fetch: three float32 polynomial coefficients for an sRGB colour (clamped to [0,1])."""
        res = self.res
        c = [f32(max(min(float(f32(v)), 1.0), 0.0)) for v in rgb]
        i = 0
        for j in (1, 2):
            if c[j] >= c[i]:
                i = j
        z = c[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = f32(f32(res - 1) / z)
            x = f32(c[(i + 1) % 3] * scale)
            y = f32(c[(i + 2) % 3] * scale)
        def to_u32(v):  # C cast float -> uint32 (NaN/negative are UB in C; clamp like x86 does for the cases that occur)
            v = float(v)
            return 0 if not np.isfinite(v) or v < 0 else int(v)
        xi = min(to_u32(x), res - 2)
        yi = min(to_u32(y), res - 2)
        zi = self._find_interval(z)
        offset = (((i * res + zi) * res + yi) * res + xi) * 3
        dx, dy, dz = 3, 3 * res, 3 * res * res
        x1 = f32(x - f32(xi)); x0 = f32(f32(1) - x1)
        y1 = f32(y - f32(yi)); y0 = f32(f32(1) - y1)
        z1 = f32((z - self.scale[zi]) / (self.scale[zi + 1] - self.scale[zi])); z0 = f32(f32(1) - z1)
        D = self.data
        out = np.empty(3, dtype=f32)
        with np.errstate(invalid="ignore"):
            for j in range(3):
                o = offset + j
                a = f32(f32(f32(D[o] * x0) + f32(D[o + dx] * x1)) * y0)
                b = f32(f32(f32(D[o + dy] * x0) + f32(D[o + dy + dx] * x1)) * y1)
                c_ = f32(f32(f32(D[o + dz] * x0) + f32(D[o + dz + dx] * x1)) * y0)
                d = f32(f32(f32(D[o + dz + dy] * x0) + f32(D[o + dz + dy + dx] * x1)) * y1)
                out[j] = f32(f32(f32(a + b) * z0) + f32(f32(c_ + d) * z1))
        return out


_model = None


def model() -> RGB2Spec:
    global _model
    if _model is None:
        _model = RGB2Spec()
    return _model


def srgb_model_eval(coeff, wavelengths):
    """include/misaki/render/srgb.h:8-19 (numpy float32; host-side checks only)."""
    c = np.asarray(coeff, dtype=f32)
    w = np.asarray(wavelengths, dtype=f32)
    if np.isinf(c[2]):
        return np.full(w.shape, f32(np.copysign(1.0, c[2]) * 0.5 + 0.5), dtype=f32)
    v = (c[0] * w + c[1]) * w + c[2]
    return np.maximum(f32(0.5) * v / np.sqrt(v * v + f32(1)) + f32(0.5), f32(0))
