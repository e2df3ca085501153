"""misaki_render_b200 -- B200-native backend for misaki-render's path-tracing hot path.

The product is the CUDA library ``lib/libmisaki_b200.so`` behind the C ABI declared in
``include/misaki_b200.h``; this package is the thin Python host layer (ctypes bindings,
scene-description builders, the C++ host front-end loader).  Importing the package does
not load any native code; ``capi.load()`` does, and fails loudly when the library is
missing -- there is no CPU fallback.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
