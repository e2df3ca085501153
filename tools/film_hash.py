#!/usr/bin/env python3
"""SHA-256 of the films a library build renders for a few jobs (MSK_B200_LIB selects the build): bit-identity A/B of kernel
rewrites that must not change a single bit.    python tools/film_hash.py"""
import hashlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from misaki_render_b200 import capi  # noqa: E402
from workloads import scenes  # noqa: E402

JOBS = [("cbox 64x48 12spp d5", lambda: scenes.cbox(64, 48), dict(spp=12, max_depth=5)),
        ("cbox 37x23 6spp (untiled film)", lambda: scenes.cbox(37, 23), dict(spp=6, max_depth=4)),
        ("bunny 128^2 32spp unbounded", lambda: scenes.bunny(128, 128, n=24), dict(spp=32, max_depth=-1, rr_depth=5)),
        ("teapot 96^2 16spp d8", lambda: scenes.teapot(96, 96, n=16), dict(spp=16, max_depth=8, rr_depth=3)),
        ("fog 64^2 8spp volpath", lambda: scenes.fog(64, 64), dict(spp=8, max_depth=-1, rr_depth=3, integrator="volpath"))]
with capi.Context(0) as ctx:
    for name, make, kw in JOBS:
        with capi.Scene(ctx, make()) as sc:
            film, st = sc.render(capi.render_desc(**kw))
            print(f"{name:34s} {hashlib.sha256(film.tobytes()).hexdigest()[:24]}  rays {st.rays_closest + st.rays_shadow}")
