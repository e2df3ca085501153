#!/usr/bin/env python3
"""One render that reaches every traversal driver -- the packet kernel (camera rays of a scene that is not tiny), the lockstep
driver with its shared-memory pool of prepared rays (bounce >= 1 and the shadow queue: more rays than the static threshold),
the static driver and the tail kernel -- for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck python tools/sanitize_job.py"""
import hashlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from misaki_render_b200 import capi  # noqa: E402
from workloads import scenes  # noqa: E402

with capi.Context(0) as ctx, capi.Scene(ctx, scenes.bunny(256, 256, n=24)) as sc:
    film, st = sc.render(capi.render_desc(spp=32, max_depth=-1, rr_depth=5))
    print("film", hashlib.sha256(film.tobytes()).hexdigest()[:24], "rays", st.rays_closest, st.rays_shadow, "bounces", st.bounces)
