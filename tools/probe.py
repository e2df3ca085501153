#!/usr/bin/env python3
"""Quick throughput probe of the BASELINE configs (development aid, not the bench)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from misaki_render_b200 import capi
from workloads import scenes

def run(name, sd, rd, reps=3):
    with capi.Context(0) as ctx:
        t0 = time.time()
        with capi.Scene(ctx, sd) as sc:
            t_scene = time.time() - t0
            info = sc.accel_info()
            best = None
            for _ in range(reps):
                film, st = sc.render(rd)
                if best is None or st.ms_render < best.ms_render:
                    best = st
            st = best
            rays = st.rays_closest + st.rays_shadow
            print(f"{name}: tris={info.ntris} nodes={info.nnodes} depth={info.max_depth} build={info.ms_build:.2f}ms scene_create={t_scene*1e3:.1f}ms | "
                  f"paths={st.paths} ms={st.ms_render:.2f} Mpaths/s={st.paths/st.ms_render/1e3:.1f} Mrays/s={rays/st.ms_render/1e3:.1f} "
                  f"rays/path={rays/st.paths:.2f} launches={st.kernel_launches} bounces={st.bounces} batches={st.batches}", flush=True)

which = sys.argv[1:] or ["c1", "c2", "c3"]
if "c1" in which:
    run("C1 cbox 256x256x16 d5", scenes.cbox(256, 256), capi.render_desc(spp=16, max_depth=5))
    run("C1x cbox 1024x1024x16 d5", scenes.cbox(1024, 1024), capi.render_desc(spp=16, max_depth=5))
if "c2" in which:
    run("C2 bunny 512x512x64", scenes.bunny(512, 512), capi.render_desc(spp=64, max_depth=-1))
if "c3" in which:
    run("C3 teapot 1024x1024x64 d16", scenes.teapot(1024, 1024), capi.render_desc(spp=64, max_depth=16))
