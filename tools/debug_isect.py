import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from misaki_render_b200 import capi
from oracle import pyoracle
from workloads import scenes
from tests.util import compare_hits, random_rays
sd = scenes.cbox(64, 64)
osc = pyoracle.OracleScene(sd)
rng = np.random.default_rng(1)
n=20000
s = np.stack([rng.random(n) * sd.width, rng.random(n) * sd.height, rng.random(n)], axis=-1).astype(np.float32)
rays = np.concatenate([osc.camera_rays(s), random_rays(20000, (0, 0, 0), (556, 548, 559), seed=2)])
with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
    gpu = sc.intersect(rays)
ref = osc.intersect(rays, brute_force=True)
t2, mb = osc.margin(rays)
r = compare_hits(gpu, ref, t2, mb)
print(r)
hit = np.isfinite(ref["t"])
bad = (np.isfinite(gpu["t"]) != hit) | (hit & np.isfinite(gpu["t"]) & ((gpu["prim"] != ref["prim"]) | (gpu["geom"] != ref["geom"]))) | (hit & np.isfinite(gpu["t"]) & (np.abs(gpu["t"]-ref["t"]) > 1e-5*ref["t"]))
for i in np.nonzero(bad)[0][:20]:
    print(i, rays[i], "gpu", gpu[i], "ref", ref[i], "t2", t2[i], "mb", mb[i])
