"""bench.py's CPU arm times a speed build of the oracle (oracle/Makefile `fast`: -O3 -march=native, FMA, and -DORC_FAST: an
ordered walk over two-box nodes and pre-gathered triangles instead of the checker's unordered BVH2 walk).  What ties that
timed arm to the checked one: the same walk compiled with the checker's own flags (`fastcheck`) must return the checker's
records and films bit for bit."""
import hashlib
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

CHILD = r"""
import hashlib, sys
from pathlib import Path
sys.path.insert(0, %r)
from oracle import pyoracle
if sys.argv[1] != "checker":
    pyoracle.LIB = Path(%r) / "oracle" / "_build" / "liboracle_fastcheck.so"
import numpy as np
from misaki_render_b200 import capi
from workloads import scenes
out = []
for sd, kw in [(scenes.bunny(48, 48, n=12), dict(max_depth=-1, rr_depth=3)), (scenes.teapot(40, 40, n=10), dict(max_depth=6, rr_depth=3)),
               (scenes.cbox(32, 32), dict(max_depth=5, rr_depth=5))]:
    osc = pyoracle.OracleScene(sd)
    film, st = osc.render(capi.render_desc(spp=4, **kw), nthreads=2)
    rays = scenes.primary_rays(sd, 48)
    hits = osc.intersect(rays)
    occ = osc.occluded(rays)
    out.append(hashlib.sha256(film.tobytes() + hits.tobytes() + np.asarray(occ).tobytes()).hexdigest() + ":%%d:%%d" %% (st.rays_closest, st.rays_shadow))
    osc.close()
print(" ".join(out))
""" % (str(ROOT), str(ROOT))


def _run(which):
    r = subprocess.run([sys.executable, "-c", CHILD, which], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-1].split()


def test_fast_walk_returns_the_checkers_records_bit_for_bit():
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "all", "fastcheck"], check=True, capture_output=True, timeout=600)
    a, b = _run("checker"), _run("fastcheck")
    assert len(a) == 3 and a == b
