"""GPU closest-hit / any-hit parity against the oracle (brute-force Moeller-Trumbore), through the C ABI."""
import numpy as np
import pytest

from misaki_render_b200 import capi
from oracle import pyoracle
from workloads import scenes
from tests.util import compare_hits, random_rays

pytestmark = pytest.mark.gpu


def _camera_rays(sd, osc, n, seed):
    rng = np.random.default_rng(seed)
    s = np.stack([rng.random(n) * sd.width, rng.random(n) * sd.height, rng.random(n)], axis=-1).astype(np.float32)
    return osc.camera_rays(s)


def test_cbox_primary_and_random(gpu_ctx):
    sd = scenes.cbox(64, 64)
    osc = pyoracle.OracleScene(sd)
    rays = np.concatenate([_camera_rays(sd, osc, 20000, 1), random_rays(20000, (0, 0, 0), (556, 548, 559), seed=2)])
    with capi.Scene(gpu_ctx, sd) as sc:
        gpu = sc.intersect(rays)
        occ = sc.occluded(rays)
    ref = osc.intersect(rays, brute_force=True)
    t2, mb = osc.margin(rays)
    r = compare_hits(gpu, ref, t2, mb, rays)
    assert r["mismatches"] == 0, r
    assert r["hits"] > 0.5 * r["n"]
    nondeg = np.isfinite(ref["t"]) & (mb > 1e-6)
    assert ((occ != 0) == np.isfinite(ref["t"]))[nondeg | ~np.isfinite(ref["t"])].all()


def test_bunny_secondary(gpu_ctx):
    sd = scenes.bunny(64, 64)
    osc = pyoracle.OracleScene(sd)
    cam = _camera_rays(sd, osc, 3000, 3)
    with capi.Scene(gpu_ctx, sd) as sc:
        gpu = sc.intersect(cam)
        ref = osc.intersect(cam, brute_force=True)
        t2, mb = osc.margin(cam)
        r = compare_hits(gpu, ref, t2, mb, cam)
        assert r["mismatches"] == 0, r
        rnd = random_rays(3000, (-2, 0.05, -2), (2, 2.5, 2), seed=4)
        gpu2 = sc.intersect(rnd)
        info = sc.accel_info()
    ref2 = osc.intersect(rnd, brute_force=True)
    t2, mb = osc.margin(rnd)
    r2 = compare_hits(gpu2, ref2, t2, mb, rnd)
    assert r2["mismatches"] == 0, r2
    assert info.ntris == sum(m["tris"].shape[0] for m in sd.meshes)


def test_edge_cases(gpu_ctx):
    sd = scenes.cbox(32, 32)
    with capi.Scene(gpu_ctx, sd) as sc:
        assert sc.intersect(np.zeros(0, dtype=capi.RAY_DTYPE)).shape == (0,)
        rays = random_rays(64, (100, 100, 100), (400, 400, 400), seed=5, tmax=1e-3)  # tmax before anything
        assert not np.isfinite(sc.intersect(rays)["t"]).any()
        assert not sc.occluded(rays).any()
        # axis-aligned rays with zero direction components
        rays = np.zeros(3, dtype=capi.RAY_DTYPE)
        rays["o"] = (278, 273, 100)
        rays["d"] = [(0, 1, 0), (0, -1, 0), (0, 0, 1)]
        rays["tmin"], rays["tmax"] = 0, np.inf
        h = sc.intersect(rays)
        osc = pyoracle.OracleScene(sd)
        ref = osc.intersect(rays, brute_force=True)
        assert (h["prim"] == ref["prim"]).all() and (h["geom"] == ref["geom"]).all()
        assert np.allclose(h["t"], ref["t"], rtol=1e-5)


@pytest.mark.parametrize("builder", ["lbvh", "ploc"])
def test_sah_optimal_collapse_same_hits_fewer_nodes(gpu_ctx, monkeypatch, builder):
    """The wide tree is collapsed from the binary one by dynamic programming over the SAH cost (k_plan, msk_bvh.cu) instead
    of greedily opening the largest child (MSK_BVH_COLLAPSE=greedy): fewer, fuller nodes over the same triangles -- same
    closest hits, same occlusion, under both binary builders."""
    sd = scenes.bunny(64, 64)
    osc = pyoracle.OracleScene(sd)
    rays = np.concatenate([_camera_rays(sd, osc, 3000, 17), random_rays(3000, (-2, 0.05, -2), (2, 2.5, 2), seed=18)])
    monkeypatch.setenv("MSK_BVH_BUILDER", builder)
    with capi.Scene(gpu_ctx, sd) as sc:
        sah, sah_occ, si = sc.intersect(rays), sc.occluded(rays), sc.accel_info()
    monkeypatch.setenv("MSK_BVH_COLLAPSE", "greedy")
    with capi.Scene(gpu_ctx, sd) as sc:
        gr, gr_occ, gi = sc.intersect(rays), sc.occluded(rays), sc.accel_info()
    assert si.ntris == gi.ntris and si.nnodes < gi.nnodes
    ref = osc.intersect(rays, brute_force=True)
    t2, mb = osc.margin(rays)
    assert compare_hits(sah, ref, t2, mb, rays)["mismatches"] == 0
    assert compare_hits(gr, ref, t2, mb, rays)["mismatches"] == 0
    nondeg = mb > 1e-6
    assert (sah["prim"] == gr["prim"])[nondeg].all()
    assert (sah_occ == gr_occ)[nondeg | ~np.isfinite(ref["t"])].all()


def test_ploc_builder_same_hits(gpu_ctx, monkeypatch):
    """MSK_BVH_BUILDER=ploc (SAH-driven clustering instead of the Morton-prefix tree): another tree over the same
    triangles must report the same closest hits and the same occlusion."""
    sd = scenes.bunny(64, 64)
    osc = pyoracle.OracleScene(sd)
    rays = np.concatenate([_camera_rays(sd, osc, 3000, 7), random_rays(3000, (-2, 0.05, -2), (2, 2.5, 2), seed=8)])
    with capi.Scene(gpu_ctx, sd) as sc:
        base, base_occ, lb = sc.intersect(rays), sc.occluded(rays), sc.accel_info()
    monkeypatch.setenv("MSK_BVH_BUILDER", "ploc")
    with capi.Scene(gpu_ctx, sd) as sc:
        ploc, ploc_occ, pl = sc.intersect(rays), sc.occluded(rays), sc.accel_info()
    assert pl.ntris == lb.ntris and pl.sah_cost > 0
    ref = osc.intersect(rays, brute_force=True)
    t2, mb = osc.margin(rays)
    assert compare_hits(ploc, ref, t2, mb, rays)["mismatches"] == 0
    nondeg = mb > 1e-6
    assert (ploc["prim"] == base["prim"])[nondeg].all()
    assert (ploc_occ == base_occ)[nondeg | ~np.isfinite(ref["t"])].all()
