"""Several GPUs driven by ONE process (msk_gpu_render_multi, the host plugin's `devices` property / MSK_DEVICES): the
reference uses every core of the machine from one Integrator::render call (src/librender/integrator.cpp:54-75); here every
listed GPU renders a sample sub-range and the first GPU sums the films with the peer-memory kernel.  On a one-GPU box the
device list repeats device 0 (two contexts, two streams -- the protocol is the same); with >= 2 GPUs it is 0,1."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from misaki_render_b200 import capi, host_api
from workloads import scenes

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _device_list(n):
    have = capi.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("ndev,spp", [(2, 8), (3, 7)])
def test_render_multi_matches_one_device(gpu_ctx, ndev, spp):
    sd = scenes.cbox(64, 48)
    rd = capi.render_desc(spp=spp, max_depth=5)
    with capi.Scene(gpu_ctx, sd) as sc:
        film1, st1 = sc.render(rd)
    ctxs = [capi.Context(d) for d in _device_list(ndev)]
    try:
        scs = [capi.Scene(c, sd) for c in ctxs]
        try:
            film, st = capi.render_multi(scs, rd)
            film_again, _ = capi.render_multi(scs, rd)
        finally:
            for s in scs:
                s.close()
    finally:
        for c in ctxs:
            c.close()
    assert st.paths == st1.paths == 64 * 48 * spp
    assert st.rays_closest == st1.rays_closest and st.rays_shadow == st1.rays_shadow  # same paths, partitioned
    np.testing.assert_allclose(film, film1, rtol=2e-5, atol=1e-6)  # float summation order only
    np.testing.assert_array_equal(film, film_again)  # the peer sum adds in device order: deterministic


def test_render_multi_rejects_mismatched_scenes(gpu_ctx):
    a, b = scenes.cbox(32, 32), scenes.cbox(48, 32)
    c2 = capi.Context(0)
    try:
        with capi.Scene(gpu_ctx, a) as sa, capi.Scene(c2, b) as sb, capi.Scene(gpu_ctx, a) as sa2:
            with pytest.raises(capi.MskError, match="differs from scene 0"):
                capi.render_multi([sa, sb], capi.render_desc(spp=2, max_depth=3))
            with pytest.raises(capi.MskError, match="share a context"):
                capi.render_multi([sa, sa2], capi.render_desc(spp=2, max_depth=3))
    finally:
        c2.close()


def test_host_plugin_uses_the_listed_devices(tmp_path):
    """misaki_b200 scene.xml with MSK_DEVICES / <integer name="devices">: same image as one GPU."""
    exe = ROOT / "misaki_render_b200" / "lib" / "misaki_b200"
    args = [str(exe), str(ROOT / "assets" / "scenes" / "cbox.xml"), "-D", "w=48", "-D", "h=32", "-D", "spp=6", "-D", "depth=4"]
    env1 = {k: v for k, v in os.environ.items() if k != "MSK_DEVICES"}
    r1 = subprocess.run(args + ["-o", str(tmp_path / "one.exr")], capture_output=True, text=True, env=env1)
    assert r1.returncode == 0, r1.stderr
    devs = ",".join(str(d) for d in _device_list(2))
    r2 = subprocess.run(args + ["-o", str(tmp_path / "two.exr")], capture_output=True, text=True, env=dict(env1, MSK_DEVICES=devs))
    assert r2.returncode == 0, r2.stderr
    assert "2 GPUs, samples partitioned" in r2.stderr
    one, two = host_api.read_exr_rgba(tmp_path / "one.exr"), host_api.read_exr_rgba(tmp_path / "two.exr")
    np.testing.assert_allclose(two, one, rtol=2e-5, atol=1e-6)
    # a count larger than the box is an error, not a silent single-GPU render
    r3 = subprocess.run(args + ["-o", str(tmp_path / "x.exr")], capture_output=True, text=True, env=dict(env1, MSK_DEVICES="99"))
    assert r3.returncode != 0 and "visible" in r3.stderr


def test_scene_description_validation(gpu_ctx):
    """Malformed descriptions fail loudly instead of reading out of bounds on the device."""
    sd = scenes.cbox(16, 16)
    d = sd.c_desc()
    old = d.bsdfs[0].k
    d.bsdfs[0].k = 12345  # an id the diffuse BSDF never reads, but the texture pass would
    try:
        with pytest.raises(capi.MskError, match="out of range"):
            capi.Scene(gpu_ctx, sd)
    finally:
        d.bsdfs[0].k = old
    light = next(i for i in range(d.nmeshes) if d.meshes[i].emitter >= 0)
    other = next(i for i in range(d.nmeshes) if d.meshes[i].emitter < 0)
    d.meshes[other].emitter = d.meshes[light].emitter  # points at an area emitter of another mesh
    try:
        with pytest.raises(capi.MskError, match="not an area emitter of this mesh"):
            capi.Scene(gpu_ctx, sd)
    finally:
        d.meshes[other].emitter = -1
    with capi.Scene(gpu_ctx, sd):
        pass
