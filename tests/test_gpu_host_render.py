"""End to end through the C++ host front-end on the GPU: XML scene file -> plugin objects -> GpuPathIntegrator::render
(C ABI) -> Film::put -> Film::develop (EXR), against the same scene described programmatically and rendered
through the C ABI directly, and against the CPU oracle."""
from pathlib import Path

import numpy as np
import pytest

from misaki_render_b200 import capi, host_api
from oracle import pyoracle
from workloads import scenes
from tests.util import relmse

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_cbox_xml_renders_like_the_programmatic_scene(gpu_ctx, tmp_path):
    out = tmp_path / "cbox.exr"
    with host_api.HostScene(ROOT / "assets" / "scenes" / "cbox.xml", params=dict(w=64, h=48, spp=8, depth=5)) as hs:
        st = hs.render(str(out))
        rd = hs.render_desc()
    assert st.paths == 64 * 48 * 8 and st.kernel_launches > 0
    img = host_api.read_exr_rgba(out)
    assert img.shape == (48, 64, 4) and np.isfinite(img).all()
    sd = scenes.cbox(64, 48)
    with capi.Scene(gpu_ctx, sd) as sc:
        film, _ = sc.render(capi.render_desc(spp=8, max_depth=5))
        rgba = sc.develop(film)
    # same meshes, spectra and seeds; sample_to_camera may differ in the last ulp (float32 inverse here, float64 there)
    assert relmse(img, rgba) < 1e-6
    np.testing.assert_allclose(img[..., 3], 1.0, rtol=1e-5)
    ofilm, _ = pyoracle.OracleScene(sd).render(capi.render_desc(spp=rd.spp, max_depth=rd.max_depth, rr_depth=rd.rr_depth))
    assert relmse(img, pyoracle.develop(ofilm)) < 1e-4


def test_command_line_renderer(tmp_path):
    import subprocess
    exe = ROOT / "misaki_render_b200" / "lib" / "misaki_b200"
    out = tmp_path / "cli.pfm"
    r = subprocess.run([str(exe), str(ROOT / "assets" / "scenes" / "cbox.xml"), "-D", "w=32", "-D", "h=32", "-D", "spp=4", "-D", "depth=3",
                        "-o", str(tmp_path / "cli.exr")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "cli.exr").exists() and "Rendering finished" in r.stderr
    img = host_api.read_exr_rgba(tmp_path / "cli.exr")
    assert img.shape == (32, 32, 4) and img[..., :3].max() > 0
