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


def _write_obj(path, verts, tris, normals=False):
    with open(path, "w") as f:
        for v in verts:
            f.write("v %.9g %.9g %.9g\n" % tuple(v[:3]))
        if normals:
            for v in verts:
                f.write("vn %.9g %.9g %.9g\n" % tuple(v[3:6]))
        for t in tris:
            f.write(("f %d//%d %d//%d %d//%d\n" % tuple(np.repeat(t + 1, 2))) if normals else ("f %d %d %d\n" % tuple(t + 1)))


def test_volpath_scene_file_through_the_host_frontend(gpu_ctx, tmp_path):
    """A scene file with the structure of the reference's assets/teapot-full/scene.xml -- `volpath`, a `twosided`
    diffuse floor with a `checkerboard` reflectance, a `dielectric` shell with an interior `homogeneous` medium, a
    `<boolean>` property, a `constant` emitter plus an area light -- rendered by the host front-end (XML -> plugins ->
    GpuVolPathIntegrator::render -> C ABI -> Film::develop) and compared with the oracle's render of the SAME flattened
    description."""
    from workloads import meshes
    v, t = meshes.cube_sphere(10, seed=7, octaves=2, amplitude=0.05, radius=0.8, center=(0, 0.9, 0), normals=True)
    _write_obj(tmp_path / "blob.obj", v, t, normals=True)
    (tmp_path / "floor.obj").write_text("v -4 0 -4\nv -4 0 4\nv 4 0 4\nv 4 0 -4\nvt 0 0\nvt 0 1\nvt 1 1\nvt 1 0\nf 1/1 2/2 3/3 4/4\n")
    (tmp_path / "light.obj").write_text("v -1 4 -1\nv 1 4 -1\nv 1 4 1\nv -1 4 1\nf 1 2 3 4\n")
    xml = """<scene>
      <integrator type="volpath"><integer name="max_depth" value="12"/><boolean name="hide_emitters" value="false"/></integrator>
      <sensor type="perspective"><float name="fov" value="38"/>
        <transform name="to_world"><lookat origin="0 2.4 -4.6" target="0 0.8 0" up="0 1 0"/></transform>
        <sampler type="independent"><integer name="sample_count" value="8"/></sampler>
        <film type="hdrfilm"><integer name="width" value="64"/><integer name="height" value="48"/></film></sensor>
      <bsdf type="twosided" id="Floor"><bsdf type="diffuse"><texture name="reflectance" type="checkerboard">
        <rgb name="color0" value="0.7, 0.7, 0.65"/><rgb name="color1" value="0.3, 0.3, 0.25"/>
        <transform name="to_uv"><scale x="8" y="8"/></transform></texture></bsdf></bsdf>
      <shape type="obj"><string name="filename" value="floor.obj"/><ref id="Floor"/></shape>
      <shape type="obj"><string name="filename" value="blob.obj"/><boolean name="faceNormals" value="true"/>
        <bsdf type="dielectric"><float name="int_ior" value="1.33"/><float name="ext_ior" value="1"/></bsdf>
        <medium type="homogeneous" name="interior"><rgb name="sigma_s" value="1.5, 1.2, 1.0"/><rgb name="sigma_a" value="0.15, 0.32, 0.74"/></medium></shape>
      <shape type="obj"><string name="filename" value="light.obj"/><emitter type="area"><rgb name="radiance" value="18, 18, 18"/></emitter></shape>
      <emitter type="constant"><rgb name="radiance" value="0.3, 0.35, 0.5"/></emitter>
    </scene>"""
    (tmp_path / "scene.xml").write_text(xml)
    out = tmp_path / "vol.exr"
    with host_api.HostScene(tmp_path / "scene.xml") as hs:
        st = hs.render(str(out))
        rd, desc = hs.render_desc(), hs.desc()
        assert rd.integrator == capi.INTEGRATOR_VOLPATH and desc.nmedia == 1 and desc.environment >= 0

        class _Flat:  # the oracle consumes the very description the integrator plugin handed to the C ABI
            width, height = desc.camera.width, desc.camera.height

            @staticmethod
            def c_desc():
                return desc
        ofilm, ost = pyoracle.OracleScene(_Flat).render(rd)
    img = host_api.read_exr_rgba(out)
    assert img.shape == (48, 64, 4) and np.isfinite(img).all() and st.paths == 64 * 48 * 8
    assert st.rays_closest <= ost.rays_closest  # Russian roulette before the trace (see test_gpu_render)
    e = relmse(img, pyoracle.develop(ofilm))
    print(f"[host volpath] relMSE={e:.3e}")
    assert e < 1e-3
