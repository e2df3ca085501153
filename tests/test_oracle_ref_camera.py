"""The camera pinned to the reference's own code.  tests/golden/ref_camera.json was produced by the reference's
include/misaki/core/transform.h, src/librender/sensor.cpp and src/librender/sensors/perspective.cpp compiled from where they
lie (oracle/ref_camera_wrap.cpp, tools/gen_golden_ref_camera.py).  Held to it here, on CPU: the oracle's camera_sample_ray
(what every GPU parity test compares k_raygen with, tests/test_gpu_intersect.py), the scene builder's camera matrices
(misaki_render_b200/scene.py) and the host front-end's PerspectiveCamera::describe + <lookat> / <translate> / <scale> /
<rotate> (host/plugins.cpp, host/xml.cpp).  1e-6 relative: the stand-in's 4x4 inverse is a cofactor inverse, Eigen's is a
packed SSE routine -- same mathematics, another order of float operations."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from misaki_render_b200 import host_api
from misaki_render_b200.scene import SceneDescription, lookat
from oracle import pyoracle

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = json.loads((ROOT / "tests" / "golden" / "ref_camera.json").read_text())
RTOL = 1e-6


def close(a, b, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=RTOL * s)


def _scene(c):
    o, t, u = c["lookat"]
    return SceneDescription(c["width"], c["height"], fov=c["fov"], near_clip=c["near_clip"], far_clip=c["far_clip"], to_world=lookat(o, t, u))


@pytest.mark.parametrize("name", sorted(GOLDEN["cameras"]))
def test_oracle_camera_rays_match_the_reference_camera(name):
    c = GOLDEN["cameras"][name]
    sd = _scene(c)
    sd.add_mesh(np.array([[0, 0, 0, 0, 0, 1, 0, 0], [1, 0, 0, 0, 0, 1, 0, 0], [0, 1, 0, 0, 0, 1, 0, 0]], np.float32), np.array([[0, 1, 2]], np.uint32),
                sd.bsdf_diffuse((0.5, 0.5, 0.5)))
    osc = pyoracle.OracleScene(sd)
    s = np.asarray(c["samples"], np.float32)
    rays = osc.camera_rays(np.stack([s[:, 1], s[:, 2], s[:, 0]], axis=1))  # the oracle takes (px, py, wavelength sample)
    osc.close()
    ref = np.asarray(c["rays"], np.float64)
    close(rays["o"], ref[:, 0:3])
    close(rays["d"], ref[:, 3:6], scale=1.0)
    close(rays["tmin"], ref[:, 6])
    close(rays["tmax"], ref[:, 7])


@pytest.mark.parametrize("name", sorted(GOLDEN["cameras"]))
def test_scene_builder_matrices_match_the_reference(name):
    c = GOLDEN["cameras"][name]
    cam = _scene(c).camera()
    s2c, ref = np.array(cam.sample_to_camera[:], np.float64).reshape(4, 4), np.array(c["sample_to_camera"])
    for i in range(4):  # row by row: the rows differ by orders of magnitude
        close(s2c[i], ref[i])
    close(np.array(cam.to_world[:]).reshape(4, 4), c["to_world"])


@pytest.mark.parametrize("name", sorted(GOLDEN["cameras"]))
def test_host_frontend_camera_matches_the_reference(name):
    c = GOLDEN["cameras"][name]
    o, t, u = (" ".join(repr(float(x)) for x in v) for v in c["lookat"])
    xml = f"""<scene><sensor type="perspective"><float name="fov" value="{c['fov']!r}"/><float name="near_clip" value="{c['near_clip']!r}"/>
      <float name="far_clip" value="{c['far_clip']!r}"/><transform name="to_world"><lookat origin="{o}" target="{t}" up="{u}"/></transform>
      <film type="hdrfilm"><integer name="width" value="{c['width']}"/><integer name="height" value="{c['height']}"/></film></sensor></scene>"""
    with host_api.HostScene(xml=xml) as hs:
        cam = hs.desc().camera
        s2c, ref = np.array(cam.sample_to_camera[:], np.float64).reshape(4, 4), np.array(c["sample_to_camera"])
        for i in range(4):
            close(s2c[i], ref[i])
        close(np.array(cam.to_world[:]).reshape(4, 4), c["to_world"])
        assert (cam.width, cam.height) == (c["width"], c["height"])
        close(cam.near_clip, c["near_clip"]); close(cam.far_clip, c["far_clip"])


@pytest.mark.parametrize("k", range(len(GOLDEN["transforms"])))
def test_host_frontend_transform_tags_match_the_reference(k):
    g = GOLDEN["transforms"][k]
    x, y, z = (repr(float(a)) for a in g["v"])
    tag = {"translate": f'<translate x="{x}" y="{y}" z="{z}"/>', "scale": f'<scale x="{x}" y="{y}" z="{z}"/>',
           "rotate": f'<rotate x="{x}" y="{y}" z="{z}" angle="{math.degrees(g["angle"])!r}"/>'}[g["kind"]]
    xml = f"""<scene><sensor type="perspective"><transform name="to_world">{tag}</transform>
      <film type="hdrfilm"><integer name="width" value="8"/><integer name="height" value="4"/></film></sensor></scene>"""
    with host_api.HostScene(xml=xml) as hs:
        close(np.array(hs.desc().camera.to_world[:]).reshape(4, 4), g["matrix"])
