"""The AOV integrator (reference src/librender/integrators/aov.cpp:22-144): depth / position / uv / geometric and
shading normal of the primary hit plus the nested path tracer's RGBA, every channel splatted through the
reconstruction filter like XYZAW (integrator.cpp:103-126, hdrfilm.cpp:83-86 divides them by W on develop).

CPU part: known answers for the oracle's restatement.  GPU part: msk_gpu_render_aov vs the oracle."""
import numpy as np
import pytest

from misaki_render_b200 import capi
from oracle import pyoracle
from workloads import scenes
from tests.util import relmse

ALL = ["depth", "position", "uv", "geo_normal", "sh_normal"]


def _wall_scene(w=24, h=24):
    """A camera at the origin looking down +z at an axis-aligned wall z = 5 that covers the whole view."""
    from misaki_render_b200.scene import SceneDescription, lookat
    from workloads import meshes
    sd = SceneDescription(w, h, fov=40.0, near_clip=0.1, far_clip=100.0, to_world=lookat((0, 0, 0), (0, 0, 1), (0, 1, 0)))
    v, t = meshes.quad((-50, -50, 5), (50, -50, 5), (50, 50, 5), (-50, 50, 5))
    sd.add_mesh(v, t, sd.bsdf_diffuse((0.5, 0.5, 0.5)))
    return sd


def test_oracle_aov_known_answers():
    sd = _wall_scene()
    rd = capi.render_desc(spp=4, max_depth=3)
    film, st = pyoracle.OracleScene(sd).render_aov(rd, ALL)
    assert film.shape == (24, 24, 5 + 12)
    w = film[..., 4:5]
    assert (w > 0).all()
    a = film[..., 5:] / w  # HDRFilm::image: AOV channels / W
    depth, pos, uv, gn, sn = a[..., 0], a[..., 1:4], a[..., 4:6], a[..., 6:9], a[..., 9:12]
    np.testing.assert_allclose(pos[..., 2], 5.0, rtol=1e-5)            # every hit lies on the wall
    np.testing.assert_allclose(np.abs(gn[..., 2]), 1.0, rtol=1e-5)     # geometric normal = +-z
    np.testing.assert_allclose(gn[..., :2], 0.0, atol=1e-6)
    np.testing.assert_allclose(sn, gn, atol=1e-6)                      # no vertex normals: sh_frame.n = n (mesh.cpp:97-99)
    # depth is the distance along the (normalised) ray: |p| for a camera at the origin
    np.testing.assert_allclose(depth[12, 12], np.linalg.norm(pos[12, 12]), rtol=2e-3)
    assert depth.min() >= 5.0 - 1e-4 and depth[0, 0] > depth[12, 12]
    assert (uv >= -1e-6).all() and (uv <= 1 + 1e-6).all()              # no texcoords: si.uv = barycentrics (mesh.cpp:65)
    assert np.all(film[..., :3] == 0.0)                                # no nested integrator: zero radiance (stated deviation)
    assert st.rays_closest == 24 * 24 * 4 and st.rays_shadow == 0


def test_oracle_aov_nested_path_matches_path_integrator():
    sd = scenes.cbox(32, 32)
    rd = capi.render_desc(spp=4, max_depth=4)
    osc = pyoracle.OracleScene(sd)
    plain, _ = osc.render(rd)
    film, _ = osc.render_aov(rd, ["depth", "integrator"])
    assert film.shape[-1] == 5 + 1 + 4
    np.testing.assert_array_equal(film[..., :5], plain)  # the AOV integrator's extra ray_intersect draws no random numbers
    np.testing.assert_allclose(film[..., 9], film[..., 4], rtol=1e-6)  # A of the nested RGBA is 1 per sample
    assert film[..., 6:9].max() > 0


def test_oracle_aov_miss_reads_zero():
    from misaki_render_b200.scene import SceneDescription, lookat
    from workloads import meshes
    sd = SceneDescription(16, 16, fov=40.0, near_clip=0.1, far_clip=100.0, to_world=lookat((0, 0, 0), (0, 0, 1), (0, 1, 0)))
    v, t = meshes.quad((-50, -50, -5), (50, -50, -5), (50, 50, -5), (-50, 50, -5))  # behind the camera
    sd.add_mesh(v, t, sd.bsdf_diffuse((0.5, 0.5, 0.5)))
    film, _ = pyoracle.OracleScene(sd).render_aov(capi.render_desc(spp=2, max_depth=2), ALL)
    assert np.all(film[..., 5:] == 0.0) and (film[..., 4] > 0).all()


def test_aov_invalid_type_is_an_error():
    with pytest.raises(RuntimeError):
        pyoracle.OracleScene(_wall_scene(8, 8)).render_aov(capi.render_desc(spp=1), [17])


# ------------------------------------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
def test_gpu_aov_geometry_matches_oracle(gpu_ctx):
    """Interpolated normals + texcoords (bunny-class mesh), misses against the background."""
    sd = scenes.bunny(64, 64, n=12)
    rd = capi.render_desc(spp=4, max_depth=3)
    with capi.Scene(gpu_ctx, sd) as sc:
        film, stats = sc.render_aov(rd, ALL)
    ofilm, _ = pyoracle.OracleScene(sd).render_aov(rd, ALL)
    assert film.shape == ofilm.shape
    np.testing.assert_allclose(film[..., 4], ofilm[..., 4], rtol=1e-5)
    w = np.maximum(ofilm[..., 4:5], 1e-20)
    a, o = film[..., 5:] / w, ofilm[..., 5:] / w
    # silhouette pixels: a primary ray grazing an edge may hit on one side and miss on the other
    close = np.isclose(a, o, rtol=2e-4, atol=2e-4).all(axis=-1)
    assert close.mean() > 0.995, close.mean()
    assert np.all(film[..., :3] == 0.0)
    assert stats.rays_closest == 64 * 64 * 4 and stats.rays_shadow == 0


@pytest.mark.gpu
def test_gpu_aov_nested_path_matches_oracle_and_plain_render(gpu_ctx):
    sd = scenes.cbox(64, 64)
    rd = capi.render_desc(spp=8, max_depth=5)
    types = ["sh_normal", "integrator", "depth"]
    with capi.Scene(gpu_ctx, sd) as sc:
        plain, _ = sc.render(rd)
        film, stats = sc.render_aov(rd, types)
        small, _ = sc.render_aov(capi.render_desc(spp=8, max_depth=5, paths_per_batch=64 * 64 * 3), types)
        rgba = sc.develop(film)
    np.testing.assert_array_equal(film[..., :5], plain)           # same paths, same film
    np.testing.assert_allclose(small, film, rtol=2e-5, atol=1e-6)  # batch partition
    ofilm, _ = pyoracle.OracleScene(sd).render_aov(rd, types)
    assert relmse(rgba, pyoracle.develop(ofilm[..., :5])) < 1e-4
    w = np.maximum(ofilm[..., 4:5], 1e-20)
    np.testing.assert_allclose(film[..., 5:8] / w, ofilm[..., 5:8] / w, atol=2e-3)           # shading normal
    e = relmse(film[..., 8:11] / w, ofilm[..., 8:11] / w)                                     # nested RGB
    assert e < 1e-4, e
    np.testing.assert_allclose(film[..., 11], ofilm[..., 11], rtol=1e-5)                      # nested A
    np.testing.assert_allclose(film[..., 12] / w[..., 0], ofilm[..., 12] / w[..., 0], rtol=1e-3, atol=1e-2)  # depth
    assert stats.rays_shadow > 0


@pytest.mark.gpu
def test_gpu_aov_rejects_bad_descriptions(gpu_ctx):
    sd = scenes.cbox(16, 16)
    with capi.Scene(gpu_ctx, sd) as sc:
        with pytest.raises(capi.MskError):
            sc.render_aov(capi.render_desc(spp=1), [9])
        with pytest.raises(capi.MskError):
            sc.render_aov(capi.render_desc(spp=1), ["integrator", "integrator"])


# ------------------------------------------------------------------------------------------------ host plugin
def _cbox_aov_xml(w, h, spp, depth, aovs="dd:depth,nn:sh_normal", nested=True):
    from pathlib import Path
    import re
    root = Path(__file__).resolve().parent.parent
    text = (root / "assets" / "scenes" / "cbox.xml").read_text()
    for k, v in dict(w=w, h=h, spp=spp, depth=depth).items():
        text = text.replace(f"${k}", str(v))
    inner = f'<integrator type="path" name="img"><integer name="max_depth" value="{depth}"/></integrator>' if nested else ""
    block = f'<integrator type="aov"><string name="aovs" value="{aovs}"/>{inner}</integrator>'
    text, n = re.subn(r'<integrator type="path">.*?</integrator>', block, text, count=1, flags=re.S)
    assert n == 1
    return text, str(root / "assets" / "scenes")


def test_host_aov_plugin_is_registered_and_validates():
    from misaki_render_b200 import host_api
    assert "aov" in host_api.registered_plugins()
    text, base = _cbox_aov_xml(16, 16, 1, 2, aovs="x:nonsense")
    with pytest.raises(host_api.HostError, match="Invalid AOV type"):
        host_api.HostScene(xml=text, base_dir=base)
    text, base = _cbox_aov_xml(16, 16, 1, 2)
    with host_api.HostScene(xml=text, base_dir=base) as hs:  # loads without a GPU; rendering needs one
        assert hs.desc().nmeshes == 8


@pytest.mark.gpu
def test_gpu_host_aov_render_writes_named_channels(gpu_ctx, tmp_path):
    from misaki_render_b200 import host_api
    text, base = _cbox_aov_xml(48, 48, 4, 4)
    out = tmp_path / "aov.exr"
    with host_api.HostScene(xml=text, base_dir=base) as hs:
        st = hs.render(str(out))
    assert st.paths == 48 * 48 * 4
    names, img = host_api.read_exr_channels(out)
    # hdrfilm.cpp:52-59: R,G,B,A then the AOV channels; EXR stores them sorted by name
    assert sorted(names) == names
    assert set(names) == {"R", "G", "B", "A", "dd", "nn.X", "nn.Y", "nn.Z", "img.R", "img.G", "img.B", "img.A"}
    ch = {n: img[..., i] for i, n in enumerate(names)}
    sd = scenes.cbox(48, 48)
    with capi.Scene(gpu_ctx, sd) as sc:
        film, _ = sc.render_aov(capi.render_desc(spp=4, max_depth=4), ["depth", "sh_normal", "integrator"])
        rgba = sc.develop(film)
    w = np.maximum(film[..., 4], 1e-20)
    assert relmse(np.stack([ch["R"], ch["G"], ch["B"]], -1), rgba[..., :3]) < 1e-6
    np.testing.assert_allclose(ch["dd"], film[..., 5] / w, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(ch["nn.Y"], film[..., 7] / w, atol=1e-4)
    np.testing.assert_allclose(ch["img.A"], 1.0, rtol=1e-5)
    assert relmse(np.stack([ch["img.R"], ch["img.G"], ch["img.B"]], -1), film[..., 9:12] / w[..., None]) < 1e-6
