"""Known-answer tests pinning the oracle's building blocks (SURVEY.md section 8c list).  The reference has
no tests of its own; these check the restatement against published vectors and closed forms."""
import math

import numpy as np

from misaki_render_b200 import capi
from misaki_render_b200 import scene as mscene
from oracle import pyoracle as po
from workloads import meshes, scenes


def test_pcg32_published_vector():
    # pcg32-demo: pcg32_srandom(42, 54) -> first six outputs (O'Neill, pcg-c-basic)
    want = [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]
    assert list(po.pcg32_uints(42, 54, 6)) == want


def test_sampler_floats_follow_pcg32():
    """IndependentSampler::seed(s): rng.seed(s + base_seed, PCG32_DEFAULT_STREAM); next_float32 = bits trick."""
    stream = 0xda3e39cb94b95bdb
    for seed, base in [(0, 0), (12345, 0), (7, 1000)]:
        u = po.pcg32_uints(seed + base, stream, 8)
        want = ((u >> 9) | 0x3f800000).astype(np.uint32).view(np.float32) - np.float32(1)
        np.testing.assert_array_equal(po.pcg32_floats(seed, 8, base), want)
    f = po.pcg32_floats(99, 10000)
    assert f.min() >= 0 and f.max() < 1 and abs(f.mean() - 0.5) < 0.02


def test_sample_wavelength():
    wl, w = po.sample_wavelength(0.0)
    assert 359.9 < wl[0] < 360.5  # u = 0 maps to the lower end of the visible range
    for u in np.linspace(0, 0.999, 50):
        wl, w = po.sample_wavelength(float(u))
        assert (wl >= 359.9).all() and (wl <= 830.1).all()
        # pdf of the "rgb" wavelength sampler is sech^2: weight = 1/pdf = 253.82 cosh^2(0.0072 (l - 538))
        np.testing.assert_allclose(w, 253.82 * np.cosh(0.0072 * (wl.astype(np.float64) - 538)) ** 2, rtol=1e-5)
    # shifted samples: u_i = frac(u + i/4)
    wl_a, _ = po.sample_wavelength(0.1)
    wl_b, _ = po.sample_wavelength(0.35)
    np.testing.assert_allclose(wl_a[1], wl_b[0], rtol=1e-6)


def test_warps():
    rng = np.random.default_rng(0)
    for _ in range(200):
        u, v = rng.random(2)
        b = po.warp(0, u, v)
        assert b[0] >= 0 and b[1] >= 0 and b[0] + b[1] <= 1 + 1e-6
        d = po.warp(1, u, v)
        assert d[0] ** 2 + d[1] ** 2 <= 1 + 1e-6
        h = po.warp(2, u, v)
        np.testing.assert_allclose(np.linalg.norm(h), 1, atol=1e-6)
        assert h[2] >= 0
        s = po.warp(3, u, v)
        np.testing.assert_allclose(np.linalg.norm(s), 1, atol=1e-6)
    np.testing.assert_array_equal(po.warp(1, 0.5, 0.5)[:2], [0, 0])
    # cosine hemisphere: E[z] = 2/3
    zs = [po.warp(2, *rng.random(2))[2] for _ in range(4000)]
    assert abs(np.mean(zs) - 2 / 3) < 0.02


def test_fresnel():
    np.testing.assert_allclose(po.fresnel(1.0, 1.5)[0], 0.04, rtol=1e-6)
    F, ct, eta_it, eta_ti = po.fresnel(1.0, 1.5)
    assert ct == -1.0 and eta_it == np.float32(1.5) and eta_ti == np.float32(1 / 1.5)
    assert po.fresnel(0.0, 1.5)[0] == 1.0
    assert po.fresnel(0.3, 1.0)[0] == 0.0
    assert po.fresnel(-0.2, 1.5)[0] == 1.0  # total internal reflection from inside
    np.testing.assert_allclose(po.fresnel(-1.0, 1.5)[0], 0.04, rtol=1e-6)
    # conductor: eta = 0, k = 1 is a perfect mirror (conductor.cpp:14-17 "initially set up for mirror")
    np.testing.assert_allclose(po.fresnel_conductor(0.7, [0] * 4, [1] * 4), 1.0, rtol=1e-6)
    # k = 0 reduces to the dielectric formula
    np.testing.assert_allclose(po.fresnel_conductor(0.8, [1.5] * 4, [0] * 4), po.fresnel(0.8, 1.5)[0], rtol=1e-5)


def test_ggx_normalisation_and_sampling():
    for au, av in [(0.1, 0.1), (0.3, 0.3), (0.5, 0.2)]:
        th = np.linspace(0, np.pi / 2, 1500)[1:-1]
        ph = np.linspace(0, 2 * np.pi, 720, endpoint=False)
        tot = 0.0
        for p in ph[::4]:
            vals = np.array([po.ggx(0, au, av, (math.sin(t) * math.cos(p), math.sin(t) * math.sin(p), math.cos(t)))[0] for t in th[::3]])
            tot += np.sum(vals * np.cos(th[::3]) * np.sin(th[::3])) * (th[3] - th[0]) * (ph[4] - ph[0])
        assert abs(tot - 1) < 0.03, (au, av, tot)  # integral of D(m) cos(theta_m) over the hemisphere = 1
        rng = np.random.default_rng(1)
        for _ in range(100):
            r = po.ggx(1, au, av, (0, 0, 1), (rng.random(), rng.random(), 0))
            m, pdf = r[:3], r[3]
            np.testing.assert_allclose(np.linalg.norm(m), 1, atol=1e-5)
            D = po.ggx(0, au, av, m)[0]
            if pdf > 0 and D > 0:
                np.testing.assert_allclose(pdf, D * m[2], rtol=2e-3)
    assert po.ggx(2, 0.1, 0.1, (0, 0, 1), (0, 0, 1))[0] == 1.0
    assert po.ggx(2, 0.1, 0.1, (0.6, 0, -0.8), (0.99995, 0, 0.01))[0] == 0.0  # v.m > 0 but v below the horizon


def test_gaussian_filter_table():
    r, t = po.gaussian_filter(0.5)
    assert r == 2.0 and t[32] == 0 and (np.diff(t[:32]) <= 0).all()
    np.testing.assert_allclose(t[:32].sum() * 2 * r / 32, 1.0, rtol=1e-6)
    r2, t2 = mscene.gaussian_filter(0.5)  # the product's host-side table
    assert r2 == r
    np.testing.assert_allclose(t2, t, rtol=2e-7, atol=1e-12)


def test_d65_white_has_unit_luminance():
    """XYZ of D65 * (1/10568) through the wavelength sampler integrates to Y ~= 1 (d65.cpp:33)."""
    sd = mscene.SceneDescription(4, 4)
    sid = sd.spectrum_d65(1.0)
    osc = po.OracleScene(sd)
    acc = np.zeros(3)
    n = 4000
    for u in (np.arange(n) + 0.5) / n:
        wl, w = po.sample_wavelength(float(u))
        acc += po.spectrum_to_xyz(osc.spectrum_eval(sid, wl) * w, wl)
    xyz = acc / n
    assert abs(xyz[1] - 1.0) < 0.01
    assert abs(xyz[0] - 0.9505) < 0.01 and abs(xyz[2] - 1.089) < 0.012  # D65 white point


def test_uniform_spectrum_all_or_nothing():
    sd = mscene.SceneDescription(4, 4)
    sid = sd.spectrum_uniform(0.7)
    osc = po.OracleScene(sd)
    np.testing.assert_array_equal(osc.spectrum_eval(sid, [400, 500, 600, 700]), np.float32(0.7))
    np.testing.assert_array_equal(osc.spectrum_eval(sid, [400, 500, 600, 900]), 0)  # uniform.cpp:20-25: .all()


def _bsdf_scene():
    sd = mscene.SceneDescription(4, 4)
    ids = dict(diffuse=sd.bsdf_diffuse((0.8, 0.8, 0.8)),
               rc=sd.bsdf_roughconductor(eta=(0.200438, 0.924033, 1.10221), k=(3.91295, 2.45285, 2.14219), alpha=0.3,
                                         specular_reflectance=sd.spectrum_uniform(1.0)),
               rd=sd.bsdf_roughdielectric(int_ior=1.5, ext_ior=1.0, alpha=0.3, specular_reflectance=sd.spectrum_uniform(1.0),
                                          specular_transmittance=sd.spectrum_uniform(1.0)),
               cond=sd.bsdf_conductor(eta=(0.2, 0.9, 1.1), k=(3.9, 2.4, 2.1)), diel=sd.bsdf_dielectric(1.5, 1.0),
               two=sd.bsdf_diffuse((0.5, 0.5, 0.5), twosided=True))
    return sd, ids


def test_bsdf_sample_consistent_with_eval_and_pdf():
    """For the smooth BSDFs weight * pdf == eval (sample() returns f cos / pdf), except the reference's own
    inconsistencies: roughconductor sample() omits specular_reflectance and roughdielectric sample() omits
    specular_transmittance while eval() applies them (roughconductor.cpp:79 vs :99) -- so the test uses a
    uniform reflectance of exactly 1 (an <rgb> white upsamples to ~0.96..1, not 1)."""
    sd, ids = _bsdf_scene()
    osc = po.OracleScene(sd)
    rng = np.random.default_rng(3)
    wl = [450, 520, 600, 680]
    for name in ("diffuse", "rc", "rd"):
        checked = 0
        for _ in range(300):
            ct = rng.uniform(0.2, 1.0) * (1 if name != "rd" or rng.random() < 0.5 else -1)
            st = math.sqrt(1 - ct * ct)
            phi = rng.uniform(0, 2 * np.pi)
            wi = (st * math.cos(phi), st * math.sin(phi), ct)
            r = osc.bsdf(ids[name], wi, wl, rng.random(3), (0, 0, 1))
            if r["pdf"] <= 0 or not r["weight"].any():
                continue
            r2 = osc.bsdf(ids[name], wi, wl, (0, 0, 0), r["wo"])
            if r2["eval_pdf"] <= 0:
                continue
            np.testing.assert_allclose(r2["eval_pdf"], r["pdf"], rtol=5e-3)
            if name != "rd":
                np.testing.assert_allclose(r["weight"] * r["pdf"], r2["eval"], rtol=5e-3, atol=1e-7)
            else:
                # reference quirk (q9): roughdielectric samples m from the alpha-SCALED distribution
                # (roughdielectric.cpp:70-75) but its weight assumes the unscaled one (:109-110), so
                # weight * pdf == eval * D_scaled(m) / D(m); only the pdf identity holds.
                assert np.isfinite(r["weight"]).all() and (r["weight"] >= 0).all()
            checked += 1
        assert checked > 100, name


def test_bsdf_energy_and_delta_lobes():
    sd, ids = _bsdf_scene()
    osc = po.OracleScene(sd)
    rng = np.random.default_rng(4)
    wl = [450, 520, 600, 680]
    wi = (0.3, 0.1, math.sqrt(1 - 0.1))
    for name in ("diffuse", "rc", "rd", "cond", "diel"):
        ws = np.array([osc.bsdf(ids[name], wi, wl, rng.random(3), (0, 0, 1))["weight"] for _ in range(2000)])
        assert ws.mean(axis=0).max() <= 1.02, name  # white furnace: albedo <= 1
    r = osc.bsdf(ids["cond"], wi, wl, (0.1, 0.2, 0.3), (0, 0, 1))
    np.testing.assert_allclose(r["wo"], (-wi[0], -wi[1], wi[2]), rtol=1e-6)
    assert r["pdf"] == 1 and r["type"] == 0x20 and not r["eval"].any()
    # one-sided diffuse sees nothing from below; the two-sided adapter mirrors it
    assert osc.bsdf(ids["diffuse"], (0, 0, -1), wl, (0.1, 0.2, 0.3), (0, 0, 1))["pdf"] == 0
    r = osc.bsdf(ids["two"], (0, 0, -1), wl, (0.1, 0.2, 0.3), (0, 0, -1))
    assert r["pdf"] > 0 and r["wo"][2] < 0 and r["eval"].all()
    # smooth glass at normal incidence reflects 4 %
    refl = np.mean([osc.bsdf(ids["diel"], (0, 0, 1), wl, (0, u, 0), (0, 0, 1))["type"] == 0x20 for u in rng.random(5000)])
    assert abs(refl - 0.04) < 0.01


def test_bvh_equals_brute_force():
    sd = scenes.bunny(32, 32, n=12)
    osc = po.OracleScene(sd)
    rng = np.random.default_rng(5)
    rays = np.zeros(3000, dtype=capi.RAY_DTYPE)
    rays["o"] = rng.uniform(-3, 3, (3000, 3)) + (0, 1.5, 0)
    d = rng.normal(size=(3000, 3))
    rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmin"], rays["tmax"] = 1e-4, np.inf
    a, b = osc.intersect(rays), osc.intersect(rays, brute_force=True)
    for k in ("t", "u", "v", "prim", "geom"):
        np.testing.assert_array_equal(a[k], b[k])
    assert np.isfinite(a["t"]).mean() > 0.3
    np.testing.assert_array_equal(osc.occluded(rays) != 0, np.isfinite(b["t"]))
    # barycentric convention: o + t d == (1-u-v) p0 + u p1 + v p2
    hit = np.isfinite(b["t"])
    for i in np.nonzero(hit)[0][:200]:
        m = sd.meshes[b["geom"][i]]
        f = m["tris"][b["prim"][i]]
        p = m["verts"][f, :3].astype(np.float64)
        q = (1 - b["u"][i] - b["v"][i]) * p[0] + b["u"][i] * p[1] + b["v"][i] * p[2]
        np.testing.assert_allclose(q, rays["o"][i].astype(np.float64) + b["t"][i] * rays["d"][i], atol=2e-4)


def test_render_is_deterministic_and_thread_independent():
    sd = scenes.cbox(40, 24)  # ragged: not a multiple of the 32-pixel block
    rd = capi.render_desc(spp=3, max_depth=4)
    osc = po.OracleScene(sd)
    a, sa = osc.render(rd, nthreads=1)
    b, sb = osc.render(rd, nthreads=5)
    np.testing.assert_array_equal(a, b)
    assert sa.paths == 40 * 24 * 3 and sa.rays_closest == sb.rays_closest
    # the W channel is the sum of filter weights; the filter is normalised, so it averages spp per pixel
    w = a[4:-4, 4:-4, 4]
    assert abs(w.mean() - 3.0) < 0.1
    assert np.isfinite(a).all() and (a[..., 4] > 0).all()


def test_sample_range_accumulates():
    sd = scenes.cbox(16, 16)
    osc = po.OracleScene(sd)
    whole, _ = osc.render(capi.render_desc(spp=4, max_depth=3))
    part, _ = osc.render(capi.render_desc(spp=4, max_depth=3, sample_end=2))
    part, _ = osc.render(capi.render_desc(spp=4, max_depth=3, sample_begin=2, clear_film=False), film=part)
    np.testing.assert_allclose(part, whole, rtol=1e-5, atol=1e-7)


def test_direct_light_closed_form():
    """One emitting quad above a diffuse floor, max_depth 2: radiance leaving the floor towards the camera is
    rho/pi * E with E the irradiance of a parallel square light directly overhead (closed form for a point on
    the axis: E = L * 4 * atan-form; here checked against numerical quadrature)."""
    sd = mscene.SceneDescription(9, 9, fov=2.0, near_clip=0.1, far_clip=100,
                                 to_world=mscene.lookat((0, 3, 0), (0, 0, 0), (0, 0, 1)))
    fv, ft = meshes.quad((-50, 0, -50), (-50, 0, 50), (50, 0, 50), (50, 0, -50))
    sd.add_mesh(fv, ft, sd.bsdf_diffuse(sd.spectrum_uniform(0.5)))
    h, a = 2.0, 0.5
    lv, lt = meshes.quad((-a, h, -a), (a, h, -a), (a, h, a), (-a, h, a))  # faces down; camera is above it
    # shift the light sideways so it does not block the camera ray
    lv[:, 0] += 1.5
    sd.add_mesh(lv, lt, sd.bsdf_diffuse(sd.spectrum_uniform(0.0)), radiance=sd.spectrum_uniform(3.0))
    osc = po.OracleScene(sd)
    film, _ = osc.render(capi.render_desc(spp=4096, max_depth=2))
    # uniform spectra => result(lambda) = const = rho/pi * L * G; XYZ.y = mean(ybar * weight) * const
    xs = np.linspace(1.5 - a, 1.5 + a, 400)
    zs = np.linspace(-a, a, 400)
    X, Z = np.meshgrid(xs, zs)
    r2 = X ** 2 + Z ** 2 + h ** 2
    G = np.sum((h * h) / (r2 * r2)) * (xs[1] - xs[0]) * (zs[1] - zs[0])
    want = 0.5 / np.pi * 3.0 * G
    # luminance of a unit constant spectrum through the sampler
    n = 2000
    ybar = np.mean([po.spectrum_to_xyz(po.sample_wavelength(float(u))[1], po.sample_wavelength(float(u))[0])[1] for u in (np.arange(n) + 0.5) / n])
    got = film[4, 4, 1] / film[4, 4, 4]
    assert abs(got - want * ybar) < 0.03 * want * ybar, (got, want * ybar)


def test_checkerboard_texture_selects_children_by_uv():
    """textures/checkerboard.cpp:25-33: uv' = Transform3f(to_uv.extract()) * (u, v, 1); frac; color0 iff (u' > .5) == (v' > .5).
    extract() keeps the top-left 3x3 of the 4x4 (transform.h:142-148): the z COLUMN is the uv offset, the translation
    column is dropped."""
    from misaki_render_b200.scene import SceneDescription, scale, translate
    sd = SceneDescription(8, 8)
    a, b, c = sd.spectrum_uniform(0.25), sd.spectrum_uniform(0.5), sd.spectrum_uniform(1.0)
    plain = sd.spectrum_checkerboard(a, b)
    scaled = sd.spectrum_checkerboard(a, b, to_uv=scale((4, 2, 1)))
    moved = sd.spectrum_checkerboard(a, b, to_uv=translate((0.5, 0.0, 0.0)))  # translation is NOT applied
    m = np.eye(4, dtype=np.float32); m[0, 2] = 0.5
    zcol = sd.spectrum_checkerboard(a, b, to_uv=m)                              # the z column is
    nested = sd.spectrum_checkerboard(c, scaled)
    v, t = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32)
    sd.add_mesh(v, t, sd.bsdf_diffuse(nested))
    osc = po.OracleScene(sd)
    wl = np.array([400, 500, 600, 700], np.float32)
    ev = lambda sid, u, v: float(osc.texture_eval(sid, u, v, wl)[0])
    assert [ev(plain, *uv) for uv in [(0.25, 0.25), (0.75, 0.75), (0.75, 0.25), (0.25, 0.75)]] == [0.25, 0.25, 0.5, 0.5]
    assert ev(plain, 1.25, -0.75) == 0.25 and ev(plain, -0.25, 0.25) == 0.5          # frac() of negative coordinates
    assert ev(plain, 0.5, 0.5) == 0.25 and ev(plain, 0.5, 0.75) == 0.5               # strict > .5
    assert [ev(scaled, u, 0.1) for u in (0.05, 0.2, 0.3, 0.45)] == [0.25, 0.5, 0.25, 0.5]
    assert ev(moved, 0.25, 0.25) == ev(plain, 0.25, 0.25) and ev(zcol, 0.25, 0.25) == ev(plain, 0.75, 0.25)
    assert ev(nested, 0.25, 0.25) == 1.0 and ev(nested, 0.8, 0.1) == ev(scaled, 0.8, 0.1) == 0.25
    assert ev(nested, 0.1, 0.6) == ev(scaled, 0.1, 0.6) == 0.25 and ev(nested, 0.2, 0.6) == 0.5


def test_checkerboard_radiance_seen_directly():
    """An area light filling the view, max_depth 1: every sample returns Le = radiance->eval(si) (path.cpp:44-47,
    area.cpp:51-54) with si.uv = the hit barycentrics for a mesh without texcoords (mesh.cpp:65).  With uniform
    children 1 and 3 every pixel away from a cell border is one of two values in ratio 3."""
    from misaki_render_b200.scene import SceneDescription, lookat, scale
    from workloads import meshes
    sd = SceneDescription(32, 32, fov=30.0, to_world=lookat((0, 0, -2), (0, 0, 0), (0, 1, 0)))
    tex = sd.spectrum_checkerboard(sd.spectrum_uniform(1.0), sd.spectrum_uniform(3.0), to_uv=scale((2, 2, 1)))
    lv, lt = meshes.quad((-4, -4, 0), (-4, 4, 0), (4, 4, 0), (4, -4, 0))  # normal towards -z (the camera)
    sd.add_mesh(lv, lt, sd.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=tex)
    film, _ = po.OracleScene(sd).render(capi.render_desc(spp=16, max_depth=1))
    flat = SceneDescription(32, 32, fov=30.0, to_world=lookat((0, 0, -2), (0, 0, 0), (0, 1, 0)))
    flat.add_mesh(lv, lt, flat.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=flat.spectrum_uniform(1.0))
    base, _ = po.OracleScene(flat).render(capi.render_desc(spp=16, max_depth=1))  # same seeds, radiance 1 everywhere
    r = film[..., 1] / base[..., 1]
    ones, threes = np.abs(r - 1) < 1e-5, np.abs(r - 3) < 1e-5
    assert (ones | threes).mean() > 0.6 and ones.sum() > 50 and threes.sum() > 50
    assert r.min() >= 1 - 1e-5 and r.max() <= 3 + 1e-5


# ---- volumetric path tracer (integrators/volpath.cpp, media/homogeneous.cpp) -------------------
def _lit_wall(sigma_a, sigma_s, width=16):
    from misaki_render_b200.scene import SceneDescription, lookat
    sd = SceneDescription(width, width, fov=20.0, to_world=lookat((0, 0, -2), (0, 0, 0), (0, 1, 0)))
    lv, lt = meshes.quad((-4, -4, 0), (-4, 4, 0), (4, 4, 0), (4, -4, 0))  # emitting wall at distance 2, facing the camera
    sd.add_mesh(lv, lt, sd.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=sd.spectrum_uniform(1.0))
    sd.sensor_medium = sd.add_medium(sigma_a=sigma_a, sigma_s=sigma_s)
    return sd


def test_volpath_beer_lambert_in_an_absorbing_medium():
    """Camera inside a purely absorbing homogeneous medium looking at an emitter at distance d: free-flight
    sampling (homogeneous.cpp:21-53) either 'scatters' with sigma_s = 0 (throughput 0) or reaches the surface with
    weight transmittance / pdf = 1, so E[L] = Le exp(-sigma_a d).  d in [2, 2 / cos(half diagonal fov)]."""
    rd = capi.render_desc(spp=256, max_depth=1, integrator="volpath")
    clear, _ = po.OracleScene(_lit_wall(0.0, 0.0)).render(rd)
    y0 = (clear[..., 1] / clear[..., 4]).mean()
    assert y0 > 0
    for sigma in (0.25, 0.5, 1.0):
        f, _ = po.OracleScene(_lit_wall(float(sigma), 0.0)).render(rd)
        ratio = (f[..., 1] / f[..., 4]).mean() / y0
        lo, hi = math.exp(-sigma * 2.0 / math.cos(math.radians(14.2))), math.exp(-sigma * 2.0)
        assert lo * 0.97 <= ratio <= hi * 1.03, (sigma, ratio, lo, hi)
    # a medium without extinction is transparent (and does not produce NaNs: exp(0 * -inf) guard)
    f, _ = po.OracleScene(_lit_wall(0.0, 0.0)).render(capi.render_desc(spp=4, max_depth=3, integrator="volpath"))
    assert np.isfinite(f).all()


def test_volpath_without_media_agrees_with_the_path_tracer():
    """Both estimators are unbiased for the same transport (volpath.cpp adds NEE without MIS and counts emitter hits
    only on camera / delta chains; path.cpp MIS-weights BSDF-sampled hits), so their high-spp images agree."""
    sd = scenes.cbox(32, 32)
    osc = po.OracleScene(sd)
    a, _ = osc.render(capi.render_desc(spp=384, max_depth=6, integrator="volpath"))
    b, _ = osc.render(capi.render_desc(spp=384, max_depth=6))
    A, B = po.develop(a)[..., :3], po.develop(b)[..., :3]
    assert abs(A.mean() - B.mean()) < 0.02 * B.mean()
    assert np.abs(A - B).mean() < 0.03 * B.mean()


def test_volpath_scattering_medium_conserves_energy_between_emitting_walls():
    """White-furnace style check of the medium code: inside a closed box whose walls all emit radiance 1 and
    reflect nothing, a non-absorbing scattering medium (albedo 1) leaves the radiance at 1 for every density: each
    scattering event multiplies by sigma_s T / pdf and continues, each surface hit adds the wall's emission -- but
    volpath.cpp adds NEE at every scattering event ON TOP of emitter hits (emitted_radiance stays set), so the
    estimate is 1 + (expected number of scattering events) * 1.  The test pins that kept quirk: the excess over 1
    grows with sigma_s and vanishes without a medium."""
    from misaki_render_b200.scene import SceneDescription, lookat
    def box(sigma_s):
        sd = SceneDescription(8, 8, fov=40.0, to_world=lookat((0, 0, -0.5), (0, 0, 1), (0, 1, 0)))
        q = lambda a, b, c, d: meshes.quad(d, c, b, a)  # inward-facing normals
        walls = [q((-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)), q((-1, -1, -1), (-1, 1, -1), (1, 1, -1), (1, -1, -1)),
                 q((-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1)), q((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1)),
                 q((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1)), q((-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1))]
        for v, t in walls:
            sd.add_mesh(v, t, sd.bsdf_diffuse(0.0), radiance=sd.spectrum_uniform(1.0))
        if sigma_s is not None:
            sd.sensor_medium = sd.add_medium(sigma_a=0.0, sigma_s=float(sigma_s))
        return sd
    rd = capi.render_desc(spp=256, max_depth=-1, rr_depth=1000, integrator="volpath")
    vals = []
    for s in (None, 0.5, 2.0):
        f, _ = po.OracleScene(box(s)).render(rd)
        vals.append(float((f[..., 1] / f[..., 4]).mean()))
    base = vals[0]
    assert base > 0 and vals[1] > 1.1 * base and vals[2] > 1.5 * vals[1] - 0.5 * base
