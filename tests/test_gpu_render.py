"""GPU wavefront render vs the oracle's CPU render of the same scene with the same per-(pixel, sample)
seeds.  Float parity: transcendental functions and FMA contraction differ between glibc and CUDA, so the
bound is relMSE <= 1e-6 on the developed linear-sRGB image; scenes with dielectric chains use the chaotic bound below."""
import numpy as np
import pytest

from misaki_render_b200 import capi
from oracle import pyoracle
from workloads import scenes
from tests.util import relmse

pytestmark = pytest.mark.gpu

EQUAL_SEED_RELMSE = 1e-6  # float reassociation only (measured 1e-12 .. 2e-8)
# Refraction chains through a curved dielectric amplify ulp-level differences of the hit point (watertight vs
# Moeller-Trumbore barycentrics, CUDA vs glibc transcendentals) until a handful of the 32 768 paths take another
# branch; at 8 spp one such path moves relMSE by ~1e-5.  The bound for those scenes is 1e-3 plus a cap on the
# fraction of pixels that differ visibly.
CHAOTIC_RELMSE, CHAOTIC_BAD_PIXELS = 1e-3, 0.03  # (measured 0.0195 with IEEE division / sqrt in the shade stage, 0.0225 with the <= 2 ulp ones of the product build)


def bad_pixel_fraction(img, ref):
    img, ref = np.asarray(img, np.float64)[..., :3], np.asarray(ref, np.float64)[..., :3]
    return float(np.mean(np.any(np.abs(img - ref) > 0.02 * (np.abs(ref) + 0.1), axis=-1)))


def _both(gpu_ctx, sd, rd):
    with capi.Scene(gpu_ctx, sd) as sc:
        film, stats = sc.render(rd)
        rgba = sc.develop(film)
    ofilm, ost = pyoracle.OracleScene(sd).render(rd)
    return film, rgba, ofilm, pyoracle.develop(ofilm), stats, ost


def test_cbox_equal_seed(gpu_ctx):
    sd = scenes.cbox(64, 64)
    rd = capi.render_desc(spp=16, max_depth=5)
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    assert np.isfinite(film).all()
    np.testing.assert_allclose(film[..., 4], ofilm[..., 4], rtol=1e-5)  # filter weights: same positions
    e = relmse(rgba, oref)
    assert e < EQUAL_SEED_RELMSE, e
    assert stats.paths == 64 * 64 * 16
    # same paths => same ray counts up to rare branch flips
    assert abs(int(stats.rays_closest) - int(ost.rays_closest)) <= 1e-3 * ost.rays_closest
    # the GPU path does not trace shadow rays whose NEE contribution is exactly zero (BSDF value 0 for a light
    # direction below the surface): an output-equivalent skip, so its any-hit count is a subset of the oracle's
    assert 0.7 * ost.rays_shadow <= stats.rays_shadow <= ost.rays_shadow


def test_cbox_unbounded_depth_rr(gpu_ctx):
    sd = scenes.cbox(48, 48)
    rd = capi.render_desc(spp=8, max_depth=-1, rr_depth=3)
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    assert relmse(rgba, oref) < EQUAL_SEED_RELMSE
    assert stats.bounces > 5


def test_sample_range_partition_matches_whole(gpu_ctx):
    """Any partition of the sample range reproduces the whole job (multi-GPU sharding rests on this)."""
    sd = scenes.cbox(48, 48)
    with capi.Scene(gpu_ctx, sd) as sc:
        whole, _ = sc.render(capi.render_desc(spp=8, max_depth=4))
        part, _ = sc.render(capi.render_desc(spp=8, max_depth=4, sample_begin=0, sample_end=3))
        part, _ = sc.render(capi.render_desc(spp=8, max_depth=4, sample_begin=3, sample_end=8, clear_film=False), film=part)
        small_batches, _ = sc.render(capi.render_desc(spp=8, max_depth=4, paths_per_batch=48 * 48 * 2))
        again, _ = sc.render(capi.render_desc(spp=8, max_depth=4))
    np.testing.assert_array_equal(whole, again)  # deterministic film: gather, no float atomics
    np.testing.assert_allclose(part, whole, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(small_batches, whole, rtol=2e-5, atol=1e-6)


def _report(name, e, stats, ost):
    print(f"[{name}] relMSE={e:.3e} gpu rays c/s={stats.rays_closest}/{stats.rays_shadow} oracle={ost.rays_closest}/{ost.rays_shadow} "
          f"bounces={stats.bounces}")


def test_roughconductor_equal_seed(gpu_ctx):
    """C2's material (roughconductor.cpp:53-120, GGX, unbounded eta/k spectra) on a small bunny-class mesh with
    interpolated shading normals; unbounded depth with Russian roulette (path.cpp:112-120)."""
    sd = scenes.bunny(64, 64, n=12)
    rd = capi.render_desc(spp=8, max_depth=-1, rr_depth=5)
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    e = relmse(rgba, oref)
    _report("roughconductor", e, stats, ost)
    assert np.isfinite(film).all()
    np.testing.assert_allclose(film[..., 4], ofilm[..., 4], rtol=1e-5)
    assert e < EQUAL_SEED_RELMSE, e
    assert abs(int(stats.rays_closest) - int(ost.rays_closest)) <= 2e-3 * ost.rays_closest


def test_roughdielectric_environment_equal_seed(gpu_ctx):
    """C3's material (roughdielectric.cpp:58-190: reflection/refraction choice, eta-scaled radiance, RR with
    eta^2) plus a quad light and the constant environment (constant.cpp:21-81), depth 16."""
    sd = scenes.teapot(64, 64, n=12)
    rd = capi.render_desc(spp=8, max_depth=16, rr_depth=5)
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    e = relmse(rgba, oref)
    bad = bad_pixel_fraction(rgba, oref)
    _report("roughdielectric+env", e, stats, ost)
    print(f"[roughdielectric+env] pixels differing by more than 2%: {bad:.4f}")
    assert np.isfinite(film).all()
    assert e < CHAOTIC_RELMSE, e
    assert bad < CHAOTIC_BAD_PIXELS, bad
    assert abs(int(stats.rays_closest) - int(ost.rays_closest)) <= 2e-3 * ost.rays_closest


def _material_zoo(width=64, height=64):
    """Every BSDF type of include/misaki_b200.h in one scene, two area lights (Scene::sample_emitter_direct's
    N > 1 branch, scene.cpp:72-80) and an environment."""
    from misaki_render_b200.scene import SceneDescription, lookat
    from workloads import meshes
    sd = SceneDescription(width, height, fov=40.0, near_clip=0.1, far_clip=100.0,
                          to_world=lookat((0.0, 2.5, -6.0), (0.0, 0.8, 0.0), (0, 1, 0)))
    gv, gt = meshes.quad((-6, 0, -6), (-6, 0, 6), (6, 0, 6), (6, 0, -6))
    sd.add_mesh(gv, gt, sd.bsdf_diffuse((0.6, 0.5, 0.4)))
    for cx, half, rad in ((-2.0, 0.7, (12, 10, 8)), (2.0, 0.5, (6, 9, 14))):
        lv, lt = meshes.quad((cx - half, 4.0, -half), (cx + half, 4.0, -half), (cx + half, 4.0, half), (cx - half, 4.0, half))
        sd.add_mesh(lv, lt, sd.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=rad)
    gold = dict(eta=(0.143, 0.375, 1.442), k=(3.983, 2.386, 1.603))
    mats = [
        sd.bsdf_conductor(**gold),
        sd.bsdf_roughconductor(alpha=(0.05, 0.3), twosided=True, **gold),
        sd.bsdf_dielectric(int_ior=1.5, ext_ior=1.0),
        sd.bsdf_roughdielectric(int_ior=1.33, ext_ior=1.0, alpha=0.25, specular_transmittance=(0.9, 0.95, 1.0)),
        sd.bsdf_diffuse((0.2, 0.7, 0.3), twosided=True),
    ]
    for i, mat in enumerate(mats):
        v, t = meshes.cube_sphere(6, seed=10 + i, octaves=2, amplitude=0.1, radius=0.55, center=(-3.0 + 1.5 * i, 0.7, 0.3 * (i % 2)),
                                  normals=(i != 2), uvs=(i % 2 == 0))
        sd.add_mesh(v, t, mat, has_normals=(i != 2), has_uvs=(i % 2 == 0))
    # an open one-sided quad seen from behind (cos_theta(wi) <= 0 paths of diffuse.cpp:24-25)
    bv, bt = meshes.quad((-1, 0.2, 2.5), (1, 0.2, 2.5), (1, 2.2, 2.5), (-1, 2.2, 2.5))
    sd.add_mesh(bv, bt, sd.bsdf_diffuse((0.8, 0.8, 0.2)))
    sd.add_constant_environment((0.3, 0.35, 0.5))
    return sd


@pytest.mark.parametrize("max_depth,rr_depth,hide", [(6, 5, False), (-1, 3, True)])
def test_material_zoo_equal_seed(gpu_ctx, max_depth, rr_depth, hide):
    sd = _material_zoo()
    rd = capi.render_desc(spp=8, max_depth=max_depth, rr_depth=rr_depth, hide_emitters=hide)
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    e = relmse(rgba, oref)
    _report(f"zoo depth={max_depth}", e, stats, ost)
    assert np.isfinite(film).all()
    np.testing.assert_allclose(film[..., 3], ofilm[..., 3], rtol=1e-5)  # alpha channel: hide_emitters / misses
    bad = bad_pixel_fraction(rgba, oref)
    print(f"[zoo] pixels differing by more than 2%: {bad:.4f}")
    assert e < CHAOTIC_RELMSE, e  # the zoo holds a smooth and a rough dielectric
    assert bad < CHAOTIC_BAD_PIXELS, bad


def test_base_seed_and_convergence(gpu_ctx):
    """Measurement (ii) of SURVEY 8d in miniature: against a high-spp render, the GPU's error at N spp must
    equal the oracle's error at N spp (both are the same estimator), and a different base_seed must give a
    different but statistically equivalent image."""
    sd = scenes.cbox(32, 32)
    with capi.Scene(gpu_ctx, sd) as sc:
        ref_film, _ = sc.render(capi.render_desc(spp=4096, max_depth=5, base_seed=12345))
        ref = sc.develop(ref_film)
        a_film, _ = sc.render(capi.render_desc(spp=32, max_depth=5))
        b_film, _ = sc.render(capi.render_desc(spp=32, max_depth=5, base_seed=777))
        a, b = sc.develop(a_film), sc.develop(b_film)
    ofilm, _ = pyoracle.OracleScene(sd).render(capi.render_desc(spp=32, max_depth=5))
    o = pyoracle.develop(ofilm)
    ea, eb, eo = relmse(a, ref), relmse(b, ref), relmse(o, ref)
    print(f"[convergence] relMSE vs 4096spp: gpu={ea:.4e} gpu(seed 777)={eb:.4e} oracle={eo:.4e}")
    assert not np.array_equal(a_film, b_film)
    assert abs(ea - eo) <= 0.02 * eo  # same seeds: same estimator up to float noise
    assert 0.5 * eo < eb < 2.0 * eo   # other seeds: same variance


def test_checkerboard_textures_equal_seed(gpu_ctx):
    """"checkerboard" (textures/checkerboard.cpp) on a diffuse reflectance (nested, with texcoords),
    on an area light's radiance (no texcoords: si.uv = barycentrics, ps.uv = the warped sample) and on a rough
    conductor's specular_reflectance.  The texture only selects a child spectrum, so parity is the usual equal-seed
    bound; the untextured scene must differ visibly (the texture is actually evaluated)."""
    sd = scenes.checkers(96, 96)
    rd = capi.render_desc(spp=16, max_depth=6)
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    e = relmse(rgba, oref)
    _report("checkers", e, stats, ost)
    assert np.isfinite(film).all() and e < EQUAL_SEED_RELMSE, e
    assert bad_pixel_fraction(rgba, oref) < 0.01
    for s in sd.spectra:  # collapse every checkerboard onto its color0
        if s.kind == capi.SPEC_CHECKERBOARD:
            s.child1 = s.child0
    sd._cdesc = None
    with capi.Scene(gpu_ctx, sd) as sc:
        film0, _ = sc.render(rd)
        rgba0 = sc.develop(film0)
    assert relmse(rgba0, oref) > 100 * EQUAL_SEED_RELMSE


def _volpath_both(gpu_ctx, sd, rd):
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    assert np.isfinite(film).all()
    np.testing.assert_allclose(film[..., 4], ofilm[..., 4], rtol=1e-5)
    return rgba, oref, stats, ost


def test_volpath_without_media_equal_seed(gpu_ctx):
    """VolumetricPathTracer::sample (volpath.cpp:26-167) without any medium: pins the surface branch (emitter term on
    camera / delta chains, NEE without MIS, depth + 1 >= rr_depth roulette) on a unit-scale scene.
    Russian roulette runs before the next ray is traced on the GPU (after it in volpath.cpp:150-164, with no draw in
    between), so the GPU traces fewer closest-hit rays for the same image."""
    rgba, oref, stats, ost = _volpath_both(gpu_ctx, scenes.checkers(64, 64), capi.render_desc(spp=16, max_depth=6, rr_depth=3, integrator="volpath"))
    e = relmse(rgba, oref)
    _report("volpath checkers", e, stats, ost)
    assert e < EQUAL_SEED_RELMSE, e
    assert stats.rays_closest <= ost.rays_closest
    # Cornell-box scale (coordinates ~550): the visibility ray of eval_transmittance uses the scaled offset
    # RayEpsilon (1 + max|p|) of scene.cpp:91-93 instead of the unscaled one of scene.cpp:146-149 (see oracle.cpp)
    rgba, oref, stats, ost = _volpath_both(gpu_ctx, scenes.cbox(64, 64), capi.render_desc(spp=16, max_depth=6, rr_depth=3, integrator="volpath"))
    e = relmse(rgba, oref)
    _report("volpath cbox", e, stats, ost)
    assert e < EQUAL_SEED_RELMSE, e


def test_volpath_fog_and_scattering_interior_equal_seed(gpu_ctx):
    """Camera in a thin fog (sensor medium), a smooth-dielectric blob with a dense scattering interior and the fog
    as exterior medium: free-flight sampling, medium NEE with transmittance, isotropic phase sampling, medium
    transitions at the boundary (homogeneous.cpp, isotropic.cpp, scene.cpp:114-184, interaction.cpp:10-13).
    Refraction + many scattering events amplify ulp differences, hence the chaotic-scene bound."""
    sd = scenes.fog(96, 96)
    rd = capi.render_desc(spp=16, max_depth=-1, rr_depth=5, integrator="volpath")
    rgba, oref, stats, ost = _volpath_both(gpu_ctx, sd, rd)
    e = relmse(rgba, oref)
    _report("volpath fog", e, stats, ost)
    assert e < CHAOTIC_RELMSE, e
    assert bad_pixel_fraction(rgba, oref) < CHAOTIC_BAD_PIXELS
    assert stats.rays_closest <= ost.rays_closest  # roulette before the trace, see above
    # the medium matters: the same scene through the path tracer (media ignored) is a different image
    with capi.Scene(gpu_ctx, sd) as sc:
        film0, _ = sc.render(capi.render_desc(spp=16, max_depth=-1, rr_depth=5))
        assert relmse(sc.develop(film0), oref) > 20 * CHAOTIC_RELMSE


def test_volpath_bounded_depth_and_absorbing_fog(gpu_ctx):
    """max_depth cut-offs inside both branches (volpath.cpp:57-58,127-128) and a purely absorbing sensor medium."""
    sd = scenes.fog(64, 64, sensor_in_fog=False)
    sd.sensor_medium = sd.add_medium(sigma_a=(0.05, 0.1, 0.2), sigma_s=0.0)
    for depth in (1, 2, 4):
        rgba, oref, stats, ost = _volpath_both(gpu_ctx, sd, capi.render_desc(spp=8, max_depth=depth, integrator="volpath"))
        e = relmse(rgba, oref)
        _report(f"volpath depth {depth}", e, stats, ost)
        assert e < CHAOTIC_RELMSE, (depth, e)
        assert stats.bounces == depth


def test_render_edge_cases(gpu_ctx):
    """Ragged and degenerate inputs through the whole wavefront: a film whose size is no multiple of the 32-pixel
    reference block, of the film tile or of the warp; a mesh containing zero-area and duplicate triangles; several
    emitters (uniform light selection, scene.cpp:76-87) with one of them below the horizon; one-path batches."""
    from misaki_render_b200.scene import SceneDescription, lookat
    from workloads import meshes
    sd = SceneDescription(37, 23, fov=45.0, near_clip=0.1, far_clip=100.0, to_world=lookat((0, 1.5, -4), (0, 0.5, 0), (0, 1, 0)))
    gv, gt = meshes.quad((-3, 0, -3), (-3, 0, 3), (3, 0, 3), (3, 0, -3))
    gt = np.concatenate([gt, [[0, 0, 1], [1, 1, 1], gt[0]]]).astype(np.uint32)  # degenerate + duplicate triangles
    sd.add_mesh(gv, gt, sd.bsdf_diffuse((0.6, 0.6, 0.6)))
    for k, (x, rad) in enumerate([(-1.5, (30, 5, 5)), (1.5, (5, 5, 30))]):
        lv, lt = meshes.quad((x - .4, 2.5, -.4), (x + .4, 2.5, -.4), (x + .4, 2.5, .4), (x - .4, 2.5, .4))
        sd.add_mesh(lv, lt, sd.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=rad)
    lv, lt = meshes.quad((-1, -2, -1), (1, -2, -1), (1, -2, 1), (-1, -2, 1))  # a light under the floor: always occluded
    sd.add_mesh(lv, lt, sd.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=(9, 9, 9))
    rd = capi.render_desc(spp=6, max_depth=4)
    film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
    assert film.shape == (23, 37, 5) and np.isfinite(film).all()
    np.testing.assert_allclose(film[..., 4], ofilm[..., 4], rtol=1e-5)
    e = relmse(rgba, oref)
    _report("edge cases", e, stats, ost)
    assert e < EQUAL_SEED_RELMSE, e
    with capi.Scene(gpu_ctx, sd) as sc:
        tiny, st = sc.render(capi.render_desc(spp=6, max_depth=4, paths_per_batch=1))  # one sample of every pixel per batch
        assert st.batches == 6
        np.testing.assert_allclose(tiny, film, rtol=2e-5, atol=1e-6)
        none, st0 = sc.render(capi.render_desc(spp=6, max_depth=4, sample_begin=3, sample_end=3))  # empty sample range
        assert st0.paths == 0 and not none.any()
        zero, stz = sc.render(capi.render_desc(spp=2, max_depth=0))  # max_depth 0: the loop body never runs (path.cpp:33)
        assert stz.rays_closest == 0 and not zero[..., :3].any() and (zero[..., 4] > 0).all()


def test_render_scene_without_geometry(gpu_ctx):
    """No shapes at all: every camera ray escapes; with a constant environment each sample returns its radiance
    (path.cpp:34-41, constant.cpp:79-81), without one the image is black but the filter weights are still there."""
    from misaki_render_b200.scene import SceneDescription
    for env in (None, (0.5, 0.6, 0.8)):
        sd = SceneDescription(16, 16, fov=40.0)
        if env is not None:
            sd.add_constant_environment(env)
        rd = capi.render_desc(spp=4, max_depth=3)
        film, rgba, ofilm, oref, stats, ost = _both(gpu_ctx, sd, rd)
        assert np.isfinite(film).all() and stats.rays_shadow == 0
        np.testing.assert_allclose(film, ofilm, rtol=2e-5, atol=1e-7)
        assert (rgba[..., :3].max() > 0) == (env is not None)


@pytest.mark.parametrize("integrator", ["path", "volpath"])
def test_tail_kernel_is_bit_identical(monkeypatch, integrator):
    """Unbounded-depth jobs finish with k_tail (one per-path launch) once the queue is short.  It runs the same device
    functions in the same order as the wavefront stages, so the film must not change by a single bit, and the ray
    counts must agree."""
    sd = scenes.bunny(96, 96, n=16) if integrator == "path" else scenes.fog(96, 96)
    rd = capi.render_desc(spp=8, max_depth=-1, rr_depth=3, integrator=integrator)
    with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
        film_tail, st_tail = sc.render(rd)
    monkeypatch.setenv("MSK_TAIL_THRESHOLD", "0")  # read when the context is created
    with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
        film_wave, st_wave = sc.render(rd)
    assert st_tail.tail_rays_closest > 0 and st_wave.tail_rays_closest == 0
    np.testing.assert_array_equal(film_tail, film_wave)
    assert (st_tail.rays_closest, st_tail.rays_shadow, st_tail.shaded_vertices) == (st_wave.rays_closest, st_wave.rays_shadow, st_wave.shaded_vertices)
    assert st_tail.bounces == st_wave.bounces or st_wave.bounces == st_tail.bounces + 1  # the wavefront runs one bounce over an empty queue
    assert st_tail.kernel_launches < st_wave.kernel_launches


def test_graph_replay_and_tiny_scene_scheduling_are_bit_identical(monkeypatch):
    """Bounded-depth jobs replay a cached CUDA graph of their fixed launch sequence, and scenes of a few wide nodes always
    use the static traversal: scheduling only -- the film, the AOV film and the counters must not change by a bit, a second
    call (the replay proper) included, and a changed render description must re-capture."""
    sd = scenes.cbox(64, 48)
    rd, rd2 = capi.render_desc(spp=4, max_depth=5), capi.render_desc(spp=6, max_depth=3, sample_begin=1, sample_end=5)
    types = [capi.AOV_DEPTH, capi.AOV_INTEGRATOR_RGBA]
    with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
        film_a, st_a = sc.render(rd)
        film_b, st_b = sc.render(rd)      # replay of the cached graph
        film_c, st_c = sc.render(rd2)     # another sequence: re-capture
        film_d, _ = sc.render(rd)
        aov_a, ast_a = sc.render_aov(rd, types)
    np.testing.assert_array_equal(film_a, film_b)
    np.testing.assert_array_equal(film_a, film_d)
    assert (st_a.rays_closest, st_a.rays_shadow, st_a.kernel_launches, st_a.bounces) == (st_b.rays_closest, st_b.rays_shadow, st_b.kernel_launches, st_b.bounces)
    monkeypatch.setenv("MSK_GRAPH", "0")  # read when the context is created
    monkeypatch.setenv("MSK_STATIC_NODES", "0")
    with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
        film_p, st_p = sc.render(rd)
        film_q, st_q = sc.render(rd2)
        aov_p, ast_p = sc.render_aov(rd, types)
    np.testing.assert_array_equal(film_a, film_p)
    np.testing.assert_array_equal(film_c, film_q)
    np.testing.assert_array_equal(aov_a, aov_p)
    assert (st_a.rays_closest, st_a.rays_shadow, st_a.kernel_launches) == (st_p.rays_closest, st_p.rays_shadow, st_p.kernel_launches)
    assert (st_c.paths, st_c.rays_closest) == (st_q.paths, st_q.rays_closest) and ast_a.rays_closest == ast_p.rays_closest


def test_tiled_path_enumeration_is_bit_identical(monkeypatch):
    """Path slots enumerate the pixels of a sample in 8x4 tiles when the film size allows (coherent warps of camera
    rays); seeds depend on (pixel, sample) and the film records stay per pixel, so nothing in the film may change --
    including the AOV channels, which are captured per path slot."""
    sd = scenes.cbox(64, 48)
    rd = capi.render_desc(spp=4, max_depth=4)
    types = [capi.AOV_DEPTH, capi.AOV_UV, capi.AOV_INTEGRATOR_RGBA]
    with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
        film_t, _ = sc.render(rd)
        aov_t, _ = sc.render_aov(rd, types)
    monkeypatch.setenv("MSK_TILED_SLOTS", "0")
    with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
        film_l, _ = sc.render(rd)
        aov_l, _ = sc.render_aov(rd, types)
    np.testing.assert_array_equal(film_t, film_l)
    np.testing.assert_array_equal(aov_t, aov_l)


@pytest.mark.parametrize("spp", [4, 12, 32])
def test_samples_per_warp_enumeration_is_bit_identical(monkeypatch, spp):
    """A warp holds up to 32 consecutive samples of a compact pixel group (slot_decode): 1x1 pixel x 32 samples down to
    8x4 pixels x 1 sample, by what divides the batch's sample count (12 spp -> 4 samples of 4x2 pixels).  Every setting,
    and the sample-major round-1 order (0), must give the same film and AOV bits."""
    sd = scenes.cbox(64, 48)
    rd = capi.render_desc(spp=spp, max_depth=4)
    types = [capi.AOV_DEPTH, capi.AOV_UV, capi.AOV_INTEGRATOR_RGBA]
    films = []
    for spw in ["0", "1", "2", "8", "32"]:
        monkeypatch.setenv("MSK_SAMPLES_PER_WARP", spw)
        with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
            films.append((sc.render(rd)[0], sc.render_aov(rd, types)[0]))
    assert films[0][0].any()
    for f, a in films[1:]:
        np.testing.assert_array_equal(f, films[0][0])
        np.testing.assert_array_equal(a, films[0][1])


@pytest.mark.parametrize("spp", [5, 32])
def test_camera_ray_packets_are_bit_identical(monkeypatch, spp):
    """The camera-ray queue of a scene that is not tiny is traversed warp by warp as a packet (msk_traverse.cuh:
    traverse_packet: one walk of the tree per warp, children tested against the bounds of the 32 rays, every lane tests the
    triangles with its own ray); a ray accepts the same hits in the same order as alone, so film, AOVs and ray counters must
    equal those of per-ray traversal -- 32 samples of one pixel per warp, and 5 samples (warps of 8x4 pixels)."""
    sd = scenes.bunny(96, 64, n=20)
    rd = capi.render_desc(spp=spp, max_depth=-1, rr_depth=4)
    types = [capi.AOV_DEPTH, capi.AOV_UV, capi.AOV_INTEGRATOR_RGBA]
    out = []
    for packets in ["0", "1"]:
        monkeypatch.setenv("MSK_PACKET_CAMERA", packets)
        monkeypatch.setenv("MSK_STATIC_NODES", "0")  # no scene counts as tiny: the packet kernel runs whatever the mesh size
        with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
            film, st = sc.render(rd)
            out.append((film, sc.render_aov(rd, types)[0], int(st.rays_closest), int(st.rays_shadow)))
    assert out[0][0].any()
    np.testing.assert_array_equal(out[1][0], out[0][0])
    np.testing.assert_array_equal(out[1][1], out[0][1])
    assert out[1][2:] == out[0][2:]


@pytest.mark.parametrize("integrator,max_depth", [("path", -1), ("path", 3), ("volpath", -1)])
def test_batches_in_flight_are_bit_identical(monkeypatch, integrator, max_depth):
    """Up to four batches are in flight at once, each on its own stream and path pool (Renderer::render); the film
    kernels of consecutive batches are chained by events, so for a fixed batch partition the film and the ray counters
    must not depend on how many lanes run it -- polled unbounded jobs (tail kernel included), bounded ones, volpath."""
    sd = scenes.fog(64, 48, n=8) if integrator == "volpath" else scenes.cbox(64, 48)
    rd = capi.render_desc(spp=20, max_depth=max_depth, rr_depth=3, integrator=integrator, paths_per_batch=64 * 48 * 4)
    out = []
    for lanes in ["1", "2", "3", "4"]:
        monkeypatch.setenv("MSK_INFLIGHT", lanes)
        monkeypatch.setenv("MSK_GRAPH", "0")
        with capi.Context(0) as ctx, capi.Scene(ctx, sd) as sc:
            film, st = sc.render(rd)
            film2, _ = sc.render(rd)
            np.testing.assert_array_equal(film, film2)
            out.append((film, st))
    assert out[0][1].batches == 5 and out[0][0].any()
    for film, st in out[1:]:
        np.testing.assert_array_equal(film, out[0][0])
        assert (st.rays_closest, st.rays_shadow, st.shaded_vertices, st.batches) == \
            (out[0][1].rays_closest, out[0][1].rays_shadow, out[0][1].shaded_vertices, out[0][1].batches)


def test_develop_on_the_device_is_bit_identical(gpu_ctx):
    """msk_gpu_develop_dev (k_develop: HDRFilm::image on a device film) against msk_gpu_develop (host loop) and the oracle's
    develop, which is pinned bit for bit to the reference's compiled hdrfilm.cpp -- zero-weight pixels included."""
    import torch
    sd = scenes.cbox(64, 48)
    with capi.Scene(gpu_ctx, sd) as sc:
        film, _ = sc.render(capi.render_desc(spp=4, max_depth=3))
        film[5, 7, :] = 0.0   # a pixel no sample reached: W == 0 develops to 0, not NaN (hdrfilm.cpp:63-66)
        film[9, 3, 4] = 0.0
        host = sc.develop(film)
        stream = torch.cuda.ExternalStream(gpu_ctx.stream)
        with torch.cuda.stream(stream):
            d_film = torch.from_numpy(film).cuda()
            d_rgba = torch.empty((sd.height, sd.width, 4), dtype=torch.float32, device="cuda")
            sc.develop_dev(d_film.data_ptr(), d_rgba.data_ptr())
            dev = d_rgba.cpu().numpy()
    np.testing.assert_array_equal(dev, host)
    np.testing.assert_array_equal(dev, pyoracle.develop(film))
    assert np.isfinite(dev).all() and not dev[5, 7].any()


def test_c2_full_size_properties(gpu_ctx):
    """BASELINE configs[1] at its full size (512x512, 64 spp, unbounded depth): size-independent properties instead of a
    full oracle render -- bit-determinism run to run, invariance under a partition of the sample range (what multi-GPU
    sharding and batching rest on), and agreement of the image mean with the oracle's render of the first 4 samples per
    pixel (same seeds: those samples are a subset of the job's)."""
    sd = scenes.bunny(512, 512)
    rd = capi.render_desc(spp=64, max_depth=-1, rr_depth=5)
    with capi.Scene(gpu_ctx, sd) as sc:
        a, st = sc.render(rd)
        b, _ = sc.render(rd)
        np.testing.assert_array_equal(a, b)
        assert st.paths == 512 * 512 * 64 and np.isfinite(a).all()
        part, _ = sc.render(capi.render_desc(spp=64, max_depth=-1, rr_depth=5, sample_begin=0, sample_end=24))
        part, _ = sc.render(capi.render_desc(spp=64, max_depth=-1, rr_depth=5, sample_begin=24, sample_end=64, clear_film=False), film=part)
        np.testing.assert_allclose(part, a, rtol=5e-5, atol=1e-5)
        # filter weights: away from the film border every sample deposits a total weight of ~1 (normalised table, rfilter.cpp:12-27)
        np.testing.assert_allclose(a[8:-8, 8:-8, 4].mean(), 64.0, rtol=2e-2)
        first4, _ = sc.render(capi.render_desc(spp=64, max_depth=-1, rr_depth=5, sample_begin=0, sample_end=4))
        rgba4 = sc.develop(first4)
    ofilm, _ = pyoracle.OracleScene(sd).render(capi.render_desc(spp=64, max_depth=-1, rr_depth=5, sample_begin=0, sample_end=4))
    oref4 = pyoracle.develop(ofilm)
    e, bad = relmse(rgba4, oref4), bad_pixel_fraction(rgba4, oref4)
    print(f"[C2 full size, first 4 of 64 spp] relMSE={e:.3e}, pixels off by > 2 %: {bad * 512 * 512:.0f} of {512 * 512}")
    # 1 Mi unbounded-depth paths over 69 k triangles: a handful graze a silhouette edge where the watertight test and the
    # oracle's Moeller-Trumbore disagree about the hit primitive, and at 4 spp one such path moves its pixel visibly; the
    # 1e-6 bound holds for the image minus those pixels, the whole image is held to 1e-5 and to < 0.01 % of pixels off
    assert e < 10 * EQUAL_SEED_RELMSE and bad < 1e-4
    with capi.Scene(gpu_ctx, sd) as sc:  # the 64-spp image and its first 4 samples estimate the same mean
        rgba = sc.develop(a)
    np.testing.assert_allclose(rgba[..., :3].mean(), rgba4[..., :3].mean(), rtol=0.03)


def test_cbox_converges_to_the_reference_codes_image(gpu_ctx):
    """The GPU path against the reference's OWN code, without the oracle in between: tests/golden/ref_cbox48_converged.npz
    is BASELINE config C1's Cornell box rendered at 4096 spp by the reference's compiled render loop
    (tools/gen_golden_ref_converged.py; integrator.cpp, path.cpp, scene.cpp, imageblock.cpp, hdrfilm.cpp from
    /root/reference).  The reference never seeds per pixel, so the comparison is statistical: the GPU
    image approaches the fixture at the Monte-Carlo rate, with the noise level the reference's own renders show."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "ref_cbox48_converged.npz")
    sd = scenes.cbox(48, 48)
    e = {}
    with capi.Scene(gpu_ctx, sd) as sc:
        for spp in (256, 1024):
            film, _ = sc.render(capi.render_desc(spp=spp, max_depth=-1, rr_depth=5))
            e[spp] = relmse(sc.develop(film), g["image"])
    # measured with the oracle (whose films the GPU reproduces to relMSE ~1e-12 on this scene): 6.14e-4 and 1.66e-4
    # against the reference's own 6.18e-4 and 1.76e-4
    assert abs(e[256] / float(g["ref_relmse_256"]) - 1) < 0.2, e
    assert e[1024] < 1.5 * float(g["ref_relmse_1024"]), e
    assert 2.8 < e[256] / e[1024] < 5.5, e
