"""The BASELINE.json configurations as SceneDescriptions.

C1  cbox()      Cornell box, diffuse + area light, 256x256, 16 spp, max depth 5
C2  bunny()     bunny-class closed mesh (69 312 tris) with GGX rough conductor, 512x512, 64 spp
C3  teapot()    teapot-class closed mesh (150 528 tris) with rough dielectric + area light + constant
                environment, 1024x1024, 256 spp, depth 16
C4  cbox(1920, 1080), 4096 spp
C5  sphere10m() 9 998 244-triangle displaced sphere, intersection sweep
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from misaki_render_b200 import capi
from misaki_render_b200.scene import SceneDescription, lookat, scale, translate

from . import meshes

f32 = np.float32
ASSETS = Path(__file__).resolve().parent.parent / "assets"

CBOX_SHAPES = [  # order and values of reference assets/cbox/scene.xml:30-101
    ("luminaire", (0.936461, 0.740433, 0.705267)), ("floor", (0.885809, 0.698859, 0.666422)),
    ("ceiling", (0.885809, 0.698859, 0.666422)), ("back", (0.885809, 0.698859, 0.666422)),
    ("greenwall", (0.105421, 0.37798, 0.076425)), ("redwall", (0.570068, 0.0430135, 0.0443706)),
    ("smallbox", (0.45, 0.30, 0.90)), ("largebox", (0.45, 0.30, 0.90)),
]


def cbox(width=256, height=256):
    """Scene objects are children _arg_0.._arg_9 of <scene>; Scene::m_shapes follows std::map key order
    (properties.cpp:166-176), i.e. _arg_0 (integrator), _arg_1 (sensor), _arg_2.._arg_9 (shapes): file order."""
    sd = SceneDescription(width, height, fov=49.3077, near_clip=10, far_clip=2800,
                          to_world=lookat((278, 273, -800), (278, 273, -799), (0, 1, 0)))
    for name, refl in CBOX_SHAPES:
        tw = translate((0, -0.5, 0)) if name == "luminaire" else None
        v, t, hn, hu = meshes.read_obj(ASSETS / "cbox" / "meshes" / f"cbox_{name}.obj", tw)
        sd.add_mesh(v, t, sd.bsdf_diffuse(refl), radiance=(40, 40, 40) if name == "luminaire" else None,
                    has_normals=hn, has_uvs=hu)
    return sd


def _ground_and_light(sd, y=0.0, half=6.0, light_y=5.0, light_half=1.0, radiance=(20, 20, 20), ground_refl=(0.5, 0.5, 0.5)):
    gv, gt = meshes.quad((-half, y, -half), (-half, y, half), (half, y, half), (half, y, -half))
    sd.add_mesh(gv, gt, sd.bsdf_diffuse(ground_refl))
    lv, lt = meshes.quad((-light_half, light_y, -light_half), (light_half, light_y, -light_half),
                         (light_half, light_y, light_half), (-light_half, light_y, light_half))
    sd.add_mesh(lv, lt, sd.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=radiance)


def bunny(width=512, height=512, n=76):
    """C2: bunny-class mesh (cube-sphere n=76 -> 69 312 triangles, seed 1, 3 octaves, vertex normals) with the
    Figure_2 material (results/Figure_2_RoughConductor/roughconductor.xml:20-26) over a diffuse ground."""
    sd = SceneDescription(width, height, fov=35.0, near_clip=0.1, far_clip=100.0,
                          to_world=lookat((0.0, 2.2, -4.5), (0.0, 0.9, 0.0), (0, 1, 0)))
    _ground_and_light(sd)
    v, t = meshes.cube_sphere(n, seed=1, octaves=3, amplitude=0.15, radius=0.9, center=(0, 1.05, 0), normals=True)
    mat = sd.bsdf_roughconductor(eta=(0.200438, 0.924033, 1.10221), k=(3.91295, 2.45285, 2.14219), alpha=0.1,
                                 specular_reflectance=(1.0, 1.0, 1.0), distribution="ggx")
    sd.add_mesh(v, t, mat, has_normals=True)
    return sd


def teapot(width=1024, height=1024, n=112):
    """C3: teapot-class mesh (cube-sphere n=112 -> 150 528 triangles, seed 2, normals + uvs) with the Figure_3
    rough dielectric (results/Figure_3_RoughDielectric/roughdielectric.xml:19-25), a quad light and the
    constant environment of assets/teapot-full/scene.xml:86-91."""
    sd = SceneDescription(width, height, fov=35.0, near_clip=0.1, far_clip=100.0,
                          to_world=lookat((0.0, 2.4, -4.8), (0.0, 0.8, 0.0), (0, 1, 0)))
    _ground_and_light(sd, radiance=(15, 15, 15), ground_refl=(0.6, 0.55, 0.5))
    v, t = meshes.cube_sphere(n, seed=2, octaves=3, amplitude=0.10, radius=0.85, center=(0, 0.95, 0), stretch=(1.25, 0.85, 1.0),
                              normals=True, uvs=True)
    mat = sd.bsdf_roughdielectric(int_ior=1.5, ext_ior=1.0, alpha=0.1, distribution="ggx")
    sd.add_mesh(v, t, mat, has_normals=True, has_uvs=True)
    sd.add_constant_environment((0.5, 0.6, 0.8))
    return sd


CBOX_UNIFORM = [  # (mesh, reflectance, radiance or None): the Cornell box with wavelength-independent spectra
    ("luminaire", 0.78, 12.0), ("floor", 0.73, None), ("ceiling", 0.73, None), ("back", 0.73, None),
    ("greenwall", 0.25, None), ("redwall", 0.45, None), ("smallbox", 0.6, None), ("largebox", 0.5, None),
]


def cbox_uniform(width=64, height=64):
    """The Cornell box geometry with `uniform` spectra everywhere: the scene on which the reference's own compiled
    PathTracer::sample is compared with the oracle (tools/gen_golden_ref_math.py, tests/test_oracle_ref_math.py) --
    constant spectra keep the comparison on the integrator / emitter / BSDF arithmetic."""
    sd = SceneDescription(width, height, fov=49.3077, near_clip=10, far_clip=2800,
                          to_world=lookat((278, 273, -800), (278, 273, -799), (0, 1, 0)))
    for name, refl, rad in CBOX_UNIFORM:
        tw = translate((0, -0.5, 0)) if name == "luminaire" else None
        v, t, hn, hu = meshes.read_obj(ASSETS / "cbox" / "meshes" / f"cbox_{name}.obj", tw)
        sd.add_mesh(v, t, sd.bsdf_diffuse(float(refl)), radiance=None if rad is None else sd.spectrum_uniform(float(rad)),
                    has_normals=hn, has_uvs=hu)
    return sd


OPEN_UNIFORM_ENV = 0.6


def open_uniform(width=64, height=64):
    """An open scene with uniform spectra for the second pinned path-tracer comparison: ground quad, a small blob with
    vertex normals, a quad light and a `constant` environment (two emitters: uniform light selection, scene.cpp:76-87;
    escaped BSDF rays: the stale-pdf environment quirk q4 / q8 of path.cpp:90-108).  Returns (scene, per-mesh
    (reflectance, radiance or None))."""
    sd = SceneDescription(width, height, fov=40.0, near_clip=0.1, far_clip=100.0,
                          to_world=lookat((0.0, 2.2, -4.5), (0.0, 0.7, 0.0), (0, 1, 0)))
    params = []
    gv, gt = meshes.quad((-4, 0, -4), (-4, 0, 4), (4, 0, 4), (4, 0, -4))
    sd.add_mesh(gv, gt, sd.bsdf_diffuse(0.55)); params.append((0.55, None))
    v, t = meshes.cube_sphere(4, seed=9, octaves=2, amplitude=0.08, radius=0.8, center=(0, 0.9, 0), normals=True)
    sd.add_mesh(v, t, sd.bsdf_diffuse(0.7), has_normals=True); params.append((0.7, None))
    lv, lt = meshes.quad((-1, 3.5, -1), (1, 3.5, -1), (1, 3.5, 1), (-1, 3.5, 1))
    sd.add_mesh(lv, lt, sd.bsdf_diffuse(0.5), radiance=sd.spectrum_uniform(9.0)); params.append((0.5, 9.0))
    sd.add_constant_environment(sd.spectrum_uniform(OPEN_UNIFORM_ENV))
    return sd, params


def film_samples(W, H, n, seed, nch=5):
    """Seeded (position, channel values) samples for the pinned film-accumulation comparison (shared by
    tools/gen_golden_ref_math.py and tests/test_oracle_ref_math.py so the fixture only stores hashes and probe pixels).
    The first eight positions sit on film corners, block corners (32) and pixel centres."""
    rng = np.random.default_rng(seed)
    smp = np.concatenate([rng.uniform(0, W, (n, 1)), rng.uniform(0, H, (n, 1)), rng.uniform(0, 2, (n, nch))], axis=1).astype(f32)
    smp[:8, :2] = [[0, 0], [W - 1e-3, H - 1e-3], [0.5, 0.5], [31.999, 31.999], [32, 32], [W / 2, 0], [0, H / 2], [W - 0.5, 0.25]]
    return np.ascontiguousarray(smp, dtype=f32)


def checkers(width=96, height=96, n=12):
    """Textured scene for the "checkerboard" texture (textures/checkerboard.cpp): a ground quad with
    texcoords whose reflectance is a checkerboard of a colour and a NESTED checkerboard, a quad light WITHOUT
    texcoords whose radiance is a checkerboard (si.uv = barycentrics / the warped sample, mesh.cpp:65,114), and a
    rough conductor with a checkerboard specular_reflectance on interpolated uvs."""
    sd = SceneDescription(width, height, fov=40.0, near_clip=0.1, far_clip=100.0,
                          to_world=lookat((0.0, 2.6, -5.0), (0.0, 0.7, 0.0), (0, 1, 0)))
    half = 5.0
    gv = np.zeros((4, 8), f32)
    gv[:, :3] = [(-half, 0, -half), (-half, 0, half), (half, 0, half), (half, 0, -half)]
    gv[:, 6:8] = [(0, 0), (0, 1), (1, 1), (1, 0)]
    gt = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    fine = sd.spectrum_checkerboard(sd.spectrum_srgb((0.8, 0.7, 0.2)), sd.spectrum_srgb((0.1, 0.1, 0.5)), to_uv=scale((32, 32, 1)))
    # the top-left 3x3 is kept (transform.h:142-148): the z column of the 4x4 acts as the uv offset
    m = scale((4, 4, 1)); m[0, 2], m[1, 2] = 0.25, 0.125
    ground = sd.spectrum_checkerboard(sd.spectrum_srgb((0.75, 0.75, 0.75)), fine, to_uv=m)
    sd.add_mesh(gv, gt, sd.bsdf_diffuse(ground), has_uvs=True)
    lv, lt = meshes.quad((-1.2, 4.5, -1.2), (1.2, 4.5, -1.2), (1.2, 4.5, 1.2), (-1.2, 4.5, 1.2))
    light = sd.spectrum_checkerboard(sd.spectrum_srgb_d65((30, 24, 18)), sd.spectrum_srgb_d65((4, 8, 20)), to_uv=scale((3, 3, 1)))
    sd.add_mesh(lv, lt, sd.bsdf_diffuse((0.5, 0.5, 0.5)), radiance=light)
    v, t = meshes.cube_sphere(n, seed=5, octaves=2, amplitude=0.05, radius=0.8, center=(0, 0.9, 0), normals=True, uvs=True)
    spec = sd.spectrum_checkerboard(sd.spectrum_srgb((1.0, 1.0, 1.0)), sd.spectrum_srgb((0.9, 0.3, 0.2)), to_uv=scale((6, 6, 1)))
    mat = sd.bsdf_roughconductor(eta=(0.200438, 0.924033, 1.10221), k=(3.91295, 2.45285, 2.14219), alpha=0.15,
                                 specular_reflectance=spec)
    sd.add_mesh(v, t, mat, has_normals=True, has_uvs=True)
    return sd


def fog(width=96, height=96, n=16, sensor_in_fog=True):
    """Scene for the "volpath" integrator (integrators/volpath.cpp): the camera sits in a thin
    homogeneous fog (sensor medium, sensor.cpp:12-18) that is also the exterior medium of a smooth-dielectric blob
    whose interior is a dense, coloured scattering medium (shape.cpp:28-39); diffuse ground + quad light as in C2."""
    sd = SceneDescription(width, height, fov=35.0, near_clip=0.1, far_clip=100.0,
                          to_world=lookat((0.0, 2.2, -4.5), (0.0, 0.9, 0.0), (0, 1, 0)))
    _ground_and_light(sd, radiance=(25, 25, 25))
    haze = sd.add_medium(sigma_a=(0.01, 0.01, 0.01), sigma_s=(0.06, 0.07, 0.09))
    wax = sd.add_medium(sigma_a=(0.2, 0.6, 1.2), sigma_s=(3.0, 2.5, 2.0))
    v, t = meshes.cube_sphere(n, seed=4, octaves=2, amplitude=0.08, radius=0.85, center=(0, 1.0, 0), normals=True)
    sd.add_mesh(v, t, sd.bsdf_dielectric(int_ior=1.33, ext_ior=1.0), has_normals=True, interior_medium=wax, exterior_medium=haze)
    if sensor_in_fog:
        sd.sensor_medium = haze
    return sd


def sphere10m(nu=3163, nv=1582, width=64, height=64):
    """C5: the intersection-sweep mesh (single geomID)."""
    sd = SceneDescription(width, height, fov=40.0, near_clip=0.01, far_clip=100.0,
                          to_world=lookat((0.0, 0.0, -3.0), (0.0, 0.0, 0.0), (0, 1, 0)))
    v, t = meshes.sphere_grid(nu, nv)
    sd.add_mesh(v, t, sd.bsdf_diffuse((0.5, 0.5, 0.5)))
    return sd


def primary_rays(sd, res):
    """res x res pinhole rays through pixel centres of a res x res film with sd's camera (C5 'primary')."""
    cam = SceneDescription(res, res, fov=sd.fov, near_clip=sd.near_clip, far_clip=sd.far_clip, to_world=sd.to_world).camera()
    s2c = np.array(cam.sample_to_camera[:], dtype=f32).reshape(4, 4)
    c2w = np.array(cam.to_world[:], dtype=f32).reshape(4, 4)
    ys, xs = np.meshgrid(np.arange(res, dtype=f32) + f32(0.5), np.arange(res, dtype=f32) + f32(0.5), indexing="ij")
    p = np.stack([xs.reshape(-1), ys.reshape(-1), np.zeros(res * res, f32), np.ones(res * res, f32)], axis=-1)
    q = (p @ s2c.T).astype(f32)
    d = q[:, :3] / q[:, 3:4]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    rays = np.zeros(res * res, dtype=capi.RAY_DTYPE)
    rays["o"] = c2w[:3, 3]
    rays["d"] = (d @ c2w[:3, :3].T).astype(f32)
    rays["tmin"] = f32(sd.near_clip) / d[:, 2]
    rays["tmax"] = f32(sd.far_clip) / d[:, 2]
    return rays


def secondary_rays(sd_meshes, rays, hits, seed=0):
    """C5 'incoherent secondary': from each primary hit one cosine-hemisphere direction about the geometric
    normal (numpy Generator seeded per call), origin offset as SceneInteraction::spawn_ray."""
    ok = np.isfinite(hits["t"])
    h = hits[ok]
    verts, tris = sd_meshes
    f = tris[h["prim"]]
    p0, p1, p2 = verts[f[:, 0], :3], verts[f[:, 1], :3], verts[f[:, 2], :3]
    b1, b2 = h["u"][:, None], h["v"][:, None]
    p = (p0 * (1 - b1 - b2) + p1 * b1 + p2 * b2).astype(f32)
    n = np.cross(p1 - p0, p2 - p0)
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(f32)
    n = np.where((np.einsum("ij,ij->i", n, rays["d"][ok]) > 0)[:, None], -n, n)
    rng = np.random.default_rng(seed)
    u1, u2 = rng.random(p.shape[0]), rng.random(p.shape[0])
    r, phi = np.sqrt(u1), 2 * np.pi * u2
    lx, ly, lz = r * np.cos(phi), r * np.sin(phi), np.sqrt(np.maximum(0.0, 1 - u1))
    a = np.where(np.abs(n[:, 0:1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    t1 = np.cross(n, a); t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(n, t1)
    d = (t1 * lx[:, None] + t2 * ly[:, None] + n * lz[:, None]).astype(f32)
    out = np.zeros(p.shape[0], dtype=capi.RAY_DTYPE)
    out["o"], out["d"] = p, d
    out["tmin"] = (1 + np.abs(p).max(axis=1)) * f32(8.940697e-05)
    out["tmax"] = np.inf
    return out
