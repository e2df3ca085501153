"""The oracle's math layer pinned against the reference ITSELF: tests/golden/ref_math.json holds outputs of the
reference's own headers (include/misaki/core/{mathutils,warp,frame,spectrum,distribution}.h,
include/misaki/render/{fresnel,microfacet,srgb}.h, CIE table of src/librender/spectrum.cpp) compiled in the build
container by oracle/Makefile.ref against a minimal Eigen stand-in and called by tools/gen_golden_ref_math.py.
Every value is compared BIT FOR BIT (same compiler, same libm, -ffp-contract=off on both sides); the only tolerance
is for NaN payloads.  Covered: PCG32 and the sampler's float construction, the four warps, coordinate_system / Frame,
Fresnel (dielectric, conductor), reflect / refract, GGX eval / pdf / sample / G / smith_g1, sample_wavelength,
spectrum_to_xyz + xyz_to_srgb, srgb_model_eval, Distribution1D::init / sample_reuse, and SmoothDiffuse::sample / eval /
pdf from the reference's own bsdfs/diffuse.cpp (the one BSDF plugin its build compiles)."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "ref_math.json").read_text())
f32 = np.float32


def F(bits):
    return np.array(bits, dtype=np.uint32).view(f32)


def bits_of(a):
    return np.ascontiguousarray(a, dtype=f32).reshape(-1).view(np.uint32)


def same_bits(got, want_bits, what):
    got = np.ascontiguousarray(got, dtype=f32).reshape(-1)
    want = F(want_bits)
    assert got.shape == want.shape, what
    ok = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert ok.all(), f"{what}: got {got} want {want}"


def orc_math(which, values, nout):
    a = np.ascontiguousarray(values, dtype=f32)
    out = np.zeros(nout, f32)
    po.lib().orc_math(which, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def test_pcg32_and_sampler_floats():
    for c in GOLDEN["pcg32"]:
        assert [int(x) for x in po.pcg32_uints(c["state"], c["seq"], 16)] == c["uints"]
        if c["seq"] == 0xda3e39cb94b95bdb:  # IndependentSampler::seed uses PCG32_DEFAULT_STREAM (independent.cpp:20-26)
            same_bits(po.pcg32_floats(c["state"], 16), c["floats"], f"next_float32 state={c['state']}")


def test_warps():
    for c in GOLDEN["warp"]:
        u, v = F(c["uv"])
        same_bits(po.warp(c["which"], float(u), float(v)), c["out"], f"warp {c['which']} ({u}, {v})")


def test_coordinate_system_and_frame():
    for c in GOLDEN["frame"]:
        st = orc_math(0, F(c["n"]), 6)
        same_bits(st[:3], c["s"], "coordinate_system s"); same_bits(st[3:], c["t"], "coordinate_system t")
        lw = orc_math(1, np.concatenate([F(c["n"]), F(c["v"])]), 6)
        same_bits(lw[:3], c["local"], "Frame::to_local"); same_bits(lw[3:], c["world"], "Frame::to_world")


def test_fresnel_dielectric_conductor_reflect_refract():
    for c in GOLDEN["fresnel"]:
        same_bits(po.fresnel(float(F(c["cos"])[0]), float(F(c["eta"])[0])), c["out"], "fresnel")
    for c in GOLDEN["fresnel_conductor"]:
        eta, k = F(c["eta"]), F(c["k"])
        got = po.fresnel_conductor(float(F(c["cos"])[0]), np.append(eta, eta[0]), np.append(k, k[0]))  # the oracle's is 4-wide
        same_bits(got[:3], c["out"], "fresnel_conductor")
    for c in GOLDEN["reflect_refract"]:
        out = orc_math(2, np.concatenate([F(c["wi"]), F(c["m"]), F(c["ct"]), F(c["ti"])]), 6)
        same_bits(out[:3], c["reflect"], "reflect(wi, m)"); same_bits(out[3:], c["refract"], "refract(wi, m, cos_theta_t, eta_ti)")


def test_ggx_microfacet_distribution():
    for c in GOLDEN["ggx"]:
        au, av = float(F(c["au"])[0]), float(F(c["av"])[0])
        a, b, cc, want = F(c["a"]), F(c["b"]), F(c["c"]), c["out"]
        w = c["which"]
        if w == 0:
            same_bits(po.ggx(0, au, av, a)[:1], want[:1], "ggx eval")
        elif w == 1:
            same_bits(orc_math(3, np.concatenate([[au, av], a, b, b]), 2)[:1], want[:1], "ggx pdf")
        elif w == 2:
            same_bits(po.ggx(1, au, av, a, b), want, "ggx sample")
        elif w == 3:
            same_bits(orc_math(3, np.concatenate([[au, av], a, b, cc]), 2)[1:2], want[:1], "ggx G")
        else:
            same_bits(po.ggx(2, au, av, a, b)[:1], want[:1], "ggx smith_g1")


def test_spectral_sampling_and_colour():
    for c in GOLDEN["sample_wavelength"]:
        wl, w = po.sample_wavelength(float(F(c["u"])[0]))
        same_bits(wl, c["wl"], "sample_wavelength wavelengths"); same_bits(w, c["weight"], "sample_wavelength weights")
    for c in GOLDEN["spectrum_to_xyz"]:
        xyz = po.spectrum_to_xyz(F(c["value"]), F(c["wl"]))
        same_bits(xyz, c["xyz"], "spectrum_to_xyz")
        same_bits(orc_math(4, xyz, 3), c["rgb"], "xyz_to_srgb")
    for c in GOLDEN["srgb_model_eval"]:
        same_bits(po.srgb_model_eval(F(c["c"]), F(c["wl"])), c["out"], "srgb_model_eval")


def test_distribution1d():
    for c in GOLDEN["distribution"]:
        pdf, u = F(c["pdf"]), F(c["u"])
        idx = np.empty(len(u), np.uint32); re = np.empty(len(u), f32); cdf = np.empty(len(pdf) + 1, f32)
        po.lib().orc_distribution_sample_reuse(pdf.ctypes.data_as(C.c_void_p), C.c_size_t(len(pdf)), u.ctypes.data_as(C.c_void_p), C.c_size_t(len(u)),
                                               idx.ctypes.data_as(C.c_void_p), re.ctypes.data_as(C.c_void_p), cdf.ctypes.data_as(C.c_void_p))
        same_bits(cdf, c["cdf"], "Distribution1D cdf")
        assert [int(i) for i in idx] == c["index"]
        same_bits(re, c["reused"], "sample_reuse")


def test_smooth_diffuse_plugin():
    """src/librender/bsdfs/diffuse.cpp itself (compiled over stand-ins for Object / Properties / Texture)."""
    from misaki_render_b200.scene import SceneDescription
    scenes_by_refl = {}
    for c in GOLDEN["bsdf_diffuse"]:
        refl = float(F(c["reflectance"])[0])
        if refl not in scenes_by_refl:
            sd = SceneDescription(8, 8)
            b = sd.bsdf_diffuse(refl)  # float -> "uniform" spectrum: the constant texture of the reference-side wrapper
            sd.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], f32), np.array([[0, 1, 2]], np.uint32), b)
            scenes_by_refl[refl] = (po.OracleScene(sd), b)
        osc, b = scenes_by_refl[refl]
        r = osc.bsdf(b, F(c["wi"]), F(c["wl"]), F(c["smp"]), F(c["wo"]))
        got = np.concatenate([r["wo"], [r["pdf"], r["eta"], float(r["type"])], r["weight"]]).astype(f32)
        same_bits(got, c["sample"], "SmoothDiffuse::sample")
        same_bits(r["eval"], c["eval"], "SmoothDiffuse::eval")
        same_bits([r["eval_pdf"]], c["pdf"], "SmoothDiffuse::pdf")


def test_gaussian_filter_sampler_and_spectrum_plugins():
    """filters/gaussian.cpp + rfilter.cpp, samplers/independent.cpp + sampler.cpp, spectra/regular.cpp, spectra/uniform.cpp."""
    from misaki_render_b200 import capi
    from misaki_render_b200.scene import SceneDescription
    for c in GOLDEN["gaussian_filter"]:
        radius, table = po.gaussian_filter(float(F(c["stddev"])[0]))
        same_bits([radius], c["radius"], "GaussianFilter radius")
        same_bits(table, c["table"], "ReconstructionFilter::init_discretization table")
        assert int(np.ceil(radius - 0.5)) == c["border"]
    for c in GOLDEN["independent_sampler"]:  # next2d() = { next1d(), next1d() }: x first
        n = c["n1"] + 2 * c["n2"]
        same_bits(po.pcg32_floats(c["seed"], n, c["base_seed"]), c["out"], "IndependentSampler seed / next1d / next2d")
    for c in GOLDEN["regular_spectrum"]:
        sd = SceneDescription(8, 8)
        lo, hi = F(c["range"])
        sid = sd._add_spec(capi.SPEC_REGULAR, table=F(c["values"]).copy(), lmin=float(lo), lmax=float(hi))
        sd.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], f32), np.array([[0, 1, 2]], np.uint32), sd.bsdf_diffuse(0.5))
        same_bits(po.OracleScene(sd).spectrum_eval(sid, F(c["wl"])), c["out"], "RegularSpectrum::eval")
    sd = SceneDescription(8, 8)
    sid = sd.spectrum_uniform(float(F(GOLDEN["uniform_spectrum"][0]["value"])[0]))
    sd.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], f32), np.array([[0, 1, 2]], np.uint32), sd.bsdf_diffuse(0.5))
    osc = po.OracleScene(sd)
    for c in GOLDEN["uniform_spectrum"]:
        same_bits(osc.spectrum_eval(sid, F(c["wl"])), c["out"], "UniformSpectrum::eval")


def test_mesh_hit_reconstruction_and_sampling():
    """src/librender/{mesh,shape,records,interaction}.cpp themselves: Mesh::compute_scene_interaction +
    PreliminaryIntersection::compute_scene_interaction + initialize_sh_frame (position from barycentrics, geometric and
    shading normals, uv, dp_du / dp_dv with and without texcoords, the shading frame, wi), Mesh::area_distr_build,
    Mesh::sample_position, Shape::sample_direct / pdf_direct."""
    L = po.lib()
    for m in GOLDEN["mesh"]:
        v = F(m["verts"]).reshape(-1, 8).copy(); t = np.array(m["tris"], np.uint32).reshape(-1, 3).copy()
        nv, nt = v.shape[0], t.shape[0]
        vp, tp = v.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p)
        for h in m["hits"]:
            u, vv, tt = F(h["uvt"])
            o, d = F(h["o"]).copy(), F(h["d"]).copy()
            out = np.empty(27, f32)
            L.orc_mesh_interaction(vp, nv, tp, nt, m["normals"], m["uvs"], h["prim"], C.c_float(u), C.c_float(vv), C.c_float(tt),
                                   o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            same_bits(out, h["out"], f"compute_scene_interaction normals={m['normals']} uvs={m['uvs']} prim={h['prim']}")
        cdf = np.empty(nt + 1, f32)
        for s_ in m["samples"]:
            smp, ref_p = F(s_["sample"]).copy(), F(s_["ref_p"]).copy()
            out = np.empty(22, f32)
            L.orc_mesh_sampling(vp, nv, tp, nt, m["normals"], m["uvs"], smp.ctypes.data_as(C.c_void_p), ref_p.ctypes.data_as(C.c_void_p),
                                out.ctypes.data_as(C.c_void_p), cdf.ctypes.data_as(C.c_void_p))
            same_bits(out, s_["out"], f"sample_position / sample_direct normals={m['normals']} uvs={m['uvs']}")
        same_bits(cdf, m["cdf"], "area distribution")


def test_path_tracer_sample_on_the_cornell_box():
    """PathTracer::sample ITSELF (integrators/path.cpp) on the reference's own Scene::sample_emitter_direct /
    pdf_emitter_direct, AreaLight, SmoothDiffuse, Mesh / Shape / interaction code and IndependentSampler, compiled from
    where they lie (oracle/ref_path_wrap.cpp): 400 paths through the Cornell box with uniform spectra, unbounded depth,
    Russian roulette from depth 5 (hard-wired in path.cpp:135-136).  Ray-triangle intersection -- Embree's job in the
    reference -- is the same brute-force Moeller-Trumbore routine on both sides, so what is compared is the integrator loop
    with its MIS / emitter / roulette logic and every quirk the oracle restates (q2, q4, q7, q8).
    One switch: path.cpp:71-72 passes `sampler->next1d(), sampler->next2d()` as two arguments of one call, whose evaluation
    order C++ leaves unspecified.  GCC (which compiled the golden vectors) evaluates right to left; the determinism contract
    of the oracle and the GPU fixes left to right.  The oracle replays the vectors with the test-only right-to-left
    switch -- everything else is the code the GPU is compared against."""
    from misaki_render_b200 import capi
    from workloads import scenes
    rd = capi.render_desc(spp=1, max_depth=-1, rr_depth=5)
    # second scene: open, two emitters (quad light + `constant` environment, emitters/constant.cpp): uniform light
    # selection (scene.cpp:76-87), escaped BSDF rays with the stale NEE record (path.cpp:90-108, quirks q4 / q8), shading normals
    for key, sd, floor in [("path_sample", scenes.cbox_uniform(64, 64), 0.4), ("path_sample_env", scenes.open_uniform(64, 64)[0], 0.95)]:
        osc = po.OracleScene(sd)
        nonzero = 0
        for c in GOLDEN[key]:
            tmin, tmax = F(c["t"])
            got = osc.sample_ray(rd, c["seed"], F(c["o"]), F(c["d"]), float(tmin), float(tmax), F(c["wl"]), bsdf_draws_right_to_left=True)
            same_bits(got, c["out"], f"PathTracer::sample {key} seed={c['seed']}")
            nonzero += bool(np.any(got != 0))
        # (44 % of the Cornell-box camera rays miss the box: a square film around it); the comparison is not vacuous
        assert nonzero > floor * len(GOLDEN[key]), (key, nonzero)


def test_aov_integrator_sample():
    """AOVIntegrator::sample ITSELF (integrators/aov.cpp:87-144: depth, position, uv, geometric / shading normal and the
    nested path tracer's RGBA through spectrum_to_xyz + xyz_to_srgb), compiled like the path tracer.  On a miss the
    reference reads position / normals / uv of an uninitialised interaction; the oracle (and the GPU) write zeros there,
    so those 11 channels are compared on hits only."""
    from misaki_render_b200 import capi
    from workloads import scenes
    osc = po.OracleScene(scenes.open_uniform(64, 64)[0])
    rd = capi.render_desc(spp=1, max_depth=-1, rr_depth=5)
    types = np.array([capi.AOV_DEPTH, capi.AOV_POSITION, capi.AOV_UV, capi.AOV_GEO_NORMAL, capi.AOV_SH_NORMAL, capi.AOV_INTEGRATOR_RGBA], np.int32)
    hits = 0
    for c in GOLDEN["aov_sample"]:
        tmin, tmax = F(c["t"])
        o, d, wl = F(c["o"]).copy(), F(c["d"]).copy(), F(c["wl"]).copy()
        aovs, res = np.zeros(16, f32), np.zeros(4, f32)
        assert po.lib().orc_aov_sample_ray(osc.h, C.byref(rd), C.c_uint64(c["seed"]), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                           C.c_float(float(tmin)), C.c_float(float(tmax)), wl.ctypes.data_as(C.c_void_p), 1,
                                           types.ctypes.data_as(C.c_void_p), len(types), aovs.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p)) == 0
        want = F(c["out"])
        hit = want[0] != 0
        hits += bool(hit)
        sel = slice(0, 16) if hit else [0, 12, 13, 14, 15]
        same_bits(aovs[sel], [c["out"][i] for i in (range(16) if hit else sel)], f"AOVIntegrator::sample aovs seed={c['seed']}")
        same_bits(res, c["out"][16:], "AOVIntegrator::sample result")
    assert hits > 0.5 * len(GOLDEN["aov_sample"])


def test_film_accumulation():
    """src/librender/imageblock.cpp itself: ImageBlock::put(pos, value) (the filtered splat with its block-relative
    position arithmetic), ImageBlock::put(block) / accumulate_2d (the merge of the padded block into the film) and
    BlockGenerator (the spiral), driven as SamplingIntegrator::render + HDRFilm drive them, on films whose size is no
    multiple of the block size.  The whole film must match bit for bit (SHA-256) -- this is the film the GPU gather is
    compared with."""
    import hashlib
    from workloads import scenes
    for c in GOLDEN["film"]:
        W, H, nch = c["W"], c["H"], 5
        smp = scenes.film_samples(W, H, c["n"], c["sample_seed"])
        assert hashlib.sha256(smp.tobytes()).hexdigest() == c["samples_sha256"], "the seeded samples changed (numpy Generator stream?)"
        nb = len(c["order"]) // 4
        film = np.empty((H, W, nch), f32); order = np.empty((nb, 4), np.int32)
        assert po.lib().orc_film_accumulate(C.c_float(float(F(c["stddev"])[0])), W, H, nch, c["block_size"], smp.ctypes.data_as(C.c_void_p),
                                            C.c_size_t(c["n"]), film.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p)) == 0
        assert [int(v) for v in order.reshape(-1)] == c["order"], "BlockGenerator spiral order"
        same_bits(np.stack([film[y, x] for y, x in c["probes"]]), c["probe_values"], "film probe pixels")
        assert hashlib.sha256(film.tobytes()).hexdigest() == c["sha256"], f"film {W}x{H}"


def test_obj_loader_of_the_host_frontend(tmp_path):
    """The PRODUCT's OBJ loader (misaki_render_b200/host/plugins.cpp: OBJMesh) against the reference's own
    shapes/obj.cpp compiled here: same vertex de-duplication order, quad split (v1 v2 v3)(v4 v1 v3), texcoord flip and
    8-float layout, on the Cornell-box meshes and a synthetic file with quads, split vertices, normals and texcoords --
    this fixes geomID / primID / vertex order at the boundary.  (Attributes a file does not have are uninitialised
    memory in the reference; they are zero here and skipped.)"""
    from misaki_render_b200 import host_api
    root = Path(__file__).resolve().parent.parent
    minimal = ('<scene><sensor type="perspective"><film type="hdrfilm"><integer name="width" value="8"/><integer name="height" value="4"/></film></sensor>'
               '<shape type="obj"><string name="filename" value="%s"/><boolean name="filp_tex_coords" value="%s"/></shape></scene>')
    for c in GOLDEN["obj"]:
        if c["text"] is None:
            path = root / c["file"]
        else:
            path = tmp_path / c["file"]
            path.write_text(c["text"])
        with host_api.HostScene(xml=minimal % (path, "true" if c["flip"] else "false")) as hs:
            m = hs.meshes()[0]
        nv, nf, hn, hu = c["counts"]
        assert (m["verts"].shape[0], m["tris"].shape[0], bool(m["has_normals"]), bool(m["has_uvs"])) == (nv, nf, bool(hn), bool(hu)), c["file"]
        assert [int(x) for x in m["tris"].reshape(-1)] == c["faces"], c["file"]
        want = F(c["verts"]).reshape(-1, 8)
        cols = list(range(3)) + (list(range(3, 6)) if hn else []) + (list(range(6, 8)) if hu else [])
        same_bits(np.ascontiguousarray(m["verts"][:, cols]), want[:, cols].reshape(-1).view(np.uint32), f"OBJ vertices {c['file']} flip={c['flip']}")


def test_colour_to_spectrum_plugins():
    """What an <rgb> tag becomes (xml.cpp:269-277): the reference's own srgb.cpp (srgb_model_fetch over ext/rgb2spec and
    the table its optimiser generated), spectra/srgb.cpp, spectra/srgb_d65.cpp (scale = 2 max(rgb), D65 table x scale / 10568)
    and spectra/d65.cpp expanded into spectra/regular.cpp, against the scene builder's spectra evaluated by the oracle --
    the same MskSpectrum records the product's host front-end flattens to (tests/test_host_frontend.py compares those)."""
    from misaki_render_b200.scene import SceneDescription
    sd = SceneDescription(8, 8)
    ids = []
    for c in GOLDEN["colour_spectrum"]:
        rgb, scale = F(c["rgb"]), float(F(c["scale"])[0])
        if c["kind"] == 0:
            ids.append(sd.spectrum_srgb(rgb))
        elif c["kind"] == 1:
            ids.append(sd.spectrum_srgb_d65(rgb, scale))
        else:
            ids.append(sd.spectrum_d65(scale))
    sd.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], f32), np.array([[0, 1, 2]], np.uint32), sd.bsdf_diffuse(0.5))
    osc = po.OracleScene(sd)
    for c, sid in zip(GOLDEN["colour_spectrum"], ids):
        same_bits(osc.spectrum_eval(sid, F(c["wl"])), c["out"], f"colour spectrum kind={c['kind']} rgb={F(c['rgb'])}")


def test_sampling_integrator_render_whole_films():
    """SamplingIntegrator::render / render_block / render_sample THEMSELVES (src/librender/integrator.cpp): the tile loop
    over the spiral, the order of random draws per camera sample (position 2, wavelength 1, aperture 2), the path tracer,
    spectrum_to_xyz, the XYZAW channels, the filtered splat and the merge into the film -- the reference's own code end to
    end, with a serial stand-in for tbb::parallel_for and the camera ray of each sample supplied as an input.  The reference
    never seeds its sampler per pixel (one clone per TBB task), so the oracle replays the film in its test-only
    reference-seeding mode (ORC_RENDER_REFERENCE_SEEDING); every other line is the code the GPU is compared with.
    The whole film must match bit for bit."""
    import hashlib
    from misaki_render_b200 import capi
    from workloads import scenes
    for c in GOLDEN["render"]:
        W, H = c["W"], c["H"]
        sd = {"cbox": lambda: scenes.cbox_uniform(W, H), "open": lambda: scenes.open_uniform(W, H)[0],
              "cbox_rgb": lambda: scenes.cbox(W, H)}[c["scene"]]()  # cbox_rgb: BASELINE config C1's scene itself
        rd = capi.render_desc(spp=c["spp"], max_depth=-1, rr_depth=5)
        rd.flags |= 0x80000000  # ORC_RENDER_REFERENCE_SEEDING (oracle.h)
        film, _ = po.OracleScene(sd).render(rd, nthreads=1)
        same_bits(np.stack([film[y, x] for y, x in c["probes"]]), c["probe_values"], f"render {c['scene']} probe pixels")
        assert hashlib.sha256(np.ascontiguousarray(film, f32).tobytes()).hexdigest() == c["sha256"], f"film {c['scene']} {W}x{H}"
        assert film[..., :3].max() > 0


def test_hdrfilm_image_develop():
    """HDRFilm::image (src/librender/films/hdrfilm.cpp:48-90, compiled with film.cpp over the reference's own ImageBlock
    storage; only the OpenImageIO-backed Image is a buffer stand-in): XYZ -> linear sRGB, / W with the W == 0 guard,
    alpha = A / W, AOV channels / W.  Checked bit for bit against the oracle's develop AND the product's host develop
    (misaki_render_b200/host/imageio.cpp, what GpuPathIntegrator's film runs before the EXR writer)."""
    from misaki_render_b200 import host_api
    for c in GOLDEN["hdrfilm_image"]:
        W, H, nch = c["W"], c["H"], c["nch"]
        film = np.array(c["film"], np.uint32).view(f32).reshape(H, W, nch)
        want = np.array(c["image"], np.uint32).view(f32).reshape(H, W, nch - 1)
        same_bits(host_api.develop_channels(film), bits_of(want), f"host develop_channels {W}x{H}x{nch}")
        same_bits(host_api.develop(film[..., :5]), bits_of(want[..., :4]), f"host develop {W}x{H}")
        same_bits(po.develop(film[..., :5]), bits_of(want[..., :4]), f"oracle develop {W}x{H}")


@pytest.mark.skipif(not po.REF_CODE_LIB.exists(), reason="oracle/_ref is built only where /root/reference exists")
def test_reference_loop_helper_reproduces_its_golden_film_and_runs_threaded():
    """pyoracle.ReferenceLoop (the compiled reference loop behind `bench.py --impl reference --ref-kind reference`): in
    its single-task mode it reproduces the committed golden film of config C1's scene; with several threads (the timing
    mode: every task clones the sampler again, integrator.cpp:57) the film is a valid render of the same scene."""
    import hashlib
    from workloads import scenes
    c = next(c for c in GOLDEN["render"] if c["scene"] == "cbox_rgb")
    sd = scenes.cbox(c["W"], c["H"])
    loop = po.ReferenceLoop(sd, [r for _, r in scenes.CBOX_SHAPES],
                            [(40, 40, 40) if n == "luminaire" else (-1, -1, -1) for n, _ in scenes.CBOX_SHAPES])
    film, _ = loop.render(c["spp"], threads=1)
    assert hashlib.sha256(film.tobytes()).hexdigest() == c["sha256"]
    par, _ = loop.render(c["spp"], threads=4)
    assert np.isfinite(par).all()
    np.testing.assert_allclose(par[..., 4].sum(), film[..., 4].sum(), rtol=1e-2)            # as many samples, other positions
    assert abs(par[..., 1].sum() / film[..., 1].sum() - 1) < 0.25                            # same scene, other random numbers


def test_oracle_converges_to_the_reference_codes_image():
    """SURVEY 8(d)(ii) against the reference's OWN code: tests/golden/ref_cbox48_converged.npz is BASELINE config C1's
    Cornell box rendered by the compiled reference loop at 4096 spp (tools/gen_golden_ref_converged.py).  The oracle, with
    its own per-(pixel, sample) seeds, must approach that image at the Monte-Carlo rate -- relMSE ~ 1 / spp, at 256 spp
    within 10 % of what the reference's own 256-spp render shows against it -- which a biased integrator cannot do."""
    from misaki_render_b200 import capi
    from tests.util import relmse
    from workloads import scenes
    g = np.load(Path(__file__).parent / "golden" / "ref_cbox48_converged.npz")
    osc = po.OracleScene(scenes.cbox(48, 48))
    e = {}
    for spp in (64, 256, 1024):
        film, _ = osc.render(capi.render_desc(spp=spp, max_depth=-1, rr_depth=5))
        e[spp] = relmse(po.develop(film), g["image"])
    assert abs(e[256] / float(g["ref_relmse_256"]) - 1) < 0.10, e
    assert e[1024] < 1.25 * float(g["ref_relmse_1024"]), e
    assert 3.0 < e[64] / e[256] < 5.0 and 3.0 < e[256] / e[1024] < 5.0, e  # ~ 1 / spp down to the fixture's own noise floor


def test_aov_integrator_whole_film():
    """The reference's render loop driving AOVIntegrator (integrator.cpp:36-41 channel list from aov_names(), render_sample's
    aovs + 5, HDRFilm with 21 channels): a whole film of the Cornell box -- XYZAW plus depth, position, uv, geometric /
    shading normal and the nested path tracer's RGBA -- bit for bit against the oracle's AOV render in reference-seeding mode."""
    import hashlib
    from misaki_render_b200 import capi
    from workloads import scenes
    c = GOLDEN["render_aov"]
    rd = capi.render_desc(spp=c["spp"], max_depth=-1, rr_depth=5)
    rd.flags |= 0x80000000  # ORC_RENDER_REFERENCE_SEEDING (oracle.h)
    types = [capi.AOV_DEPTH, capi.AOV_POSITION, capi.AOV_UV, capi.AOV_GEO_NORMAL, capi.AOV_SH_NORMAL, capi.AOV_INTEGRATOR_RGBA]
    film, _ = po.OracleScene(scenes.cbox_uniform(c["W"], c["H"])).render_aov(rd, types, nthreads=1)
    assert film.shape == (c["H"], c["W"], 21)
    same_bits(np.stack([film[y, x] for y, x in c["probes"]]), c["probe_values"], "AOV film probe pixels")
    assert hashlib.sha256(np.ascontiguousarray(film, f32).tobytes()).hexdigest() == c["sha256"]
