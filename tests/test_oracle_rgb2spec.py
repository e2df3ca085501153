"""rgb2spec is the one part of the reference that compiles in this image, so the oracle's restatement
(oracle.cpp orc_rgb2spec_fetch, srgb_model_eval) and the product's host-side restatement
(misaki_render_b200/rgb2spec.py) are PINNED against it: bit-exact coefficients from the committed golden
vectors (tests/golden/rgb2spec_fetch.json, made by tools/gen_golden_rgb2spec.py) and, when oracle/_ref is
present, against the live reference library."""
import json
from pathlib import Path

import numpy as np
import pytest

from misaki_render_b200 import rgb2spec
from oracle import pyoracle

GOLDEN = json.loads((Path(__file__).parent / "golden" / "rgb2spec_fetch.json").read_text())


def _bits(a):
    return [np.float32(x).tobytes().hex() for x in a]


def test_oracle_fetch_matches_reference_golden():
    m = pyoracle.Rgb2Spec(rgb2spec.DEFAULT_COEFF)
    for case in GOLDEN["cases"]:
        assert _bits(m.fetch(case["rgb"])) == case["coeff_hex"], case["rgb"]


def test_product_host_fetch_matches_reference_golden():
    m = rgb2spec.model()
    for case in GOLDEN["cases"]:
        assert _bits(m.fetch(case["rgb"])) == case["coeff_hex"], case["rgb"]


def test_srgb_model_eval_matches_reference_eval_precise():
    """srgb.h:8-19 evaluates the same sigmoid-of-polynomial as rgb2spec_eval_precise (rgb2spec.c:121-135)."""
    wl = GOLDEN["wavelengths"]
    for case in GOLDEN["cases"]:
        if case["eval"][0] is None:
            continue
        for k in range(0, len(wl) - 3):
            got = pyoracle.srgb_model_eval(case["coeff"], wl[k:k + 4])
            np.testing.assert_allclose(got, case["eval"][k:k + 4], rtol=0, atol=2e-6)


def test_gray_upsamples_to_gray():
    c = rgb2spec.model().fetch((0.5, 0.5, 0.5))
    v = pyoracle.srgb_model_eval(c, [400, 500, 600, 700])
    np.testing.assert_allclose(v, 0.5, atol=2e-3)
    c = rgb2spec.model().fetch((1, 1, 1))
    assert (pyoracle.srgb_model_eval(c, [400, 500, 600, 700]) > 0.95).all()  # the fitted white dips to ~0.96 near 600 nm


@pytest.mark.skipif(not pyoracle.REF_LIB.exists(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_library_random_colours():
    ref, orc, host = pyoracle.RefRgb2Spec(), pyoracle.Rgb2Spec(), rgb2spec.model()
    rng = np.random.default_rng(11)
    for _ in range(300):
        c = rng.random(3).astype(np.float32) * rng.choice([1.0, 1.0, 0.1])
        r = ref.fetch(c)
        assert _bits(orc.fetch(c)) == _bits(r)
        assert _bits(host.fetch(c)) == _bits(r)
