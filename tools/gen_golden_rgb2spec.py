#!/usr/bin/env python3
"""Generates tests/golden/rgb2spec_fetch.json by running the UNMODIFIED reference ext/rgb2spec
(compiled into oracle/_ref by oracle/Makefile.ref) in this container: rgb2spec_fetch() coefficients and
rgb2spec_eval_precise() values for a fixed list of colours.  The fixture travels to the GPU box, where
/root/reference does not exist."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from oracle import pyoracle

ref = pyoracle.RefRgb2Spec()
rng = np.random.default_rng(7)
colors = [(0.5, 0.5, 0.5), (1, 1, 1), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0.936461, 0.740433, 0.705267),
          (0.885809, 0.698859, 0.666422), (0.105421, 0.37798, 0.076425), (0.570068, 0.0430135, 0.0443706), (0.45, 0.30, 0.90),
          (0.5, 0.5, 0.5 + 1e-7), (0.2, 0.2, 0.2), (1.5, -0.2, 0.3)] + [tuple(float(x) for x in rng.random(3)) for _ in range(50)]
wl = [360.0, 400.0, 455.5, 538.0, 600.25, 700.0, 830.0]
out = []
for c in colors:
    coeff = ref.fetch(c)
    out.append(dict(rgb=[float(np.float32(x)) for x in c], coeff=[float(x) for x in coeff],
                    coeff_hex=[np.float32(x).tobytes().hex() for x in coeff],
                    eval=[None if not np.isfinite(coeff).all() else float(ref.eval_precise(coeff, l)) for l in wl]))
(ROOT / "tests" / "golden").mkdir(exist_ok=True)
(ROOT / "tests" / "golden" / "rgb2spec_fetch.json").write_text(json.dumps(dict(
    source="reference ext/rgb2spec/rgb2spec.c (rgb2spec_fetch, rgb2spec_eval_precise) + srgb.coeff from `rgb2spec_opt 64`",
    wavelengths=wl, cases=out), indent=1))
print("wrote", len(out), "cases")
