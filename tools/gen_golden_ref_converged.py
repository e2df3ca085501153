"""Converged image of BASELINE config C1's scene from the reference's OWN compiled render loop (pyoracle.ReferenceLoop over
oracle/_ref, built by oracle/Makefile.ref from /root/reference): the Cornell box at 48x48, 4096 samples per pixel, developed
to linear sRGB.  The camera rays come from the reference's own sensors/perspective.cpp.  Committed as tests/golden/ref_cbox48_converged.npz so that the oracle (CPU tests) and the GPU path (-m gpu
tests) are compared with the reference's code directly -- statistically, SURVEY 8(d)(ii), because the reference never seeds
per pixel.  Also stores the reference's own relMSE at 256 spp against that image, the yardstick for "same noise level".
Run here (needs /root/reference for the build of oracle/_ref): python tools/gen_golden_ref_converged.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402
from workloads import scenes  # noqa: E402

W = H = 48


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def main():
    sd = scenes.cbox(W, H)
    loop = po.ReferenceLoop(sd, [r for _, r in scenes.CBOX_SHAPES],
                            [(40, 40, 40) if n == "luminaire" else (-1, -1, -1) for n, _ in scenes.CBOX_SHAPES],
                            camera="reference")  # the reference's own PerspectiveCamera supplies the rays (oracle/ref_camera_wrap.cpp)
    film, secs = loop.render(4096, threads=8)
    image = po.develop(film)[..., :3]
    own = {}
    for spp in (256, 1024):
        f, _ = loop.render(spp, threads=8)
        own[spp] = relmse(po.develop(f)[..., :3], image)
    out = ROOT / "tests" / "golden" / "ref_cbox48_converged.npz"
    np.savez_compressed(out, image=image.astype(np.float32), spp=np.int32(4096),
                        ref_relmse_256=np.float64(own[256]), ref_relmse_1024=np.float64(own[1024]))
    print(f"wrote {out} ({out.stat().st_size} bytes; {secs:.1f} s of reference code); the reference's own relMSE against it: {own}")


if __name__ == "__main__":
    main()
