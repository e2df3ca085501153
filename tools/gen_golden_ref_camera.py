"""Golden camera rays from the reference's OWN compiled camera (oracle/ref_camera_wrap.cpp over include/misaki/core/
transform.h, src/librender/sensor.cpp and src/librender/sensors/perspective.cpp; oracle/Makefile.ref): for the cameras of
the BASELINE configs (C1/C4 Cornell box, C2 bunny, C3 teapot) and a non-square film, the matrices perspective.cpp:12-19
builds, Transform4f::lookat / translate / scale / rotate, and sample_ray() for a fixed set of (wavelength sample, pixel
position) samples.  Committed as tests/golden/ref_camera.json; tests/test_oracle_ref_camera.py holds the oracle's camera,
the scene builder's and the host front-end's PerspectiveCamera to it (1e-6 relative: the Eigen stand-in's 4x4 inverse
orders its float operations differently from Eigen's SSE routine, so this pin is not a bit pin).
Run here (needs /root/reference for the build of oracle/_ref): python tools/gen_golden_ref_camera.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402

CAMERAS = {  # name: W, H, fov, near, far, lookat(origin, target, up) -- workloads/scenes.py
    "c1_cbox_256": (256, 256, 49.3077, 10.0, 2800.0, ((278, 273, -800), (278, 273, -799), (0, 1, 0))),
    "c4_cbox_1920x1080": (1920, 1080, 49.3077, 10.0, 2800.0, ((278, 273, -800), (278, 273, -799), (0, 1, 0))),
    "c2_bunny_512": (512, 512, 35.0, 0.1, 100.0, ((0.0, 2.2, -4.5), (0.0, 0.9, 0.0), (0, 1, 0))),
    "c3_teapot_1024": (1024, 1024, 35.0, 0.1, 100.0, ((0.0, 2.4, -4.8), (0.0, 0.8, 0.0), (0, 1, 0))),
    "oblique_80x48": (80, 48, 61.5, 0.05, 37.0, ((1.5, -2.25, 3.0), (-0.25, 0.5, 0.125), (0.2, 0.9, -0.1))),
}


def main():
    rng = np.random.default_rng(20261018)
    out = {"cameras": {}, "transforms": []}
    for name, (W, H, fov, near, far, (o, t, u)) in CAMERAS.items():
        tw = po.ref_lookat(o, t, u)
        cam = po.ReferenceCamera(W, H, fov, near, far, tw)
        c2s, s2c = cam.matrices(W, H, fov, near, far)
        n = 48
        s = np.empty((n, 3), np.float32)
        s[:, 0] = rng.random(n, dtype=np.float32)
        s[:, 1] = rng.random(n, dtype=np.float32) * W
        s[:, 2] = rng.random(n, dtype=np.float32) * H
        s[:4, 1:] = [[0, 0], [W, H], [W / 2, H / 2], [0.5, H - 0.5]]  # corners, centre, a pixel centre
        rays = cam.sample_rays(s)
        cam.close()
        out["cameras"][name] = {"width": W, "height": H, "fov": fov, "near_clip": near, "far_clip": far, "lookat": [list(map(float, o)), list(map(float, t)), list(map(float, u))],
                                "to_world": tw.astype(float).tolist(), "camera_to_sample": c2s.astype(float).tolist(), "sample_to_camera": s2c.astype(float).tolist(),
                                "samples": s.astype(float).tolist(), "rays": rays.astype(float).tolist()}
    for kind, v, angle in (("translate", (1.5, -2.0, 0.25), 0.0), ("scale", (2.0, 0.5, -3.0), 0.0), ("rotate", (0.0, 1.0, 0.0), 0.6108652), ("rotate", (0.267261, 0.534522, 0.801784), -1.1)):
        m, inv = po.ref_transform(kind, v, angle)
        out["transforms"].append({"kind": kind, "v": list(map(float, v)), "angle": angle, "matrix": m.astype(float).tolist(), "inverse": inv.astype(float).tolist()})
    p = ROOT / "tests" / "golden" / "ref_camera.json"
    p.write_text(json.dumps(out, separators=(",", ":")))
    print(f"wrote {p} ({p.stat().st_size} bytes): {len(out['cameras'])} cameras x 48 rays, {len(out['transforms'])} transforms")


if __name__ == "__main__":
    main()
