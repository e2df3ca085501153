#!/usr/bin/env python3
"""Writes assets/cbox/meshes/cbox_*.obj from the public Cornell box measurements
(https://www.graphics.cornell.edu/online/box/data.html), one OBJ per shape named as
assets/cbox/scene.xml of the reference expects.  Quads are wound so that the geometric
normal (p1-p0)x(p2-p0) of every room surface faces into the room, of every block face
faces outwards and of the luminaire faces down: the reference's diffuse BSDF and area
light are one-sided (diffuse.cpp:23, area.cpp:39,52)."""
from pathlib import Path
import numpy as np

ROOM_CENTER = np.array([278.0, 274.4, 279.6])

def quads():
    q = {}
    q["floor"] = [[(552.8, 0, 0), (0, 0, 0), (0, 0, 559.2), (549.6, 0, 559.2)]]
    q["ceiling"] = [[(556, 548.8, 0), (556, 548.8, 559.2), (0, 548.8, 559.2), (0, 548.8, 0)]]
    q["back"] = [[(549.6, 0, 559.2), (0, 0, 559.2), (0, 548.8, 559.2), (556, 548.8, 559.2)]]
    q["greenwall"] = [[(0, 0, 559.2), (0, 0, 0), (0, 548.8, 0), (0, 548.8, 559.2)]]
    q["redwall"] = [[(552.8, 0, 0), (549.6, 0, 559.2), (556, 548.8, 559.2), (556, 548.8, 0)]]
    q["luminaire"] = [[(343, 548.8, 227), (343, 548.8, 332), (213, 548.8, 332), (213, 548.8, 227)]]
    q["smallbox"] = [
        [(130, 165, 65), (82, 165, 225), (240, 165, 272), (290, 165, 114)],
        [(290, 0, 114), (290, 165, 114), (240, 165, 272), (240, 0, 272)],
        [(130, 0, 65), (130, 165, 65), (290, 165, 114), (290, 0, 114)],
        [(82, 0, 225), (82, 165, 225), (130, 165, 65), (130, 0, 65)],
        [(240, 0, 272), (240, 165, 272), (82, 165, 225), (82, 0, 225)],
    ]
    q["largebox"] = [
        [(423, 330, 247), (265, 330, 296), (314, 330, 456), (472, 330, 406)],
        [(423, 0, 247), (423, 330, 247), (472, 330, 406), (472, 0, 406)],
        [(472, 0, 406), (472, 330, 406), (314, 330, 456), (314, 0, 456)],
        [(314, 0, 456), (314, 330, 456), (265, 330, 296), (265, 0, 296)],
        [(265, 0, 296), (265, 330, 296), (423, 330, 247), (423, 0, 247)],
    ]
    return q

def oriented(name, quad):
    p = np.array(quad, dtype=np.float64)
    n = np.cross(p[1] - p[0], p[2] - p[0])
    c = p.mean(axis=0)
    if name in ("smallbox", "largebox"):
        box_c = np.array([q for f in quads()[name] for q in f], dtype=np.float64).mean(axis=0)
        want = c - box_c
    elif name == "luminaire":
        want = np.array([0.0, -1.0, 0.0])
    else:
        want = ROOM_CENTER - c
    return quad if np.dot(n, want) > 0 else quad[::-1]

def main():
    out = Path(__file__).resolve().parent.parent / "assets" / "cbox" / "meshes"
    out.mkdir(parents=True, exist_ok=True)
    for name, faces in quads().items():
        lines = [f"# Cornell box: {name} (public Cornell data), written by workloads/make_cbox.py"]
        nv = 0
        for quad in faces:
            quad = oriented(name, quad)
            for v in quad:
                lines.append("v %g %g %g" % v)
            lines.append("f %d %d %d %d" % (nv + 1, nv + 2, nv + 3, nv + 4))
            nv += 4
        (out / f"cbox_{name}.obj").write_text("\n".join(lines) + "\n")
    print("wrote", out)

if __name__ == "__main__":
    main()
