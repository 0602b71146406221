#!/usr/bin/env python
"""Generate tests/golden/* from the reference's own vendored vanilla Box2D 2.3.2.

Run in the build container (needs /root/reference, i.e. oracle/_ref/libb2ref.so built by
`make -C oracle ref`).  The fixtures travel with the repo, so the GPU box and CI can check the C
restatement (oracle/b2o_*.c), the host layer and the CUDA path without the reference present.

The reference ships no golden files: its tests run the vanilla library beside the fork.
We do the same, once, and freeze the answers here:

  narrowphase.npz   inputs + manifolds for the reference's known-answer tests
                    (Regression/tests/helloworld.cpp:268-418) and seeded random shape pairs, for all five
                    collide functions (b2CollidePolygons, b2CollideCircles, b2CollidePolygonAndCircle,
                    b2CollideEdgeAndCircle, b2CollideEdgeAndPolygon)
  polygons.npz      b2PolygonShape::Set / SetAsBox / ComputeMass on seeded random point clouds
  sincos.npz        libm sinf / cosf bit patterns (glibc 2.39, the reference's b2Rot::Set dependency)
  worlds.json       per-step FNV-1a hashes (SURVEY.md Appendix D definition) of body state and contact
                    order for every parity scene, plus HelloWorld positions with the reference defaults
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from box2dpp_b200 import scene as S  # noqa: E402
from oracle import worlds  # noqa: E402
import golden_cases as G  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    assert worlds.available("ref"), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(OUT, exist_ok=True)

    # ---- narrow phase ----
    npz = {}
    for name, (fn, gen) in G.NARROW_CASES.items():
        cases = gen()
        ints = np.zeros((len(cases), 4), np.int32)
        fl = np.zeros((len(cases), 8), np.float32)
        for i, (a, xa, b, xb) in enumerate(cases):
            ints[i], fl[i] = getattr(worlds, fn)(a, xa, b, xb, kind="ref")
        npz[name + "_ints"] = ints
        npz[name + "_floats"] = fl
    np.savez_compressed(os.path.join(OUT, "narrowphase.npz"), **npz)

    # ---- polygons ----
    counts, outs = [], []
    for fd, verts in G.polygon_cases():
        n, out = worlds.polygon_build(fd, verts, kind="ref")
        counts.append(n)
        outs.append(out)
    np.savez_compressed(os.path.join(OUT, "polygons.npz"), counts=np.array(counts, np.int32), outs=np.array(outs))

    # ---- sinf / cosf ----
    x = G.sincos_inputs()
    s, c = worlds.sincosf(x)
    np.savez_compressed(os.path.join(OUT, "sincos.npz"), s=s.view(np.uint32), c=c.view(np.uint32))

    # ---- worlds ----
    out = {}
    for name, (make, steps, vi, pi) in G.WORLD_CASES.items():
        sc = make()
        w = worlds.RefWorld(sc)
        fb = G.fixture_bodies(sc)
        sh, ch, counts = [], [], []
        for _ in range(steps):
            w.step(1.0 / 60.0, vi, pi)
            ints, _ = w.contacts()
            sh.append("%016x" % G.state_hash(w.bodies()))
            ch.append("%016x" % G.contact_order_hash(ints, fb))
            counts.append([int(len(ints)), int((ints[:, 4] & 1).sum()), int(ints[(ints[:, 4] & 1) != 0, 5].sum())])
        out[name] = {"steps": steps, "vel_iters": vi, "pos_iters": pi, "state_hash": sh, "contact_order_hash": ch,
                     "contacts_touching_points": counts}
    # HelloWorld with the reference's defaults (continuous physics ON): the un-flagged reference answer
    w = worlds.RefWorld(S.hello_world())
    rows = []
    for _ in range(60):
        w.step(1.0 / 60.0, 6, 2)
        b = w.bodies()[1]
        rows.append([int(np.float32(v).view(np.uint32)) for v in (b[0], b[1], b[6])])
    out["hello_world_default_flags_xya_bits"] = rows
    json.dump(out, open(os.path.join(OUT, "worlds.json"), "w"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
