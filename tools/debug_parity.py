#!/usr/bin/env python
"""Step one scene on the device and on the reference side by side; report the first step that differs.
Usage: python tools/debug_parity.py mixed 0 [steps] [vi] [pi]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from box2dpp_b200 import scene as S  # noqa: E402
from box2dpp_b200.batch import EVENT_ALL, WorldBatch  # noqa: E402
from oracle import worlds  # noqa: E402
from parity_util import compare_world  # noqa: E402

kind, seed = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
vi = int(sys.argv[4]) if len(sys.argv) > 4 else 8
pi = int(sys.argv[5]) if len(sys.argv) > 5 else 3
sc = {"mixed": lambda: S.mixed(seed), "bullets": lambda: S.bullets(seed), "pyramid": lambda: S.pyramid(seed),
      "hello": lambda: S.hello_world()}[kind]()
ref = worlds.RefWorld(sc, record_events=True)
batch = WorldBatch([sc], event_mask=EVENT_ALL, always_rebuild_islands=bool(int(os.environ.get("REBUILD", "0"))))
np.set_printoptions(linewidth=200, precision=9)
for step in range(steps):
    ref.step(1 / 60, vi, pi)
    batch.step(1 / 60, vi, pi)
    ri = ref.events()[0]
    if len(ri) and ri[-1, 0] != 3 and (ri[:, 0] == 3).any():
        print("step", step, "TOI-phase events:", "".join(str(t) for t in ri[np.argwhere(ri[:, 0] == 3)[-1, 0] + 1:, 0]))
    try:
        compare_world(ref, batch.bodies(0), batch.contacts(0), batch.proxies(0), batch.events(0), what=f"step {step}")
    except AssertionError as e:
        print("FIRST MISMATCH:", str(e)[:600])
        re_i, re_f = ref.events()
        ge_i, ge_f = batch.events(0)
        print("ref events", len(re_i), "gpu events", len(ge_i))
        n = min(len(re_i), len(ge_i))
        bad = [i for i in range(n) if (re_i[i] != ge_i[i]).any() or (re_f[i].view(np.uint32) != ge_f[i].view(np.uint32)).any()]
        print("first differing event index:", bad[:5])
        for i in bad[:3]:
            print("  ref", re_i[i], re_f[i], "\n  gpu", ge_i[i], ge_f[i])
        # events past the discrete phase (types 3 end the Solve; anything after the last run of 3s is TOI)
        print("ref event types:", "".join(str(t) for t in re_i[:, 0]))
        print("gpu event types:", "".join(str(t) for t in ge_i[:, 0]))
        rb, gb = ref.bodies(), batch.bodies(0)
        for c in range(11):
            d = np.argwhere(rb[:, c].view(np.uint32) != gb[:, c].view(np.uint32))[:, 0]
            if len(d):
                print("col", c, "bodies", d[:8].tolist(), "ref", rb[d[:4], c], "gpu", gb[d[:4], c])
        break
else:
    print("no mismatch in", steps, "steps")
