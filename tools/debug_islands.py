#!/usr/bin/env python
"""Validate the topology-dirty tracking: step the same scene with and without `always_rebuild_islands` and report
the first step whose island contact order differs.  Usage: python tools/debug_islands.py mixed 0 [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2dpp_b200 import scene as S  # noqa: E402
from box2dpp_b200 import batch as B  # noqa: E402
from box2dpp_b200.batch import EVENT_ALL, WorldBatch  # noqa: E402

kind, seed = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
sc = {"mixed": lambda: S.mixed(seed), "bullets": lambda: S.bullets(seed), "pyramid": lambda: S.pyramid(seed)}[kind]()
a = WorldBatch([sc], event_mask=EVENT_ALL)
b = WorldBatch([sc], event_mask=EVENT_ALL, always_rebuild_islands=True)
for step in range(steps):
    a.step(1 / 60, 8, 3)
    b.step(1 / 60, 8, 3)
    ia, ib = a.field(B.F_ISLAND_CONTACTS)[0, :, 0], b.field(B.F_ISLAND_CONTACTS)[0, :, 0]
    ea, eb = a.field(B.F_ISLAND_CONTACT_END)[0, :, 0], b.field(B.F_ISLAND_CONTACT_END)[0, :, 0]
    na, nb = int(a.field(B.F_ISLAND_COUNT)[0, 0, 0]), int(b.field(B.F_ISLAND_COUNT)[0, 0, 0])
    ev = b.events(0)[0]
    toi = ""
    if len(ev) and (ev[:, 0] == 3).any() and ev[-1, 0] != 3:
        toi = "".join(str(t) for t in ev[np.argwhere(ev[:, 0] == 3)[-1, 0] + 1:, 0])
    ca = ia[:ea[na - 1]] if na else ia[:0]
    cb = ib[:eb[nb - 1]] if nb else ib[:0]
    same = na == nb and len(ca) == len(cb) and (ca == cb).all() and (ea[:na] == eb[:nb]).all()
    if toi or not same:
        print("step", step, "islands", na, nb, "contacts", len(ca), len(cb), "same" if same else "DIFFERENT", "toi events:", toi)
    if not same:
        print(" stale  :", ca.tolist(), ea[:na].tolist())
        print(" rebuilt:", cb.tolist(), eb[:nb].tolist())
        fa, fb = a.field(B.F_BODY_FLAGS)[0, :, 0], b.field(B.F_BODY_FLAGS)[0, :, 0]
        print(" body flags differ at", np.argwhere(fa != fb)[:, 0].tolist(), fa[fa != fb], fb[fa != fb])
        A, Bc = a.field(B.F_CONTACT_A)[0], b.field(B.F_CONTACT_A)[0]
        n = int(a.field(B.F_CONTACT_COUNT)[0, 0, 0])
        print(" contact count", n, int(b.field(B.F_CONTACT_COUNT)[0, 0, 0]), "A differs at", np.argwhere((A[:n] != Bc[:n]).any(axis=1))[:, 0].tolist())
        break
