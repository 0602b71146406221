"""Diagnostic: two batches of the same worlds stepped in lockstep, compared after every step.

    python tools/diag_lockstep.py W steps [mode]     mode: 1 = b2gpu_batch_step per step, n = step_n(1) (graph replay)

Reports the FIRST step at which the two disagree, which persistent fields disagree there (the states were identical one
step earlier, so the field class names the kernel) and where (tile, lanes, item indices); then copies batch A's state
into B and goes on, so one run collects several events.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S
from box2dpp_b200.batch import WorldBatch

W = int(sys.argv[1]); steps = int(sys.argv[2]); mode = sys.argv[3] if len(sys.argv) > 3 else "1"
flags = S.WORLD_DEFAULT & ~S.WORLD_CONTINUOUS & ~S.WORLD_ALLOW_SLEEP
sb = S.pyramid_batch(W, 20, flags=flags, first_seed=0, jittered=True)
a, b = WorldBatch(sb), WorldBatch(sb)
CHEAP = [("BODY_VEL", B.F_BODY_VEL), ("BODY_SWEEP", B.F_BODY_SWEEP)]
MORE = [("CONTACT_P0", B.F_CONTACT_P0), ("CONTACT_P1", B.F_CONTACT_P1), ("CONTACT_N", B.F_CONTACT_N), ("CONTACT_A", B.F_CONTACT_A),
        ("CONTACT_K", B.F_CONTACT_K), ("BODY_XF", B.F_BODY_XF), ("PROXY_FAT", B.F_PROXY_FAT), ("ISLAND_CONTACTS", B.F_ISLAND_CONTACTS),
        ("CONTACT_COUNT", B.F_CONTACT_COUNT), ("BODY_FLAGS", B.F_BODY_FLAGS)]


def one(x):
    if mode == "n":
        x.step_n(1 / 60, 8, 3, 1)
    else:
        x.step(1 / 60, 8, 3)


def diff(name, fid):
    fa, fb = a.field(fid), b.field(fid)
    ne = fa.view(np.uint32) != fb.view(np.uint32)
    if not ne.any():
        return None
    worlds = np.nonzero(ne.any(axis=(1, 2)))[0]
    w = int(worlds[0])
    items = np.nonzero(ne[w].any(axis=1))[0]
    comps = sorted(set(int(c) for c in np.nonzero(ne[worlds].any(axis=(0, 1)))[0]))
    fl = fa[w].astype(np.float64) - fb[w].astype(np.float64) if fa.dtype == np.float32 else None
    return (f"{name}: {len(worlds)} worlds, tiles {sorted(set(int(x) // 32 for x in worlds))[:8]}, lanes "
            f"{sorted(set(int(x) % 32 for x in worlds))}, world {w}: {len(items)} items {items[:12].tolist()} comps {comps}"
            + (f" max|d| {np.abs(fl).max():.3g}" if fl is not None else ""))


events = 0
for t in range(steps):
    one(a); one(b)
    a.sync(); b.sync()
    bad = [d for d in (diff(n, f) for n, f in CHEAP) if d]
    if t >= 30 and not bad and t % 10 == 0:  # impulses too, now and then
        bad = [d for d in (diff(n, f) for n, f in MORE[:2]) if d]
    if bad:
        events += 1
        print(f"step {t}: FIRST DIFFERENCE (event {events})")
        for d in bad:
            print("   ", d)
        for n, f in MORE:
            d = diff(n, f)
            print("   ", d if d else f"{n}: equal")
        b.restore(a.snapshot())
        bad2 = [d for d in (diff(n, f) for n, f in CHEAP) if d]
        print("    after restore:", "equal" if not bad2 else bad2)
        sys.stdout.flush()
print(f"lockstep W={W} steps={steps} mode={mode} groups={os.environ.get('B2GPU_GROUPS', 'auto')}: {events} events")
