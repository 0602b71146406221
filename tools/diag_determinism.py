"""Diagnostic: do repeated runs of the same 16,384 pyramid worlds agree bit for bit?

    RUNS=12 python tools/diag_determinism.py W steps mode [generic]     mode: 1 = one b2gpu_batch_step per step, n = step_n

One batch, its initial state saved once (b2gpu_batch_snapshot_save); every run restores it, steps and downloads the body
sweeps.  Prints, per run that differs from run 0, how many worlds differ and in which tiles / lanes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S
from box2dpp_b200.batch import WorldBatch

W = int(sys.argv[1]); steps = int(sys.argv[2]); mode = sys.argv[3]; generic = len(sys.argv) > 4
runs = int(os.environ.get("RUNS", "3"))
flags = S.WORLD_DEFAULT & ~S.WORLD_CONTINUOUS & ~S.WORLD_ALLOW_SLEEP
sb = S.pyramid_batch(W, 20, flags=flags, first_seed=0, jittered=True)
b = WorldBatch(sb, generic_solver=generic)
b.step(1 / 60, 8, 3)  # the initial FindNewContacts is behind us
b.sync()
start = b.snapshot()
tag = f"groups={os.environ.get('B2GPU_GROUPS', 'auto')} mode={mode} generic={generic}"
ref = None
bad = 0
t0 = time.time()
for r in range(runs):
    b.restore(start)
    if mode == "n":
        b.step_n(1 / 60, 8, 3, steps)
    else:
        for _ in range(steps):
            b.step(1 / 60, 8, 3)
    b.sync()
    out = b.field(B.F_BODY_SWEEP)
    if ref is None:
        ref = out.copy()
        continue
    d = np.nonzero((ref.view(np.uint32) != out.view(np.uint32)).any(axis=(1, 2)))[0]
    if len(d):
        bad += 1
        for tile in sorted(set(int(x) // 32 for x in d))[:6]:
            lanes = sorted(int(x) % 32 for x in d if int(x) // 32 == tile)
            w = tile * 32 + lanes[0]
            print(f"   {tag} run {r}: tile {tile} lanes {lanes} world {w}: max |diff| {float(np.abs(ref[w] - out[w]).max()):.3g}")
print(f"{tag}: {bad} of {runs - 1} runs differ from run 0 ({steps} steps, {time.time() - t0:.0f} s)")
b.close()
