"""Diagnostic (GPU): TOI events per world-step on the settled jittered pyramid (sum of b2Contact::m_toiCount)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2dpp_b200 import scene as S
from box2dpp_b200.batch import WorldBatch, F_CONTACT_A, F_CONTACT_COUNT

W = 1024
flags = S.WORLD_DEFAULT & ~S.WORLD_ALLOW_SLEEP
sb = S.pyramid_batch(W, 20, flags=flags, first_seed=0)
b = WorldBatch(sb)
b.step_n(1 / 60, 8, 3, 90)
hist = np.zeros(64, np.int64)
for s in range(30):
    b.step(1 / 60, 8, 3)
    ca = b.field(F_CONTACT_A)           # [W][NC][4]
    n = b.field(F_CONTACT_COUNT)[:, 0, 0]
    fl = ca[:, :, 2]
    cnt = (fl >> 8) & 0xf
    mask = np.arange(fl.shape[1])[None, :] < n[:, None]
    per_world = (cnt * mask).sum(1)
    for v in per_world:
        hist[min(v, 63)] += 1
    if s % 10 == 0:
        print("step", 90 + s, "events/world mean", per_world.mean(), "max", per_world.max(), "worlds with events", (per_world > 0).mean())
print("hist", hist[:20])
