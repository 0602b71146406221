"""Diagnostic (GPU): where k_toi_events spends its cycles.  Needs the variant built by
`tools/build_variant.sh prof -DB2G_TOI_PROFILE` and B2GPU_LIB=box2dpp_b200/libb2gpu_prof.so."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2dpp_b200 import scene as S
from box2dpp_b200 import batch as B

W = 16384
flags = S.WORLD_DEFAULT & ~S.WORLD_ALLOW_SLEEP
b = B.WorldBatch(S.pyramid_batch(W, 20, flags=flags, first_seed=0))
b.step_n(1 / 60, 8, 3, 100)
b.sync()
out = (C.c_ulonglong * 16)()
B.lib().b2gpu_debug_toi_profile(out, 1)
N = 20
b.step_n(1 / 60, 8, 3, N)
b.sync()
B.lib().b2gpu_debug_toi_profile(out, 0)
names = ["find min (+TOI recompute)", "advance + update contact", "build island", "SolveTOI", "sync + invalidate", "UpdatePairs", "  ContactToi inside find-min", "  invalidate walk inside sync"]
tiles = (W + 31) // 32
for i, nm in enumerate(names):
    cyc, cnt = out[i], out[8 + i]
    print(f"{nm:28s} {cyc / max(cnt,1):12.0f} cycles/visit  {cnt / (tiles * N):6.2f} visits/tile-step  {cyc / (tiles * N) / 1965e3:8.3f} ms/tile-step")
