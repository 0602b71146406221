"""Diagnostic (GPU): ms per step of the pyramid batch under the current environment (B2GPU_LIB, B2GPU_GROUPS,
B2GPU_PIPELINE, B2GPU_ISLAND_KB); one line of output.  Used by tools/pipe_exp.sh to compare group / pipeline settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from box2dpp_b200 import scene as S
from box2dpp_b200 import batch as B

W = int(os.environ.get("WORLDS", "16384"))
flags = S.WORLD_DEFAULT & ~S.WORLD_ALLOW_SLEEP & ~S.WORLD_CONTINUOUS
b = B.WorldBatch(S.pyramid_batch(W, 20, flags=flags, first_seed=0))
b.step_n(1 / 60, 8, 3, 296)
b.sync()
t0 = time.perf_counter()
b.step_n(1 / 60, 8, 3, 200)
b.sync()
t1 = time.perf_counter()
keys = ("B2GPU_LIB", "B2GPU_GROUPS", "B2GPU_PIPELINE", "B2GPU_ISLAND_KB")
print(" ".join(f"{k[6:]}={os.environ.get(k, '-')}" for k in keys), f"worlds={W} ms_per_step={(t1 - t0) * 1e3 / 200:.3f}",
      f"status_or={int(b.status().max())}", flush=True)
