"""Diagnostic (GPU): timeline of the world groups for a few steps (B2GPU_GROUPS=2 python tools/groups_trace.py)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2dpp_b200 import scene as S
from box2dpp_b200 import batch as B

W = int(os.environ.get("WORLDS", "16384"))
flags = S.WORLD_DEFAULT & ~S.WORLD_ALLOW_SLEEP & ~S.WORLD_CONTINUOUS
b = B.WorldBatch(S.pyramid_batch(W, 20, flags=flags, first_seed=0))
b.step_n(1 / 60, 8, 3, 300)
b.sync()
L = B.lib()
L.b2gpu_batch_trace_groups.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
steps = 4
ms = np.zeros(8 * steps * 12, np.float32)
G, nk = C.c_int32(), C.c_int32()
rc = L.b2gpu_batch_trace_groups(b.h, 1 / 60, 8, 3, steps, ms.ctypes.data_as(C.c_void_p), ms.shape[0], C.byref(G), C.byref(nk))
assert rc == 0, L.b2gpu_last_error()
G, nk = G.value, nk.value
t = ms[:G * steps * (nk + 1)].reshape(G, steps, nk + 1)
names = [L.b2gpu_batch_kernel_name(b.h, i).decode().replace("k_solve_", "").replace("k_", "") for i in range(nk)]
for g in range(G):
    print(f"group {g}:")
    for s in range(steps):
        row = t[g, s]
        print(f"  step {s}: start {row[0]:7.3f} | " + " ".join(f"{names[k]}:{row[k + 1] - row[k]:.3f}" for k in range(nk)) + f" | end {row[-1]:7.3f}")
print("total", t[:, -1, -1].max(), "ms for", steps, "steps")
