"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sharding + timing reductions bench.py uses."""
import os
import subprocess
import sys
import textwrap

from box2dpp_b200.shard import owner_of, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_batch():
    for total in (0, 1, 7, 64, 65536, 131072, 1000003):
        for ws in (1, 2, 3, 4, 8):
            cover = 0
            for r in range(ws):
                first, count = shard_range(total, ws, r)
                assert first == cover
                cover += count
                if count:
                    assert owner_of(first, total, ws) == r and owner_of(first + count - 1, total, ws) == r
            assert cover == total
            sizes = [shard_range(total, ws, r)[1] for r in range(ws)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_run(tmp_path):
    """Two processes over gloo: each builds the scenes of its shard (seeded by global world index), the union of
    the shards reproduces the single-process batch, and the time reduction is the max over ranks."""
    script = textwrap.dedent("""
        import os, sys, hashlib
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from box2dpp_b200 import scene as S
        from box2dpp_b200.shard import shard_range, reduce_max, reduce_sum
        dist.init_process_group("gloo")
        rank, ws = dist.get_rank(), dist.get_world_size()
        total = 37
        first, count = shard_range(total, ws, rank)
        sb = S.pyramid_batch(count, 3, first_seed=first)
        h = hashlib.sha256()
        for i in range(count):
            sc = sb.scene(i)
            h.update(np.array([b.position[0] for b in sc.bodies], np.float32).tobytes())
        digests = [None] * ws
        dist.all_gather_object(digests, (first, count, h.hexdigest()))
        t = reduce_max(1.0 + rank, dist)
        n = reduce_sum(float(count), dist)
        if rank == 0:
            ref = S.pyramid_batch(total, 3, first_seed=0)
            ok = True
            for (f, c, dg) in digests:
                hh = hashlib.sha256()
                for i in range(f, f + c):
                    hh.update(np.array([b.position[0] for b in ref.scene(i).bodies], np.float32).tobytes())
                ok &= hh.hexdigest() == dg
            print("RESULT", ok, t, n)
        dist.destroy_process_group()
    """ % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    path = os.path.join(str(tmp_path), "two_rank.py")
    with open(path, "w") as f:
        f.write(script)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", path], capture_output=True,
                       text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "True" and float(line[2]) == 2.0 and float(line[3]) == 37.0
