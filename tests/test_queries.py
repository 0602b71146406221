"""Batched read-only queries: b2gpu_batch_query_aabb / b2gpu_batch_raycast_closest against the
reference's b2World::QueryAABB / b2World::RayCast on the same worlds at the same step."""
import numpy as np
import pytest

from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S
from parity_util import NOCCD, same_bits

pytestmark = pytest.mark.gpu


def _scenes():
    return ([S.mixed(s, flags=NOCCD) for s in range(5)] + [S.sensors(s, flags=NOCCD) for s in range(3)] +
            [S.pyramid(6, flags=NOCCD, jitter_seed=s) for s in range(3)] + [S.pyramid(5, flags=NOCCD, ground="edge")] * 2 +
            [S.bullets(1, flags=NOCCD)] * 2) * 3  # 45 worlds: two tiles, the second ragged


def test_queries_match_reference(oracles):
    if not oracles.available("ref"):
        pytest.skip("queries are checked against the live reference")
    scenes = _scenes()
    refs = [oracles.RefWorld(s) for s in scenes]
    batch = B.WorldBatch(scenes)
    rs = np.random.RandomState(11)
    for rnd, steps in enumerate([1, 40, 80]):
        for _ in range(steps):
            batch.step(1 / 60, 8, 3)
            for r in refs:
                r.step(1 / 60, 8, 3)
        W = len(scenes)
        # ---- QueryAABB: 6 boxes per world, world-fastest ----
        nq = 6
        q = np.zeros(nq * W, B.AABB_QUERY_DTYPE)
        q["world"] = np.tile(np.arange(W), nq)
        c = np.stack([rs.uniform(-15, 15, nq * W), rs.uniform(-1, 12, nq * W)], 1)
        half = rs.uniform(0.05, 6.0, (nq * W, 2))
        q["lower"], q["upper"] = c - half, c + half
        counts, hits = batch.query_aabb(q, max_hits=128)
        total = 0
        for i in range(nq * W):
            want = np.sort(refs[q["world"][i]].query_aabb(q["lower"][i], q["upper"][i]).astype(np.uint32))
            got = hits[i, :counts[i]].astype(np.uint32)
            assert counts[i] == want.shape[0], f"round {rnd} query {i}: {counts[i]} hits, reference {want.shape[0]}"
            # ascending proxy id is not ascending (fixture | child << 16): compare as sets
            assert (np.sort(got) == want).all(), f"round {rnd} query {i}"
            total += counts[i]
        assert total > 50 * W // 10
        # ---- RayCast: 16 rays per world ----
        nr = 16
        rays = np.zeros(nr * W, B.RAY_DTYPE)
        rays["world"] = np.tile(np.arange(W), nr)
        rays["p1"] = np.stack([rs.uniform(-18, 18, nr * W), rs.uniform(4, 16, nr * W)], 1)
        rays["p2"] = np.stack([rs.uniform(-18, 18, nr * W), rs.uniform(-3, 6, nr * W)], 1)
        rays["maxFraction"] = 1.0
        out = batch.raycast_closest(rays)
        nhit = 0
        for i in range(nr * W):
            want = refs[rays["world"][i]].raycast_closest(rays["p1"][i], rays["p2"][i])
            if want is None:
                assert out["fixture"][i] == -1, f"round {rnd} ray {i}: device hit {out[i]}, reference none"
                continue
            nhit += 1
            f, ch, fr, pt, nm = want
            assert out["fixture"][i] == f and out["child"][i] == ch, f"round {rnd} ray {i}: {out[i]} vs {want}"
            assert same_bits(np.float32(out["fraction"][i]).reshape(1), np.float32(fr).reshape(1)), f"round {rnd} ray {i} fraction"
            assert same_bits(out["point"][i], pt) and same_bits(out["normal"][i], nm), f"round {rnd} ray {i}: {out[i]} vs {want}"
        assert nhit > nr * W // 3
    batch.close()


def test_query_edge_cases():
    batch = B.WorldBatch([S.hello_world()] * 2)
    counts, hits = batch.query_aabb(np.zeros(0, B.AABB_QUERY_DTYPE))
    assert counts.shape == (0,)
    q = np.zeros(1, B.AABB_QUERY_DTYPE)
    q[0] = (1, (-100, -100), (100, 100))
    counts, hits = batch.query_aabb(q, max_hits=1)  # truncated list, full count
    assert counts[0] == 2 and hits[0, 0] == 0
    q["world"] = 2
    with pytest.raises(B.B2GpuError):
        batch.query_aabb(q)
    rays = np.zeros(2, B.RAY_DTYPE)
    rays[0] = (0, (0, 50), (0, 20), 1.0)   # ends above everything
    rays[1] = (1, (0, 50), (0, -5), 1.0)   # straight down onto the box at (0, 4)
    out = batch.raycast_closest(rays)
    assert out["fixture"][0] == -1
    assert out["fixture"][1] == 1 and out["normal"][1].tolist() == [0.0, 1.0] and abs(out["point"][1][1] - 5.0) < 1e-5
    batch.close()
