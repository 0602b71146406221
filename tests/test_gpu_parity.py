"""GPU parity: the CUDA path (through the C ABI) against the CPU oracles, bit for bit.

Bar (BASELINE.json north_star): proxy-pair sets and manifold point counts bit-exact; body state within
1e-4 relative after 60 steps.  Because the device carries a libm-identical sinf/cosf and the build disables
FMA contraction, these tests assert the stronger property: EVERY float of every body, manifold, fat AABB
and listener event is bit-identical to the unmodified reference, every step.
"""
import numpy as np
import pytest

from box2dpp_b200 import scene as S
from box2dpp_b200.batch import EVENT_ALL, WorldBatch
from parity_util import NOCCD, compare_world

pytestmark = pytest.mark.gpu


def _ref_cls(oracles):
    return oracles.RefWorld if oracles.available("ref") else oracles.PortWorld


def run_parity(oracles, scenes, steps, vi, pi, check_every=1, events=True, dt=1.0 / 60.0):
    Ref = _ref_cls(oracles)
    refs = [Ref(s, record_events=events) for s in scenes]
    batch = WorldBatch(scenes, event_mask=EVENT_ALL if events else 0)
    for step in range(steps):
        for r in refs:
            r.step(dt, vi, pi)
        batch.step(dt, vi, pi)
        if step % check_every == 0 or step == steps - 1:
            for w, r in enumerate(refs):
                compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w),
                              batch.events(w) if events else None, what=f"world {w} step {step}")
    assert (batch.status() == 0).all()
    batch.close()


def test_hello_world(oracles):
    # BASELINE config 1 (HelloWorld/HelloWorld.cpp:36-94), continuous physics off
    run_parity(oracles, [S.hello_world(NOCCD)], 120, 6, 2)


def test_pyramid_small(oracles):
    run_parity(oracles, [S.pyramid(5, flags=NOCCD)], 200, 8, 3)


def test_pyramid_20_full(oracles):
    # BASELINE config 2: 600 steps, per-step state parity, sleeping on (falls asleep around step 254)
    run_parity(oracles, [S.pyramid(20, flags=NOCCD)], 600, 8, 3, check_every=7, events=False)


def test_pyramid_20_events_no_sleep(oracles):
    run_parity(oracles, [S.pyramid(20, flags=NOCCD & ~1)], 90, 6, 2, check_every=5)


def test_jittered_pyramids_batch(oracles):
    # BASELINE config 3 flavour: many different worlds in one batch (divergent control flow inside a warp)
    scenes = [S.pyramid(6 + (w % 5), flags=NOCCD, jitter_seed=w) for w in range(40)]
    run_parity(oracles, scenes, 150, 8, 3, check_every=10, events=False)


def test_mixed_shapes(oracles):
    # BASELINE config 4: circles, convex polygons, chain terrain; many islands, sleeping per island
    scenes = [S.mixed(seed, flags=NOCCD) for seed in range(8)]
    run_parity(oracles, scenes, 400, 8, 3, check_every=9)


def test_edge_ground_pyramid(oracles):
    run_parity(oracles, [S.pyramid(8, flags=NOCCD, ground="edge")], 250, 8, 3, check_every=5)


def test_replicated_worlds_identical(oracles):
    sc = S.pyramid(10, flags=NOCCD)
    batch = WorldBatch(sc, replicate=70)
    batch.step_n(1.0 / 60.0, 8, 3, 80)
    batch.sync()
    xf = batch.read_transforms()
    assert (xf.view(np.uint32) == xf[0].view(np.uint32)).all()
    ref = _ref_cls(oracles)(sc)
    for _ in range(80):
        ref.step(1.0 / 60.0, 8, 3)
    compare_world(ref, batch.bodies(69), batch.contacts(69), what="replica 69")
    batch.close()
