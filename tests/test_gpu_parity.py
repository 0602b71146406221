"""GPU parity: the CUDA path (through the C ABI) against the CPU oracles, bit for bit.

Bar (BASELINE.json north_star): proxy-pair sets and manifold point counts bit-exact; body state within
1e-4 relative after 60 steps.  Because the device carries a libm-identical sinf/cosf and the build disables
FMA contraction, these tests assert the stronger property: EVERY float of every body, manifold, fat AABB
and listener event is bit-identical to the unmodified reference, every step.
"""
import json
import os

import numpy as np
import pytest

import golden_cases as G
from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S
from box2dpp_b200.batch import EVENT_ALL, WorldBatch
from parity_util import NOCCD, compare_world

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.gpu


def _ref_cls(oracles):
    """The GPU parity tests compare with the reference itself (oracle/_ref, built where /root/reference exists and shipped
    to the GPU box as a built library).  Without it they fail loudly instead of quietly comparing with the restatement;
    B2GPU_ALLOW_PORT_ORACLE=1 opts into that weaker check."""
    if oracles.available("ref"):
        return oracles.RefWorld
    if os.environ.get("B2GPU_ALLOW_PORT_ORACLE") == "1":
        return oracles.PortWorld
    pytest.fail("oracle/_ref/libb2ref.so is missing: build it with `make -C oracle ref` where /root/reference exists "
                "(or set B2GPU_ALLOW_PORT_ORACLE=1 to compare with the C restatement instead)")


def test_parity_oracle_is_the_reference_itself(oracles):
    """Every test below is a claim against the unmodified reference: fail here, by name, if its built library did not travel."""
    assert _ref_cls(oracles).KIND == "ref" or os.environ.get("B2GPU_ALLOW_PORT_ORACLE") == "1"


def run_parity(oracles, scenes, steps, vi, pi, check_every=1, events=True, dt=1.0 / 60.0, generic_solver=False):
    Ref = _ref_cls(oracles)
    refs = [Ref(s, record_events=events) for s in scenes]
    batch = WorldBatch(scenes, event_mask=EVENT_ALL if events else 0, generic_solver=generic_solver)
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


def test_mixed_shapes_every_step(oracles):
    # per-step check: catches state that is only wrong for a step (e.g. a static body's awake flag when islands
    # sharing it fall asleep at different times: b2World::Solve builds and solves islands alternately)
    run_parity(oracles, [S.mixed(seed, flags=NOCCD) for seed in (1, 6)], 400, 8, 3)


def test_generic_solver_path(oracles):
    """The fallback solver kernel (body state in global memory, used when a tile's body table exceeds shared
    memory) stays bit-exact too."""
    scenes = [S.pyramid(6 + (w % 3), flags=NOCCD, jitter_seed=w) for w in range(6)] + [S.mixed(1, flags=NOCCD)]
    run_parity(oracles, scenes, 120, 8, 3, check_every=6, generic_solver=True)


def test_edge_ground_pyramid(oracles):
    run_parity(oracles, [S.pyramid(8, flags=NOCCD, ground="edge")], 250, 8, 3, check_every=5)


# ---- continuous physics ON: b2World::SolveTOI on the device -------------------------
def _need_ref(oracles):
    if not oracles.available("ref"):
        pytest.skip("continuous physics parity needs oracle/_ref (the C restatement stops at SolveTOI); "
                    "the golden-hash tests below cover it without the library")


def test_hello_world_default_flags(oracles):
    # BASELINE config 1 exactly as HelloWorld/HelloWorld.cpp runs it (TOI event at the landing)
    _need_ref(oracles)
    run_parity(oracles, [S.hello_world()], 60, 6, 2)


def test_pyramid_20_default_flags(oracles):
    # BASELINE config 2 with the reference's defaults: 386 TOI events in the first 210 steps, asleep from 212
    _need_ref(oracles)
    run_parity(oracles, [S.pyramid(20)], 260, 8, 3, check_every=3)


def test_toi_batch_divergent(oracles):
    # different worlds in one tile: lanes with and without TOI events, different sub-step counts
    _need_ref(oracles)
    scenes = [S.pyramid(5 + (w % 4), jitter_seed=w) for w in range(20)] + [S.mixed(s) for s in range(6)] + \
             [S.bullets(s) for s in range(8)] + [S.hello_world()]
    run_parity(oracles, scenes, 200, 8, 3, check_every=8)


def test_toi_bullets_every_step(oracles):
    _need_ref(oracles)
    run_parity(oracles, [S.bullets(s, shots=9) for s in range(5)], 150, 6, 2)


def test_toi_generic_solver_and_mixed_flags(oracles):
    # worlds with and without continuous physics in the same batch, on the fallback solver path
    _need_ref(oracles)
    scenes = [S.bullets(2), S.bullets(2, flags=NOCCD), S.pyramid(6), S.pyramid(6, flags=NOCCD), S.mixed(3)]
    run_parity(oracles, scenes, 150, 8, 3, check_every=5, generic_solver=True)


def test_filters_damping_gravity_scale_fixed_rotation(oracles):
    """b2ContactFilter::ShouldCollide with non-default categoryBits / maskBits / groupIndex (positive and negative groups,
    b2WorldCallbacks.cpp:26-39), linear / angular damping (b2Island.cpp:217-218), gravityScale outside {0, 1} and
    fixedRotation bodies in contact: every float, contact and event, every step, tile solver."""
    run_parity(oracles, [S.filtered(s, flags=NOCCD) for s in range(6)], 220, 8, 3)
    run_parity(oracles, [S.filtered(s, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP) for s in range(6, 10)], 120, 6, 2)


def test_filters_damping_default_flags_and_generic_solver(oracles):
    """The same with continuous physics on (the filter also gates TOI candidates) and on the generic solver kernel."""
    run_parity(oracles, [S.filtered(s) for s in range(10, 16)], 200, 8, 3, check_every=4)
    run_parity(oracles, [S.filtered(s) for s in range(16, 20)], 120, 8, 3, check_every=4, generic_solver=True)


def test_sensors_every_step(oracles):
    # sensor fixtures (b2Contact::Update's b2TestOverlap branch): Begin/End without PreSolve/PostSolve, no wake-ups,
    # never in an island or a TOI; with and without continuous physics, sleeping allowed, mixed with solid worlds
    _need_ref(oracles)
    scenes = [S.sensors(s, flags=(S.WORLD_DEFAULT if s % 2 else NOCCD)) for s in range(6)] + [S.bullets(4), S.mixed(1)]
    run_parity(oracles, scenes, 320, 8, 3)


def test_sensors_generic_solver(oracles):
    _need_ref(oracles)
    run_parity(oracles, [S.sensors(7), S.sensors(8, flags=NOCCD)], 200, 6, 2, check_every=4, generic_solver=True)


def test_joints_every_step(oracles):
    # revolute (limits, motors), prismatic (limits, motor) and distance (rigid, soft) joints in the island solver:
    # joint edges in the island DFS, ShouldCollide's collideConnected rule, joint init / velocity / position passes
    # interleaved with the contacts', warm starting of joint impulses; with and without continuous physics
    _need_ref(oracles)
    scenes = [S.jointed(s, flags=(S.WORLD_DEFAULT if s % 2 else NOCCD)) for s in range(6)] + [S.mixed(4), S.pyramid(5)]
    run_parity(oracles, scenes, 420, 8, 3)


def test_wheel_and_weld_joints_every_step(oracles):
    # b2WheelJoint (spring, motor, point-to-line) and b2WeldJoint (rigid 3x3, soft angular spring, the both-bodies-
    # fixed-rotation branch) on chain terrain, with and without continuous physics and warm starting
    _need_ref(oracles)
    scenes = [S.vehicles(s, flags=(S.WORLD_DEFAULT if s % 2 else NOCCD)) for s in range(5)] + \
             [S.vehicles(7, flags=S.WORLD_DEFAULT & ~S.WORLD_WARM_STARTING), S.jointed(3)]
    run_parity(oracles, scenes, 420, 8, 3)


def test_joints_no_warm_starting_and_62(oracles):
    _need_ref(oracles)
    scenes = [S.jointed(9, flags=S.WORLD_DEFAULT & ~S.WORLD_WARM_STARTING), S.jointed(10, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP)]
    run_parity(oracles, scenes, 300, 6, 2, check_every=3)


def test_joints_generic_solver(oracles):
    # the same joint code on the body arrays in global memory (k_solve): the path of batches whose body table does not
    # fit shared memory
    _need_ref(oracles)
    scenes = [S.jointed(2), S.vehicles(3, flags=NOCCD), S.jointed(5, flags=NOCCD)]
    run_parity(oracles, scenes, 240, 8, 3, check_every=2, generic_solver=True)


def test_binned_and_lane_mapped_narrow_phase_agree():
    # k_collide's type-binned evaluation (worlds that mix shape types) against the lane-per-world evaluation
    scenes = [S.mixed(s) for s in range(20)] + [S.sensors(s) for s in range(8)] + [S.vehicles(s) for s in range(6)] + \
             [S.pyramid(6, jitter_seed=s) for s in range(6)]
    a = WorldBatch(scenes, event_mask=EVENT_ALL, binned_collide=True)
    b = WorldBatch(scenes, event_mask=EVENT_ALL, binned_collide=False)
    for step in range(200):
        a.step(1.0 / 60.0, 8, 3)
        b.step(1.0 / 60.0, 8, 3)
        if step % 10 == 9:
            for fid in (B.F_BODY_XF, B.F_BODY_VEL, B.F_CONTACT_A, B.F_CONTACT_K, B.F_CONTACT_P0, B.F_CONTACT_P1, B.F_CONTACT_N,
                        B.F_PROXY_FAT, B.F_EVENT_COUNT):
                assert (a.field(fid).view(np.uint32) == b.field(fid).view(np.uint32)).all(), f"field {fid} step {step}"
    assert (a.status() == 0).all() and (b.status() == 0).all()
    a.close()
    b.close()


def test_compacted_event_readback():
    # b2gpu_batch_read_events == the per-world event arenas, world by world, in order
    scenes = [S.pyramid(4 + w % 3, jitter_seed=w) for w in range(40)] + [S.sensors(1), S.hello_world()]
    batch = WorldBatch(scenes, event_mask=EVENT_ALL)
    seen = 0
    for step in range(60):
        batch.step(1.0 / 60.0, 8, 3)
        if step % 7 == 0 or step == 59:
            offsets, ev = batch.read_events()
            assert offsets[0] == 0 and offsets[-1] == ev.shape[0]
            for w in range(len(scenes)):
                ints, fl = batch.events(w)
                e = ev[offsets[w]:offsets[w + 1]]
                assert e.shape[0] == ints.shape[0]
                got = np.stack([e["type"], e["fixtureA"], e["childA"], e["fixtureB"], e["childB"], e["count"]], 1)
                assert (got == ints).all()
                want = np.concatenate([e["normalImpulses"], e["tangentImpulses"]], 1)
                assert (want.view(np.uint32) == fl.view(np.uint32)).all()
            seen += ev.shape[0]
    assert seen > 1000
    batch.close()


def test_world_groups_do_not_change_results(monkeypatch):
    # the step chain enqueued per group of tiles on separate streams (incl. the 8-step graph chunks of step_n) against
    # a single group on one stream: every field identical
    scenes = [S.pyramid(5 + w % 3, jitter_seed=w) for w in range(150)] + [S.mixed(w) for w in range(10)]  # 5 tiles
    batches = []
    for groups in ("1", "4"):
        monkeypatch.setenv("B2GPU_GROUPS", groups)
        batches.append(WorldBatch(scenes, event_mask=EVENT_ALL))
    a, b = batches
    for n in (1, 19, 8, 3):
        a.step_n(1.0 / 60.0, 8, 3, n)
        b.step_n(1.0 / 60.0, 8, 3, n)
        a.step(1.0 / 60.0, 8, 3)
        b.step(1.0 / 60.0, 8, 3)
        for fid in (B.F_BODY_XF, B.F_BODY_SWEEP, B.F_BODY_VEL, B.F_BODY_FLAGS, B.F_CONTACT_COUNT, B.F_CONTACT_A, B.F_CONTACT_P0,
                    B.F_PROXY_FAT, B.F_EVENT_COUNT, B.F_EVENT_A, B.F_STEP_COUNT):
            assert (a.field(fid).view(np.uint32) == b.field(fid).view(np.uint32)).all(), f"field {fid} after {n}"
    assert (a.status() == 0).all() and (b.status() == 0).all()
    a.close()
    b.close()


def test_snapshot_restore_continues_bit_identically():
    # b2gpu_batch_snapshot_save / load: restore into the same batch after it moved on, and into a fresh batch built from
    # the same scenes; both continue exactly like the original (contacts, impulses, joints, sleep state, TOI on)
    scenes = [S.pyramid(5, jitter_seed=1), S.mixed(2), S.jointed(1), S.vehicles(2), S.sensors(1), S.bullets(2)] * 6
    a = WorldBatch(scenes, event_mask=EVENT_ALL)
    a.step_n(1.0 / 60.0, 8, 3, 70)
    cmds = np.zeros(2, B.BODY_COMMAND_DTYPE)
    cmds[0] = (0, 3, B.CMD_SET_TRANSFORM, 1, (0.5, 6.0), (0.3, 0.0))  # a buffered move is part of the state
    cmds[1] = (2, 12, B.CMD_APPLY_LINEAR_IMPULSE, 1, (2.0, 5.0), (0.0, 3.0))
    a.body_commands(cmds)
    blob = a.snapshot()
    fresh = WorldBatch(scenes, event_mask=EVENT_ALL)
    fresh.restore(blob)
    fields = (B.F_BODY_XF, B.F_BODY_SWEEP, B.F_BODY_VEL, B.F_BODY_FLAGS, B.F_CONTACT_COUNT, B.F_CONTACT_A, B.F_CONTACT_K,
              B.F_CONTACT_P0, B.F_CONTACT_P1, B.F_PROXY_FAT, B.F_JOINT_IMPULSE, B.F_JOINT_I, B.F_EVENT_COUNT, B.F_STEP_COUNT)
    ref_states = []
    for _ in range(4):
        a.step_n(1.0 / 60.0, 8, 3, 15)
        ref_states.append([a.field(f).copy() for f in fields])
    a.restore(blob)  # rewind
    for k in range(4):
        a.step_n(1.0 / 60.0, 8, 3, 15)
        fresh.step_n(1.0 / 60.0, 8, 3, 15)
        for f, want in zip(fields, ref_states[k]):
            assert (a.field(f).view(np.uint32) == want.view(np.uint32)).all(), f"rewound batch, field {f}, block {k}"
            assert (fresh.field(f).view(np.uint32) == want.view(np.uint32)).all(), f"fresh batch, field {f}, block {k}"
    with pytest.raises(B.B2GpuError):
        WorldBatch(scenes[:6]).restore(blob)
    a.close()
    fresh.close()


def test_edge_cases_empty_world_zero_dt_and_overflow(oracles):
    """Edge cases: a world without bodies and one with only a static body next to normal worlds (ragged batch); a step with
    dt = 0 (b2World::Step then only runs Collide, b2World.cpp:926-944) matches the oracle; a contact arena that is too
    small reports B2GPU_STATUS_CONTACT_OVERFLOW instead of corrupting memory; an event arena likewise."""
    Ref = _ref_cls(oracles)
    empty = S.Scene()
    lonely = S.Scene()
    lonely.create_box(lonely.create_body(S.STATIC_BODY), 5.0, 1.0)
    scenes = [empty, S.pyramid(4, flags=NOCCD), lonely, S.hello_world(NOCCD)]
    batch = WorldBatch(scenes, event_mask=EVENT_ALL)
    refs = [Ref(s, record_events=True) for s in scenes[1:]]
    for step in range(50):
        dt = 0.0 if step in (7, 8, 30) else 1.0 / 60.0
        batch.step(dt, 8, 3)
        for r in refs:
            r.step(dt, 8, 3)
        for w, r in enumerate(refs, start=1):
            compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} step {step}")
    assert batch.bodies(0).shape[0] == 0 and batch.contacts(0)[0].shape[0] == 0
    assert (batch.status() == 0).all()
    batch.close()
    # contact arena with room for 8 contacts only: a 5-row pyramid needs ~40
    small = WorldBatch([S.pyramid(5, flags=NOCCD), S.hello_world(NOCCD)], max_contacts=8)
    small.step_n(1.0 / 60.0, 8, 3, 30)
    st = small.status()
    assert st[0] & 1 and st[1] == 0, st
    assert int(small.field(B.F_CONTACT_COUNT)[0, 0, 0]) <= 8
    small.close()
    few = WorldBatch([S.pyramid(5, flags=NOCCD)], event_mask=EVENT_ALL, max_events=4)
    few.step_n(1.0 / 60.0, 8, 3, 40)
    assert few.status()[0] & 2
    offsets, ev = few.read_events()
    assert ev.shape[0] <= 4
    few.close()


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


# ---- against the frozen reference answers (tests/golden/, no oracle library needed) ---------------
@pytest.mark.parametrize("name,kind", [("polygons", B.COLLIDE_POLYGONS), ("circles", B.COLLIDE_CIRCLES),
                                       ("polygon_circle", B.COLLIDE_POLYGON_CIRCLE),
                                       ("edge_circle", B.COLLIDE_EDGE_CIRCLE), ("edge_polygon", B.COLLIDE_EDGE_POLYGON)])
def test_device_narrow_phase_matches_golden(name, kind):
    """The reference's narrow-phase known-answer tests (Regression/tests/helloworld.cpp:268-418) and the seeded
    random pairs, evaluated by the device functions the collide kernel calls: manifold ids, point counts and
    every float bit-exact."""
    g = np.load(os.path.join(GOLD, "narrowphase.npz"))
    cases = G.NARROW_CASES[name][1]()
    a, xa, b, xb = (np.stack([c[i] for c in cases]) for i in range(4))
    ints, fl = B.collide_pairs(kind, a, xa, b, xb)
    gi, gf = g[name + "_ints"], g[name + "_floats"]
    assert (ints == gi).all(), f"{name}: ids / point counts differ at {np.argwhere(ints != gi)[:4].tolist()}"
    assert (fl.view(np.uint32) == gf.view(np.uint32)).all(), f"{name}: manifold floats differ"


def test_device_narrow_phase_large_random(oracles):
    """200k random convex-polygon pairs, device vs the C restatement (bit-exact)."""
    rs = np.random.RandomState(5)
    n = 200000
    base = [G.random_convex(rs) for _ in range(256)]
    ia, ib = rs.randint(0, 256, n), rs.randint(0, 256, n)
    a, b = np.stack(base)[ia], np.stack(base)[ib]
    ang = rs.uniform(-3.2, 3.2, (n, 2)).astype(np.float32)
    xa = np.stack([rs.uniform(-.5, .5, n), rs.uniform(-.5, .5, n), np.sin(ang[:, 0]), np.cos(ang[:, 0])], 1).astype(np.float32)
    xb = np.stack([rs.uniform(-2, 2, n), rs.uniform(-2, 2, n), np.sin(ang[:, 1]), np.cos(ang[:, 1])], 1).astype(np.float32)
    ints, fl = B.collide_pairs(B.COLLIDE_POLYGONS, a, xa, b, xb)
    assert (ints[:, 0] > 0).mean() > 0.1
    for i in rs.randint(0, n, 4000):
        ri, rf = oracles.collide_polygons(a[i], xa[i], b[i], xb[i], kind="port")
        assert (ri == ints[i]).all() and (rf.view(np.uint32) == fl[i].view(np.uint32)).all(), f"pair {i}"


@pytest.mark.parametrize("name", list(G.WORLD_CASES))
def test_worlds_match_golden_hashes(name):
    """Every parity scene against the per-step hashes frozen from the reference (body state bits, contact-list
    order with point counts, contact / touching / manifold-point totals)."""
    g = json.load(open(os.path.join(GOLD, "worlds.json")))[name]
    make, steps, vi, pi = G.WORLD_CASES[name]
    sc = make()
    fb = G.fixture_bodies(sc)
    batch = WorldBatch([sc])
    stride = 1 if steps <= 250 else 6
    for step in range(steps):
        batch.step(1.0 / 60.0, vi, pi)
        if step % stride == 0 or step == steps - 1:
            ints, _ = batch.contacts(0)
            touching = (ints[:, 4] & 1) != 0
            assert [len(ints), int(touching.sum()), int(ints[touching, 5].sum())] == g["contacts_touching_points"][step], \
                f"{name} step {step}: contact / touching / point counts"
            assert "%016x" % G.contact_order_hash(ints, fb) == g["contact_order_hash"][step], f"{name} step {step}: order"
            assert "%016x" % G.state_hash(batch.bodies(0)) == g["state_hash"][step], f"{name} step {step}: state"
    assert (batch.status() == 0).all()
    batch.close()


def test_step_io_equals_separate_calls():
    """b2gpu_batch_step_io (per-group copies on the groups' streams) == apply_forces + step + read_transforms,
    bit-for-bit, with a ragged last tile and more tiles than groups, and for every group count."""
    W = 150  # 5 tiles, the last one with 22 worlds
    sb = S.pyramid_batch(W, 6, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP, first_seed=3, jittered=True)
    rng = np.random.default_rng(5)
    a = WorldBatch(sb)
    a_max_bodies = a.max_bodies
    forces = np.zeros((W, a.max_bodies, 3), np.float32)
    seq = []
    for k in range(30):
        f = forces.copy()
        f[:, 1:, :] = rng.uniform(-40.0, 40.0, size=f[:, 1:, :].shape).astype(np.float32) * (rng.random((W, a.max_bodies - 1, 1)) < 0.3)
        seq.append(f)
    ref_out = []
    for f in seq:
        a.apply_forces(f)
        a.step(1.0 / 60.0, 8, 3)
        ref_out.append(a.read_transforms().copy())
    a.close()
    for groups in ("1", "3", None, "8"):
        if groups is None:
            os.environ.pop("B2GPU_IO_GROUPS", None)
        else:
            os.environ["B2GPU_IO_GROUPS"] = groups
        try:
            b = WorldBatch(sb)
            out = np.empty((W, b.max_bodies, 4), np.float32)
            for k, f in enumerate(seq):
                b.step_io(f, 1.0 / 60.0, 8, 3, out)
                assert (out.view(np.uint32) == ref_out[k].view(np.uint32)).all(), f"step {k}, io groups {groups}"
            # either side may be skipped
            b.step_io(None, 1.0 / 60.0, 8, 3, out)
            b.close()
            if groups is None:  # pinned buffers: the call is replayed from a captured graph (same pointers every step)
                import torch
                fp = torch.zeros((W, a_max_bodies, 3), dtype=torch.float32).pin_memory()
                op = torch.empty((W, a_max_bodies, 4), dtype=torch.float32).pin_memory()
                b = WorldBatch(sb)
                for k, f in enumerate(seq):
                    fp.numpy()[:] = f
                    b.step_io(fp.numpy(), 1.0 / 60.0, 8, 3, op.numpy())
                    assert (op.numpy().view(np.uint32) == ref_out[k].view(np.uint32)).all(), f"step {k}, graph path"
                b.close()
        finally:
            os.environ.pop("B2GPU_IO_GROUPS", None)


def test_step_io_forces_wake_sleepers_like_the_separate_calls():
    """Worlds that do not allow sleeping let b2gpu_batch_step_io bring its forces in behind k_collide / k_islands (nobody can be
    woken by them) — until a body is put to sleep by hand: from then on a force has to wake it BEFORE the step looks at the
    contacts, as b2Body::ApplyForce before b2World::Step does.  step_io == apply_forces + step + read_transforms throughout."""
    W = 96
    sb = S.pyramid_batch(W, 5, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP, first_seed=11, jittered=True)
    a, b = WorldBatch(sb), WorldBatch(sb)
    rng = np.random.default_rng(9)
    nb = a.max_bodies
    for k in range(90):
        f = (rng.uniform(-30.0, 30.0, size=(W, nb, 3)) * (rng.random((W, nb, 1)) < 0.2)).astype(np.float32)
        if k == 40:  # every settled box of the even worlds goes to sleep: b2Body::SetAwake(false)
            cmds = np.array([(w, body, B.CMD_SET_AWAKE, 0, (0, 0), (0, 0)) for w in range(0, W, 2) for body in range(1, nb)],
                            dtype=B.BODY_COMMAND_DTYPE)
            a.body_commands(cmds)
            b.body_commands(cmds)
        a.apply_forces(f)
        a.step(1.0 / 60.0, 8, 3)
        want = a.read_transforms()
        got = b.step_io(f, 1.0 / 60.0, 8, 3)
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), f"step {k}"
    for w in (0, 1, 50, 95):
        assert (a.bodies(w).view(np.uint32) == b.bodies(w).view(np.uint32)).all()
    a.close()
    b.close()


def test_step_io_with_selected_bodies_and_poses():
    """b2gpu_batch_io_select: forces for a subset of the bodies and (x, y, angle) poses of the dynamic ones only give the
    same worlds as the dense calls with zeros elsewhere, and the poses are the transforms' positions + the sweeps' angles."""
    scenes = [S.pyramid(5, flags=NOCCD, jitter_seed=w) for w in range(70)]
    dense, lean = WorldBatch(scenes), WorldBatch(scenes)
    nb = dense.max_bodies
    pushed = [1, 4, 9, nb - 1]
    dynamic = list(range(1, nb))  # body 0 is the static ground
    lean.io_select(force_bodies=pushed, pose_bodies=dynamic, xya=True)
    rs = np.random.RandomState(5)
    for step in range(40):
        f_lean = rs.uniform(-40.0, 40.0, (len(scenes), len(pushed), 3)).astype(np.float32)
        if step % 3 == 0:
            f_lean[rs.randint(0, len(scenes), 20)] = 0.0  # worlds without any force this step
        f_dense = np.zeros((len(scenes), nb, 3), np.float32)
        f_dense[:, pushed, :] = f_lean
        xf = dense.step_io(f_dense, 1.0 / 60.0, 8, 3)
        pose = lean.step_io(f_lean, 1.0 / 60.0, 8, 3)
        assert pose.shape == (len(scenes), nb - 1, 3)
        assert (pose[:, :, :2].view(np.uint32) == xf[:, 1:, :2].view(np.uint32)).all(), f"step {step}: positions"
        if step % 10 == 9:
            for w in (0, 33, 69):
                assert (dense.bodies(w).view(np.uint32) == lean.bodies(w).view(np.uint32)).all(), f"step {step} world {w}"
                assert (pose[w, :, 2].view(np.uint32) == lean.bodies(w)[1:, 6].view(np.uint32)).all(), f"step {step}: angles"
    # the separate calls honour the selection too
    lean.apply_forces(np.zeros((len(scenes), len(pushed), 3), np.float32))
    assert lean.read_transforms().shape == (len(scenes), nb - 1, 3)
    # back to every body, four floats
    lean.io_select()
    assert lean.step_io(None, 1.0 / 60.0, 8, 3).shape == (len(scenes), nb, 4)
    dense.close()
    lean.close()


def test_full_size_batch_properties(oracles):
    """BASELINE configs[2] at full size: 16,384 pyramid-20 worlds on one B200, compared where bench.py measures (steps
    90 - 125: ~590 contacts, all three position passes).  Size-independent properties:
    (a) 64 spot-checked jittered worlds equal the reference bit-for-bit at step 120; no arena overflow
        anywhere;
    (b) determinism: a second batch stepped one step at a time (no graph replay) agrees bit-for-bit with the first."""
    flags = NOCCD & ~S.WORLD_ALLOW_SLEEP
    steps = 120
    W = 16384
    sb = S.pyramid_batch(W, 20, flags=flags, first_seed=0, jittered=True)
    batch = WorldBatch(sb)
    batch.step_n(1.0 / 60.0, 8, 3, steps)
    batch.sync()
    assert (batch.status() == 0).all()
    xf = batch.read_transforms().copy()
    Ref = _ref_cls(oracles)
    rs = np.random.RandomState(2024)
    picked = [0, 1, 31, 32, 33, 1000, 4095, 4096, 8191, 8192, 12345, 16000, 16352, 16382, 16383, 777]
    picked += [int(w) for w in rs.choice(W, 48, replace=False)]
    for w in picked:
        ref = Ref(sb.scene(w))
        for _ in range(steps):
            ref.step(1.0 / 60.0, 8, 3)
        compare_world(ref, batch.bodies(w), batch.contacts(w), batch.proxies(w), what=f"jittered world {w}")
    batch.close()
    batch2 = WorldBatch(sb)
    for _ in range(steps):
        batch2.step(1.0 / 60.0, 8, 3)
    xf2 = batch2.read_transforms()
    differ = np.nonzero((xf.view(np.uint32) != xf2.view(np.uint32)).any(axis=(1, 2)))[0]
    assert len(differ) == 0, f"graph replay and single steps disagree in {len(differ)} worlds, e.g. {differ[:8]}"
    batch2.close()


def test_mixed_worlds_random_seeds_of_the_full_job(oracles):
    """BASELINE configs[3] shards 131,072 mixed-shape worlds (seed = world index) over 8 GPUs: 64 seeds drawn from the whole
    range, every float and contact against the reference every 10th step for 240 steps, continuous physics off (as benched)
    and on (the reference's default)."""
    rs = np.random.RandomState(77)
    seeds = sorted(int(x) for x in rs.choice(131072, 64, replace=False))
    run_parity(oracles, [S.mixed(s, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP) for s in seeds[:40]], 240, 8, 3, check_every=10, events=False)
    run_parity(oracles, [S.mixed(s) for s in seeds[40:]], 240, 8, 3, check_every=10, events=False)
