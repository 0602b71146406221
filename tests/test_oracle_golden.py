"""CPU tests of the oracle (TEST INFRASTRUCTURE): the C restatement oracle/b2o_*.c against

  * the golden fixtures frozen from the reference's vanilla Box2D (tests/golden/, tools/make_golden.py),
  * SURVEY.md Appendix D's independently produced checkpoints, and
  * the live reference library oracle/_ref/libb2ref.so when it is present (build container).

Bar: bit-exact everywhere (integer ids, point counts, and every float32).
"""
import json
import os

import numpy as np
import pytest

import golden_cases as G
from box2dpp_b200 import scene as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold_worlds():
    return json.load(open(os.path.join(GOLD, "worlds.json")))


@pytest.fixture(scope="module")
def gold_narrow():
    return np.load(os.path.join(GOLD, "narrowphase.npz"))


def _kinds(oracles):
    return ["port"] + (["ref"] if oracles.available("ref") else [])


@pytest.mark.parametrize("name", list(G.NARROW_CASES))
def test_narrow_phase_matches_golden(oracles, gold_narrow, name):
    fn, gen = G.NARROW_CASES[name]
    cases = gen()
    gi, gf = gold_narrow[name + "_ints"], gold_narrow[name + "_floats"]
    assert len(cases) == len(gi)
    assert (gi[:, 0] > 0).sum() > len(cases) // 10, "fixture must contain touching cases"
    for kind in _kinds(oracles):
        for i, (a, xa, b, xb) in enumerate(cases):
            ints, fl = getattr(oracles, fn)(a, xa, b, xb, kind=kind)
            assert (ints == gi[i]).all(), f"{kind} {name}[{i}]: ids/pointCount {ints} vs golden {gi[i]}"
            assert (fl.view(np.uint32) == gf[i].view(np.uint32)).all(), f"{kind} {name}[{i}]: manifold floats differ"


def test_reference_known_answer_tests(gold_narrow):
    """The reference's own narrow-phase KATs (Regression/tests/helloworld.cpp:268-293, 393-418) pin
    pointCount; these are cases 0/1 of the circles fixture and case 0 of the polygons fixture."""
    assert gold_narrow["circles_ints"][0, 0] == 1 and gold_narrow["circles_ints"][1, 0] == 1
    # two unit boxes rotated by 1 rad with centres 2 apart overlap at their corners: a 2-point face manifold
    assert gold_narrow["polygons_ints"][0, 0] == 2
    # the separated pair of Testb2FindMaxSeparation (:295-325) yields no manifold
    assert gold_narrow["polygons_ints"][2, 0] == 0


def test_polygon_build_matches_golden(oracles):
    g = np.load(os.path.join(GOLD, "polygons.npz"))
    for kind in _kinds(oracles):
        for i, (fd, verts) in enumerate(G.polygon_cases()):
            n, out = oracles.polygon_build(fd, verts, kind=kind)
            assert n == g["counts"][i], f"{kind} polygon {i}: vertex count {n} vs {g['counts'][i]}"
            assert (out.view(np.uint32) == g["outs"][i].view(np.uint32)).all(), f"{kind} polygon {i}: hull/mass differ"


def _run_world(World, name):
    make, steps, vi, pi = G.WORLD_CASES[name]
    sc = make()
    w = World(sc)
    fb = G.fixture_bodies(sc)
    for step in range(steps):
        w.step(1.0 / 60.0, vi, pi)
        ints, _ = w.contacts()
        touching = (ints[:, 4] & 1) != 0
        yield step, w.bodies(), ints, fb, [len(ints), int(touching.sum()), int(ints[touching, 5].sum())]


@pytest.mark.parametrize("name", G.NOCCD_CASES)
def test_port_worlds_match_golden(oracles, gold_worlds, name):
    """The C restatement reproduces the reference's per-step body-state and contact-order hashes."""
    g = gold_worlds[name]
    stride = 1 if g["steps"] <= 250 else 4  # the pure-Python FNV is the slow part
    for step, bodies, ints, fb, counts in _run_world(oracles.PortWorld, name):
        assert counts == g["contacts_touching_points"][step], f"{name} step {step}: contact/touching/point counts"
        if step % stride == 0 or step == g["steps"] - 1:
            assert "%016x" % G.contact_order_hash(ints, fb) == g["contact_order_hash"][step], f"{name} step {step}: contact order"
            assert "%016x" % G.state_hash(bodies) == g["state_hash"][step], f"{name} step {step}: body state"


@pytest.mark.parametrize("name", list(G.APPENDIX_D))
def test_golden_reproduces_survey_appendix_d(gold_worlds, name):
    """The frozen fixtures agree with the checkpoints SURVEY.md Appendix D obtained from vanilla with an
    independent probe: the golden file and the survey pin each other."""
    g = gold_worlds[name]
    for step, (contacts, touching, points, sh, ch) in G.APPENDIX_D[name].items():
        if contacts is not None:
            assert g["contacts_touching_points"][step] == [contacts, touching, points], f"{name} step {step}"
        assert g["state_hash"][step] == sh, f"{name} step {step}: state hash"
        if ch is not None:
            assert g["contact_order_hash"][step] == ch, f"{name} step {step}: contact-order hash"


def test_hello_world_reference_defaults(gold_worlds):
    """BASELINE config 1 with the reference's default flags (continuous physics on): Appendix D values."""
    rows = gold_worlds["hello_world_default_flags_xya_bits"]
    assert len(rows) == 60
    for step, ybits in G.APPENDIX_D_HELLO_Y.items():
        assert rows[step][1] == ybits, f"HelloWorld step {step}: y bits {rows[step][1]:08x} vs {ybits:08x}"


def test_live_reference_matches_golden(oracles, gold_worlds):
    """When the reference library is present (build container), it still produces the frozen answers."""
    if not oracles.available("ref"):
        pytest.skip("oracle/_ref/libb2ref.so not present on this box")
    for name in ("hello_world_default", "pyramid7_jitter3", "mixed_seed5"):
        g = gold_worlds[name]
        for step, bodies, ints, fb, counts in _run_world(oracles.RefWorld, name):
            assert "%016x" % G.state_hash(bodies) == g["state_hash"][step], f"{name} step {step}"
            assert "%016x" % G.contact_order_hash(ints, fb) == g["contact_order_hash"][step], f"{name} step {step}"


def test_port_equals_reference_every_float(oracles):
    """Port vs live reference, all dumps bit-for-bit (bodies, manifolds with impulses, fat AABBs, events)."""
    if not oracles.available("ref"):
        pytest.skip("oracle/_ref/libb2ref.so not present on this box")
    from box2dpp_b200 import scene as S
    from parity_util import compare_world
    scenes = [S.pyramid(6, flags=G.NOCCD, jitter_seed=9), S.mixed(2, flags=G.NOCCD), S.hello_world(G.NOCCD)]
    for si, sc in enumerate(scenes):
        r, p = oracles.RefWorld(sc, record_events=True), oracles.PortWorld(sc, record_events=True)
        for step in range(150):
            r.step(1.0 / 60.0, 8, 3)
            p.step(1.0 / 60.0, 8, 3)
            if step % 3 == 0:
                compare_world(r, p.bodies(), p.contacts(), p.proxies()[0], p.events(), what=f"scene {si} step {step}")


def test_reference_gear_and_mouse_joints_behave(oracles):
    """Semantic check of the oracle harness for the two joint types added last (gear, mouse): the gear constraint
    coordinate1 + ratio * coordinate2 stays at its initial value (b2GearJoint.cpp:287-383) and the mouse joint holds
    its crate near the target — i.e. the scene builder and ref_harness.cpp wire the defs the way the Testbed does."""
    if not oracles.available("ref"):
        pytest.skip("needs the live reference (oracle/_ref)")
    s = S.Scene()
    g = s.create_body(S.STATIC_BODY)
    s.create_box(g, 30.0, 1.0, center=(0.0, -1.0), angle=0.0)
    w1 = s.create_body(S.DYNAMIC_BODY, (-4.0, 6.0), angular_velocity=3.0)
    s.create_circle(w1, 0.5, density=5.0)
    j1 = s.create_revolute_joint(g, w1, (-4.0, 6.0))
    w2 = s.create_body(S.DYNAMIC_BODY, (-2.5, 6.0))
    s.create_circle(w2, 1.0, density=5.0)
    j2 = s.create_revolute_joint(g, w2, (-2.5, 6.0))
    ratio = 2.0
    s.create_gear_joint(j1, j2, ratio=ratio)
    crate = s.create_body(S.DYNAMIC_BODY, (5.0, 4.0))
    s.create_box(crate, 0.4, 0.4, density=1.0)
    s.create_mouse_joint(g, crate, (5.0, 4.0), max_force=1000.0)
    ref = oracles.RefWorld(s)
    for _ in range(120):
        ref.step(1.0 / 60.0, 8, 3)
    b = ref.bodies()
    a1, a2 = float(b[w1, 6]), float(b[w2, 6])
    assert abs(a1) > 0.5, "the driven wheel should have turned"
    assert abs(a1 + ratio * a2) < 0.05, (a1, a2)
    assert abs(float(b[crate, 0]) - 5.0) < 0.1 and abs(float(b[crate, 1]) - 4.0) < 0.3, b[crate, :2]


def test_reference_harness_edit_calls(oracles):
    """The oracle harness's running-world edits (what the GPU parity tests replay against): each call reaches the reference's
    own function and the harness's index bookkeeping (bodies, fixtures, joints by creation order) holds."""
    if not oracles.available("ref"):
        pytest.skip("oracle/_ref not built")
    from box2dpp_b200 import batch as B
    s = S.pyramid(3, flags=15 & ~4)
    r = oracles.RefWorld(s, record_events=True)
    for _ in range(60):
        r.step(1 / 60, 8, 3)
    n_contacts = r.contacts()[0].shape[0]
    assert n_contacts > 0
    top = s.body_count - 1
    # b2Body::SetType(static): velocity cleared, contacts of the body destroyed
    r.command(top, B.CMD_SET_TYPE, 0)
    assert r.contacts()[0].shape[0] < n_contacts and (r.bodies()[top, 7:10] == 0).all()
    r.command(top, B.CMD_SET_TYPE, 2)
    # b2Body::SetMassData / ResetMassData: GetMass follows
    r.command(1, B.CMD_SET_MASS_DATA, 0, (3.0, 2.0), (0.0, 0.0))
    assert r.bodies()[1, 11] == np.float32(3.0)
    r.command(1, B.CMD_RESET_MASS_DATA)
    assert r.bodies()[1, 11] != np.float32(3.0)
    # b2Body::CreateFixture on an existing body, then DestroyFixture: fixture indices keep counting, proxies come and go
    tmp = S.Scene()
    b = tmp.create_body(S.DYNAMIC_BODY, (0.0, 0.0))
    tmp.create_circle(b, 0.3, density=1.0)
    proxies = (r.proxies()[1] >= 0).sum()
    fx = r.create_fixture(2, tmp.fixtures[0], tmp.vertices)
    assert fx == len(s.fixtures) and (r.proxies()[1] >= 0).sum() == proxies + 1
    r.command(2, B.CMD_DESTROY_FIXTURE, fx)
    assert (r.proxies()[1] >= 0).sum() == proxies
    # b2Fixture::SetDensity: GetMassData follows at once, the body's mass at the next ResetMassData
    m_before = r.bodies()[3, 11]
    fx3 = [i for i, f in enumerate(s.fixtures) if int(f.body) == 3][0]
    r.set_density(fx3, 4.0)
    assert r.bodies()[3, 11] == m_before and r.fixture_mass_data(fx3)[0] != m_before
    r.command(3, B.CMD_RESET_MASS_DATA)
    assert r.bodies()[3, 11] == r.fixture_mass_data(fx3)[0] != m_before
    # b2World::CreateJoint / DestroyJoint: ids by creation order
    jd = S.JointDef()
    jd.type, jd.bodyA, jd.bodyB, jd.length, jd.frequencyHz, jd.dampingRatio = S.JOINT_DISTANCE, 1, 2, 1.5, 3.0, 0.5
    assert r.create_joint(jd) == 0 and r.create_joint(jd) == 1
    r.command(0, B.CMD_DESTROY_JOINT, 0)
    assert r.create_joint(jd) == 2
    # b2Fixture::SetFilterData: a fixture that collides with nothing loses its contacts at the next Step
    r.set_filter(1, 1, 0, 0)
    before = r.contacts()[0]
    r.step(1 / 60, 8, 3)
    after = r.contacts()[0]
    assert (before[:, [0, 2]] == 1).any() and not (after[:, [0, 2]] == 1).any()
    # b2Body::SetActive(false): the body's proxies leave the broad phase
    live = (r.proxies()[1] >= 0).sum()
    r.command(3, B.CMD_SET_ACTIVE, 0)
    assert (r.proxies()[1] >= 0).sum() == live - 1
    for _ in range(30):
        r.step(1 / 60, 8, 3)
    assert np.isfinite(r.bodies()).all()
