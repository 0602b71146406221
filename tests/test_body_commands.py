"""Between-step body edits: b2gpu_batch_body_commands / read_body_state against the reference.

A seeded script of b2Body calls (ApplyForce, ApplyLinearImpulse, SetLinearVelocity, SetTransform, SetAwake ...)
is replayed between steps on the vanilla reference, on the C restatement and on the device; all state must stay
bit-identical.  The CPU test pins the restatement to the reference; the GPU test is the parity test proper.
"""
import numpy as np
import pytest

from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S
from parity_util import NOCCD, compare_world, same_bits


def make_script(seed, nbodies, steps, per_step=3):
    """steps x [(body, op, wake, v, p)] from LCG(seed); bodies 0..nbodies-1 (0 is usually static: also exercised)."""
    rng = S.LCG(seed)
    script = []
    for _ in range(steps):
        cmds = []
        for _ in range(per_step):
            body = int(rng.next_u() * nbodies) % nbodies
            op = int(rng.next_u() * 9) % 9
            wake = 1 if rng.next_u() < 0.7 else 0
            if op == B.CMD_SET_TRANSFORM:
                if rng.next_u() < 0.6:
                    op = B.CMD_APPLY_LINEAR_IMPULSE  # keep teleports rarer than pushes
                v = (rng.uniform(-6.0, 6.0), rng.uniform(1.0, 12.0))
                p = (rng.uniform(-3.0, 3.0), 0.0)
            elif op in (B.CMD_APPLY_FORCE, B.CMD_APPLY_FORCE_TO_CENTER):
                v = (rng.uniform(-400.0, 400.0), rng.uniform(-100.0, 600.0))
                p = (rng.uniform(-8.0, 8.0), rng.uniform(0.0, 12.0))
            elif op in (B.CMD_APPLY_TORQUE, B.CMD_APPLY_ANGULAR_IMPULSE, B.CMD_SET_ANGULAR_VELOCITY):
                v = (rng.uniform(-20.0, 20.0) if rng.next_u() < 0.9 else 0.0, 0.0)
                p = (0.0, 0.0)
            else:
                v = (rng.uniform(-8.0, 8.0), rng.uniform(-4.0, 10.0)) if rng.next_u() < 0.9 else (0.0, 0.0)
                p = (rng.uniform(-8.0, 8.0), rng.uniform(0.0, 12.0))
            if op == B.CMD_SET_AWAKE and rng.next_u() < 0.5:
                wake = 0
            f32 = lambda t: (float(np.float32(t[0])), float(np.float32(t[1])))  # noqa: E731
            cmds.append((body, op, wake, f32(v), f32(p)))
        script.append(cmds)
    return script


def scenes_for_commands(flags):
    return [S.pyramid(5, flags=flags, jitter_seed=1), S.mixed(7, flags=flags, bodies=12), S.hello_world(flags=flags),
            S.pyramid(4, flags=flags, jitter_seed=9)]


def test_port_matches_reference_under_commands(oracles):
    if not oracles.available("ref"):
        pytest.skip("oracle/_ref not built")
    for si, scene in enumerate(scenes_for_commands(NOCCD)):
        ref, port = oracles.RefWorld(scene, record_events=True), oracles.PortWorld(scene, record_events=True)
        script = make_script(100 + si, scene.body_count, 120)
        for step, cmds in enumerate(script):
            for (body, op, wake, v, p) in cmds:
                ref.command(body, op, wake, v, p)
                port.command(body, op, wake, v, p)
            ref.step(1 / 60, 8, 3)
            port.step(1 / 60, 8, 3)
            compare_world(ref, port.bodies(), port.contacts(), port.proxies()[0], port.events(), what=f"scene {si} step {step}")


@pytest.mark.gpu
@pytest.mark.parametrize("continuous", [False, True])
def test_gpu_matches_reference_under_commands(oracles, continuous):
    flags = S.WORLD_DEFAULT if continuous else NOCCD
    if continuous and not oracles.available("ref"):
        pytest.skip("TOI needs the live reference")
    Ref = oracles.RefWorld if oracles.available("ref") else oracles.PortWorld
    scenes = scenes_for_commands(flags) * 9  # 36 worlds: two tiles, the second one ragged
    refs = [Ref(s, record_events=True) for s in scenes]
    scripts = [make_script(500 + w, s.body_count, 150) for w, s in enumerate(scenes)]
    batch = B.WorldBatch(scenes, event_mask=B.EVENT_ALL)
    for step in range(150):
        rows = []
        for w, r in enumerate(refs):
            for (body, op, wake, v, p) in scripts[w][step]:
                r.command(body, op, wake, v, p)
                rows.append((w, body, op, wake, v, p))
        if step % 2:  # any world order is accepted: descending worlds, the caller's order kept inside each world
            rows = [t for w in sorted({t[0] for t in rows}, reverse=True) for t in rows if t[0] == w]
        batch.body_commands(np.array(rows, dtype=B.BODY_COMMAND_DTYPE))
        for r in refs:
            r.step(1 / 60, 8, 3)
        batch.step(1 / 60, 8, 3)
        state = batch.read_body_state()
        for w, r in enumerate(refs):
            compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} step {step}")
            rb = r.bodies()
            nb = rb.shape[0]
            assert same_bits(state[w, :nb, 0:2], rb[:, 0:2]) and same_bits(state[w, :nb, 2], rb[:, 6])
            assert same_bits(state[w, :nb, 4:7], rb[:, 7:10]) and (state[w, :nb, 3] == rb[:, 10]).all()
    assert (batch.status() == 0).all()
    batch.close()


@pytest.mark.gpu
def test_bad_command_sets_status():
    batch = B.WorldBatch([S.hello_world()] * 3)
    cmds = np.zeros(2, B.BODY_COMMAND_DTYPE)
    cmds[0] = (1, 99, B.CMD_SET_AWAKE, 1, (0, 0), (0, 0))
    cmds[1] = (2, 1, B.CMD_SET_LINEAR_VELOCITY, 1, (1, 2), (0, 0))
    batch.body_commands(cmds)
    st = batch.status()
    assert st[1] == 4 and st[0] == 0 and st[2] == 0
    assert batch.read_body_state()[2, 1, 4:6].tolist() == [1.0, 2.0]
    with pytest.raises(B.B2GpuError):
        cmds[0]["world"] = 7
        batch.body_commands(cmds)
    batch.close()


@pytest.mark.gpu
def test_mouse_joint_set_target_matches_reference(oracles):
    """b2MouseJoint::SetTarget between steps (B2GPU_CMD_MOUSE_SET_TARGET): a crate on the ground dragged along a path by a
    mouse joint, parked for a while, then dragged again; bit-identical to the reference."""
    if not oracles.available("ref"):
        pytest.skip("checked against the live reference")
    scenes = []
    for w in range(6):
        s = S.Scene()
        g = s.create_body(S.STATIC_BODY)
        s.create_box(g, 30.0, 1.0, center=(0.0, -1.0), angle=0.0)
        crate = s.create_body(S.DYNAMIC_BODY, (0.5 * w, 0.5))
        s.create_box(crate, 0.5, 0.5, density=1.0, friction=0.4)
        other = s.create_body(S.DYNAMIC_BODY, (0.5 * w + 2.0, 0.5))
        s.create_box(other, 0.5, 0.5, density=2.0)
        s.create_mouse_joint(g, crate, (0.5 * w + 0.2, 0.7), max_force=200.0 + 50.0 * w)
        scenes.append(s)
    refs = [oracles.RefWorld(s) for s in scenes]
    batch = B.WorldBatch(scenes)
    for step in range(420):
        rows = []
        if step < 120 or step >= 330:
            if step % 3 == 0:
                for w, r in enumerate(refs):
                    t = np.float32(step) * np.float32(0.05)
                    target = (float(np.float32(0.5 * w) + np.float32(3.0) * np.sin(t)), float(np.float32(1.5) + np.cos(t)))
                    r.command(1, B.CMD_MOUSE_SET_TARGET, 1, target)
                    rows.append((w, 1, B.CMD_MOUSE_SET_TARGET, 1, target, (0, 0)))
        if step == 120:  # park it near the ground for 200 steps
            for w, r in enumerate(refs):
                target = (0.5 * w + 1.0, 0.45)
                r.command(1, B.CMD_MOUSE_SET_TARGET, 1, target)
                rows.append((w, 1, B.CMD_MOUSE_SET_TARGET, 1, target, (0, 0)))
        if rows:
            batch.body_commands(np.array(rows, dtype=B.BODY_COMMAND_DTYPE))
        for r in refs:
            r.step(1.0 / 60.0, 8, 3)
        batch.step(1.0 / 60.0, 8, 3)
        if step % 30 == 29 or step == 329:
            for w, r in enumerate(refs):
                compare_world(r, batch.bodies(w), batch.contacts(w), what=f"world {w} step {step}")
    assert (batch.status() == 0).all()
    # a body without a mouse joint: BAD_COMMAND
    batch.body_commands(np.array([(2, 2, B.CMD_MOUSE_SET_TARGET, 1, (0, 0), (0, 0))], dtype=B.BODY_COMMAND_DTYPE))
    assert batch.status()[2] == 4
    batch.close()


@pytest.mark.gpu
def test_destroy_body_matches_reference(oracles):
    """b2World::DestroyBody between steps (B2GPU_CMD_DESTROY_BODY): the body's joints, contacts (EndContact events, wake-ups
    of the neighbours) and proxies go with it; everything that remains stays bit-identical to the reference."""
    if not oracles.available("ref"):
        pytest.skip("checked against the live reference")
    scenes = [S.pyramid(6), S.pyramid(5, flags=NOCCD, jitter_seed=3), S.jointed(1), S.jointed(2, flags=NOCCD), S.vehicles(1),
              S.mixed(3), S.sensors(2), S.bullets(1)] * 5  # 40 worlds
    refs = [oracles.RefWorld(s, record_events=True) for s in scenes]
    batch = B.WorldBatch(scenes, event_mask=B.EVENT_ALL)
    rng = S.LCG(77)
    alive = [list(range(s.body_count)) for s in scenes]
    for step in range(260):
        rows = []
        if step % 9 == 4:
            for w, r in enumerate(refs):
                if len(alive[w]) <= 2 or rng.next_u() < 0.3:
                    continue
                # mostly dynamic bodies, now and then body 0 (the static ground / anchor of the joints)
                pick = 0 if (rng.next_u() < 0.05 and 0 in alive[w]) else alive[w][1 + int(rng.next_u() * (len(alive[w]) - 1)) % (len(alive[w]) - 1)]
                alive[w].remove(pick)
                r.command(pick, B.CMD_DESTROY_BODY)
                rows.append((w, pick, B.CMD_DESTROY_BODY, 0, (0, 0), (0, 0)))
                if rng.next_u() < 0.5 and len(alive[w]) > 1:  # poke a survivor in the same batch of commands
                    other = alive[w][-1]
                    r.command(other, B.CMD_APPLY_LINEAR_IMPULSE, 1, (1.0, 3.0), (0.0, 2.0))
                    rows.append((w, other, B.CMD_APPLY_LINEAR_IMPULSE, 1, (1.0, 3.0), (0.0, 2.0)))
            batch.body_commands(np.array(rows, dtype=B.BODY_COMMAND_DTYPE))
            for w, r in enumerate(refs):  # EndContact events of the destruction follow the last step's events
                compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} after destroy at step {step}")
        for r in refs:
            r.step(1 / 60, 8, 3)
        batch.step(1 / 60, 8, 3)
        if step % 3 == 0 or step % 9 == 4:
            for w, r in enumerate(refs):
                compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} step {step}")
    assert (batch.status() == 0).all()
    # a command on a destroyed body is reported, not executed
    w = next(i for i in range(len(scenes)) if len(alive[i]) < scenes[i].body_count)
    dead = next(b for b in range(scenes[w].body_count) if b not in alive[w])
    batch.body_commands(np.array([(w, dead, B.CMD_SET_AWAKE, 1, (0, 0), (0, 0))], dtype=B.BODY_COMMAND_DTYPE))
    assert batch.status()[w] == 4
    batch.close()


def _spawn_def(rng, kind):
    """(BodyDef, [FixtureDef], vertices) of one body to create mid-run, built through the Scene helpers."""
    tmp = S.Scene()
    b = tmp.create_body(S.DYNAMIC_BODY, (rng.uniform(-6.0, 6.0), rng.uniform(6.0, 11.0)), angle=rng.uniform(-1.0, 1.0),
                        linear_velocity=(rng.uniform(-2.0, 2.0), rng.uniform(-3.0, 0.0)), angular_velocity=rng.uniform(-2.0, 2.0),
                        flags=S.BD_DEFAULT | (S.BD_BULLET if kind == 3 else 0))
    if kind == 0:
        tmp.create_circle(b, rng.uniform(0.2, 0.5), density=1.0, restitution=0.1)
    elif kind == 1:
        tmp.create_box(b, rng.uniform(0.2, 0.5), rng.uniform(0.2, 0.5), density=2.0, friction=0.4)
    elif kind == 2:  # two fixtures: a convex polygon and a sensor halo
        n = 5
        r = rng.uniform(0.3, 0.6)
        tmp.create_polygon(b, [(float(np.float32(r * np.cos(2 * np.pi * j / n))), float(np.float32(r * np.sin(2 * np.pi * j / n)))) for j in range(n)], density=1.0)
        tmp.create_circle(b, 1.0, sensor=True)
    else:
        tmp.create_box(b, 0.15, 0.15, density=6.0)
    return tmp.bodies[0], list(tmp.fixtures), list(tmp.vertices)


@pytest.mark.gpu
def test_create_and_destroy_bodies_match_reference(oracles):
    """b2World::CreateBody / CreateFixture and DestroyBody between steps (b2gpu_batch_create_bodies, B2GPU_CMD_DESTROY_BODY):
    body indices, reused proxy ids (= pair order), e_newFixture's FindNewContacts at the next Step, mass data of the new
    bodies — everything stays bit-identical to the reference."""
    if not oracles.available("ref"):
        pytest.skip("checked against the live reference")
    scenes = [S.pyramid(4), S.mixed(2, bodies=10), S.jointed(1), S.hello_world(), S.pyramid(3, flags=NOCCD, jitter_seed=2),
              S.sensors(1, bodies=6)] * 6  # 36 worlds
    refs = [oracles.RefWorld(s, record_events=True) for s in scenes]
    batch = B.WorldBatch(scenes, event_mask=B.EVENT_ALL, headroom=(24, 40, 40, 800, 0))
    rng = S.LCG(4242)
    alive = [list(range(s.body_count)) for s in scenes]
    created = 0
    for step in range(240):
        if step % 7 == 3:
            spawns, destroys = [], []
            for w, r in enumerate(refs):
                u = rng.next_u()
                if u < 0.45:  # create one or two bodies
                    for _ in range(1 + int(rng.next_u() * 2) % 2):
                        bd, fds, verts = _spawn_def(rng, int(rng.next_u() * 4) % 4)
                        idx = r.create_body(bd, fds, verts)
                        spawns.append((w, bd, fds, verts, idx))
                elif u < 0.75 and len(alive[w]) > 2:  # destroy a dynamic body (possibly one that was created mid-run)
                    pick = alive[w][1 + int(rng.next_u() * (len(alive[w]) - 1)) % (len(alive[w]) - 1)]
                    alive[w].remove(pick)
                    r.command(pick, B.CMD_DESTROY_BODY)
                    destroys.append((w, pick, B.CMD_DESTROY_BODY, 0, (0, 0), (0, 0)))
            if destroys:
                batch.body_commands(np.array(destroys, dtype=B.BODY_COMMAND_DTYPE))
            if spawns:
                got = batch.create_bodies([(w, bd, fds, verts) for (w, bd, fds, verts, _) in spawns])
                assert got.tolist() == [idx for (*_, idx) in spawns]
                for (w, *_, idx) in spawns:
                    alive[w].append(idx)
                created += len(spawns)
            for w, r in enumerate(refs):
                compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} after edits at step {step}")
        for r in refs:
            r.step(1 / 60, 8, 3)
        batch.step(1 / 60, 8, 3)
        for w, r in enumerate(refs):
            compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} step {step}")
    assert created > 150 and (batch.status() == 0).all()
    # a world without headroom reports the overflow instead of corrupting its arenas
    small = B.WorldBatch([S.hello_world()])
    bd, fds, verts = _spawn_def(S.LCG(1), 1)
    assert small.create_bodies([(0, bd, fds, verts)]).tolist() == [-1] and small.status()[0] == 8
    small.close()
    batch.close()


def _edit_fixture(rng, kind):
    """[FixtureDef], vertices of one fixture to attach to an existing body."""
    tmp = S.Scene()
    b = tmp.create_body(S.DYNAMIC_BODY, (0.0, 0.0))
    if kind == 0:
        tmp.create_circle(b, rng.uniform(0.15, 0.4), density=rng.uniform(0.5, 2.0), restitution=0.2)
    elif kind == 1:
        tmp.create_box(b, rng.uniform(0.15, 0.4), rng.uniform(0.15, 0.4), density=1.5, friction=0.5)
    elif kind == 2:
        tmp.create_circle(b, rng.uniform(0.5, 0.9), sensor=True)  # no density: the mass data stays
    else:
        n = 4
        r = rng.uniform(0.25, 0.5)
        tmp.create_polygon(b, [(float(np.float32(r * np.cos(2 * np.pi * j / n + 0.3))), float(np.float32(r * np.sin(2 * np.pi * j / n + 0.3)))) for j in range(n)], density=1.0)
    return list(tmp.fixtures), list(tmp.vertices)


def _edit_joint(rng, body_a, body_b):
    jd = S.JointDef()
    kind = int(rng.next_u() * 6) % 6
    jd.bodyA, jd.bodyB = body_a, body_b
    jd.collideConnected = 1 if rng.next_u() < 0.4 else 0
    jd.localAnchorA[0], jd.localAnchorA[1] = float(np.float32(rng.uniform(-0.2, 0.2))), float(np.float32(rng.uniform(-0.2, 0.2)))
    jd.localAnchorB[0], jd.localAnchorB[1] = float(np.float32(rng.uniform(-0.2, 0.2))), float(np.float32(rng.uniform(-0.2, 0.2)))
    if kind == 0:
        jd.type, jd.length, jd.frequencyHz, jd.dampingRatio = S.JOINT_DISTANCE, float(np.float32(rng.uniform(1.0, 4.0))), 3.0, 0.6
    elif kind == 1:
        jd.type, jd.length = S.JOINT_ROPE, float(np.float32(rng.uniform(2.0, 6.0)))
    elif kind == 2:
        jd.type, jd.length, jd.maxMotor = S.JOINT_FRICTION, 5.0, 2.0
    elif kind == 3:
        jd.type, jd.frequencyHz, jd.dampingRatio = S.JOINT_WELD, 2.0, 0.7
    elif kind == 4:
        jd.type, jd.lower, jd.upper, jd.enableLimit = S.JOINT_REVOLUTE, -0.5, 0.5, 1
    else:
        jd.type, jd.length = S.JOINT_DISTANCE, float(np.float32(rng.uniform(1.5, 3.0)))
    return jd


@pytest.mark.gpu
def test_world_edits_match_reference(oracles):
    """The rest of §8f-2 against the live reference, random edits between steps, every float and contact compared after
    the edits and after every step: b2Fixture::SetFilterData (contacts flagged, b2ContactManager::Collide's e_filterFlag
    branch), b2World::CreateJoint / DestroyJoint on running worlds (collideConnected flags the pair's contacts),
    b2Body::SetType, SetActive, SetMassData, ResetMassData, DestroyFixture and CreateFixture on existing bodies, and the
    plain setters (damping, gravity scale, bullet, sleeping allowed, fixed rotation, b2Fixture::SetSensor / SetFriction /
    SetRestitution)."""
    if not oracles.available("ref"):
        pytest.skip("checked against the live reference")
    base = [S.pyramid(4), S.mixed(2, bodies=10), S.jointed(1), S.filtered(16), S.hello_world(), S.jointed(4, flags=NOCCD),
            S.pyramid(3, flags=NOCCD, jitter_seed=2), S.sensors(1, bodies=6), S.vehicles(2, flags=NOCCD)]
    scenes = base * 4  # 36 worlds (the vehicles scene brings every joint type, a mouse joint's m_bodyB->GetMass() included)
    refs = [oracles.RefWorld(s, record_events=True) for s in scenes]
    batch = B.WorldBatch(scenes, event_mask=B.EVENT_ALL, headroom=(0, 60, 90, 1500, 0), joint_headroom=24)
    rng = S.LCG(977)
    bodies = [list(range(s.body_count)) for s in scenes]
    fixtures = [{i: int(f.body) for i, f in enumerate(s.fixtures)} for s in scenes]          # live fixture -> body
    joints = [set(range(len(s.joints))) for s in scenes]
    # The reference's contract for gear joints (b2GearJoint.h:53 "you have to manually destroy the gear joint if joint1 or
    # joint2 is destroyed"; its solver reads bodyC / bodyD through their island index, which is stale when the ground is not
    # in the island): the joints a gear was built from stay, and the gears' ground body stays active and static.
    geared = [w for w, s in enumerate(scenes) if any(int(j.type) == S.JOINT_GEAR for j in s.joints)]
    for w in geared:
        for j in scenes[w].joints:
            if int(j.type) == S.JOINT_GEAR:
                joints[w] -= {int(j.enableLimit), int(j.enableMotor)}
    # (and a mouse joint's body stays dynamic: b2MouseJoint::InitVelocityConstraints asserts on a body without mass)
    mouse_bodies = [{int(j.bodyB) for j in s.joints if int(j.type) == S.JOINT_MOUSE} for s in scenes]
    types = [[int(b.type) for b in s.bodies] for s in scenes]
    counts = {}
    pick = lambda seq: seq[int(rng.next_u() * len(seq)) % len(seq)]  # noqa: E731
    for step in range(260):
        if step % 5 == 2:
            cmds, attach, new_joints, filters = [], [], [], {}
            for w, r in enumerate(refs):
                for _ in range(1 + int(rng.next_u() * 2) % 2):
                    u = rng.next_u()
                    body = pick(bodies[w])
                    if body == 0 and w in geared and 0.16 <= u < 0.40:
                        continue
                    if u < 0.16 and fixtures[w]:
                        fx = pick(sorted(fixtures[w]))
                        filters[len(cmds)] = (1 << (int(rng.next_u() * 4) % 4), 0xffff if rng.next_u() < 0.5 else int(rng.next_u() * 16) % 16,
                                              int(rng.next_u() * 5) % 5 - 2)
                        cmds.append((w, fixtures[w][fx], B.CMD_SET_FILTER, fx, (0.0, 0.0), (0.0, 0.0)))
                    elif u < 0.28:
                        t = int(rng.next_u() * 3) % 3
                        if (body == 0 and rng.next_u() < 0.8) or body in mouse_bodies[w]:
                            continue  # the ground mostly stays what it is
                        cmds.append((w, body, B.CMD_SET_TYPE, t, (0.0, 0.0), (0.0, 0.0)))
                        types[w][body] = t
                    elif u < 0.40:
                        cmds.append((w, body, B.CMD_SET_ACTIVE, 1 if rng.next_u() < 0.6 else 0, (0.0, 0.0), (0.0, 0.0)))
                    elif u < 0.50:
                        m = float(np.float32(rng.uniform(0.5, 5.0)))
                        cmds.append((w, body, B.CMD_SET_MASS_DATA, 0, (m, float(np.float32(rng.uniform(0.5, 3.0)))),
                                     (float(np.float32(rng.uniform(-0.1, 0.1))), float(np.float32(rng.uniform(-0.1, 0.1))))))
                    elif u < 0.56:
                        cmds.append((w, body, B.CMD_RESET_MASS_DATA, 0, (0.0, 0.0), (0.0, 0.0)))
                    elif u < 0.66 and fixtures[w]:
                        fx = pick(sorted(fixtures[w]))
                        cmds.append((w, fixtures[w].pop(fx), B.CMD_DESTROY_FIXTURE, fx, (0.0, 0.0), (0.0, 0.0)))
                    elif u < 0.80:
                        fds, verts = _edit_fixture(rng, int(rng.next_u() * 4) % 4)
                        attach.append((w, body, fds, verts))
                    elif u < 0.86 and len(bodies[w]) > 1:
                        other = pick([b for b in bodies[w] if b != body])
                        new_joints.append((w, _edit_joint(rng, body, other)))
                    elif u < 0.93:  # the plain setters
                        k = int(rng.next_u() * 7) % 7
                        f32 = lambda x: float(np.float32(x))  # noqa: E731
                        if k == 0:
                            cmds.append((w, body, B.CMD_SET_DAMPING, 0, (f32(rng.uniform(0.0, 0.5)), f32(rng.uniform(0.0, 0.5))), (0.0, 0.0)))
                        elif k == 1:
                            cmds.append((w, body, B.CMD_SET_GRAVITY_SCALE, 0, (f32(rng.uniform(-0.5, 2.0)), 0.0), (0.0, 0.0)))
                        elif k == 2:
                            cmds.append((w, body, B.CMD_SET_BULLET, 1 if rng.next_u() < 0.5 else 0, (0.0, 0.0), (0.0, 0.0)))
                        elif k == 3:
                            cmds.append((w, body, B.CMD_SET_SLEEPING_ALLOWED, 1 if rng.next_u() < 0.5 else 0, (0.0, 0.0), (0.0, 0.0)))
                        elif k == 4:
                            cmds.append((w, body, B.CMD_SET_FIXED_ROTATION, 1 if rng.next_u() < 0.5 else 0, (0.0, 0.0), (0.0, 0.0)))
                        elif fixtures[w]:
                            fx = pick(sorted(fixtures[w]))
                            if k == 5:
                                cmds.append((w, fixtures[w][fx], B.CMD_SET_SENSOR, fx, (1.0 if rng.next_u() < 0.4 else 0.0, 0.0), (0.0, 0.0)))
                            else:
                                cmds.append((w, fixtures[w][fx], B.CMD_SET_MATERIAL, fx, (f32(rng.uniform(0.0, 1.0)), f32(rng.uniform(0.0, 0.6))), (0.0, 0.0)))
                    elif joints[w]:
                        j = pick(sorted(joints[w]))
                        joints[w].discard(j)
                        cmds.append((w, 0, B.CMD_DESTROY_JOINT, j, (0.0, 0.0), (0.0, 0.0)))
            # the reference takes the calls in the order the device applies them: commands, then fixtures, then joints
            for i, (w, body, op, wake, v, p) in enumerate(cmds):
                if op == B.CMD_SET_FILTER:
                    refs[w].set_filter(wake, *filters[i])
                else:
                    refs[w].command(body, op, wake, v, p)
                counts[op] = counts.get(op, 0) + 1
            if cmds:
                arr = np.array(cmds, dtype=B.BODY_COMMAND_DTYPE)
                for i, f in filters.items():
                    B.put_filter(arr, i, *f)
                batch.body_commands(arr)
            if attach:
                want = [refs[w].create_fixture(body, fds[0], verts) for (w, body, fds, verts) in attach]
                got = batch.create_fixtures(attach)
                assert got.tolist() == want
                for (w, body, *_), fx in zip(attach, want):
                    fixtures[w][fx] = body
                counts["fixture"] = counts.get("fixture", 0) + len(attach)
            if new_joints:
                want = [refs[w].create_joint(jd) for (w, jd) in new_joints]
                got = batch.create_joints(new_joints)
                assert got.tolist() == want
                for (w, _), j in zip(new_joints, want):
                    joints[w].add(j)
                counts["joint"] = counts.get("joint", 0) + len(new_joints)
            for w, r in enumerate(refs):
                compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} after edits at step {step}")
        for r in refs:
            r.step(1 / 60, 8, 3)
        batch.step(1 / 60, 8, 3)
        for w, r in enumerate(refs):
            compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} step {step}")
    assert (batch.status() == 0).all(), batch.status()
    assert all(counts.get(k, 0) > 20 for k in (B.CMD_SET_FILTER, B.CMD_SET_TYPE, B.CMD_SET_ACTIVE, B.CMD_SET_MASS_DATA,
                                               B.CMD_DESTROY_FIXTURE, "fixture", "joint")), counts
    assert counts.get(B.CMD_DESTROY_JOINT, 0) > 5, counts
    assert sum(counts.get(k, 0) for k in range(B.CMD_SET_DAMPING, B.CMD_SET_MATERIAL + 1)) > 30, counts
    batch.close()


@pytest.mark.gpu
def test_set_density_matches_reference(oracles):
    """b2Fixture::SetDensity on running worlds (B2GPU_CMD_SET_DENSITY carries the fixture's mass data under the new density,
    computed by the host layer): like the reference's call it changes nothing until b2Body::ResetMassData — sometimes issued
    at once, sometimes steps later, sometimes as a side effect of another edit (SetType) — and from then on every float of the
    world agrees with the live reference; b2Body::GetMass (B2GPU_F_BODY_DAMP .w) is compared as well."""
    if not oracles.available("ref"):
        pytest.skip("checked against the live reference")
    scenes = [S.pyramid(4), S.mixed(3, bodies=10), S.hello_world(), S.sensors(1, bodies=6), S.pyramid(3, flags=NOCCD, jitter_seed=5),
              S.mixed(11, flags=NOCCD, bodies=8)] * 2
    refs = [oracles.RefWorld(s, record_events=True) for s in scenes]
    batch = B.WorldBatch(scenes, event_mask=B.EVENT_ALL)
    rng = S.LCG(4242)
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    edits = resets = 0
    mass_bad = []
    for step in range(140):
        if step % 4 == 1:
            cmds = []
            for w, (s, r) in enumerate(zip(scenes, refs)):
                fx = int(rng.next_u() * len(s.fixtures)) % len(s.fixtures)
                body = int(s.fixtures[fx].body)
                density = 0.0 if rng.next_u() < 0.2 else f32(rng.uniform(0.2, 4.0))
                md = B.host_fixture_mass_data(s.fixtures[fx], s.vertices, density)
                r.set_density(fx, density)
                assert r.fixture_mass_data(fx).view(np.uint32).tolist() == md.view(np.uint32).tolist()
                cmds.append((w, body, B.CMD_SET_DENSITY, fx, (float(md[0]), float(md[1])), (float(md[2]), float(md[3]))))
                edits += 1
                u = rng.next_u()
                if u < 0.6:
                    other = int(rng.next_u() * s.body_count) % s.body_count  # any body: pending densities of earlier edits tell now
                    for b in {body, other}:
                        r.command(b, B.CMD_RESET_MASS_DATA)
                        cmds.append((w, b, B.CMD_RESET_MASS_DATA, 0, (0.0, 0.0), (0.0, 0.0)))
                        resets += 1
                elif u < 0.7 and body != 0:  # SetType resets the mass data on its way
                    for t in (S.KINEMATIC_BODY, S.DYNAMIC_BODY):
                        r.command(body, B.CMD_SET_TYPE, t)
                        cmds.append((w, body, B.CMD_SET_TYPE, t, (0.0, 0.0), (0.0, 0.0)))
            batch.body_commands(np.array(cmds, dtype=B.BODY_COMMAND_DTYPE))
        for r in refs:
            r.step(1 / 60, 8, 3)
        batch.step(1 / 60, 8, 3)
        for w, r in enumerate(refs):
            compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} step {step}")
            if step % 4 == 1:
                rb = r.bodies()
                got = batch.field(B.F_BODY_DAMP, w, 1)[0, :rb.shape[0], 3]
                if not same_bits(rb[:, 11], got):
                    mass_bad.append((step, w, rb[:, 11].tolist(), got.tolist()))
    assert (batch.status() == 0).all(), batch.status()
    assert edits > 300 and resets > 200, (edits, resets)
    assert not mass_bad, mass_bad[:2]
    # a fixture index that is not the body's: refused like the other fixture commands
    bad = np.array([(0, 1, B.CMD_SET_DENSITY, 0, (1.0, 1.0), (0.0, 0.0))], dtype=B.BODY_COMMAND_DTYPE)
    batch.body_commands(bad)
    assert batch.status()[0] & 4  # B2GPU_STATUS_BAD_COMMAND
    batch.close()


def _late_mouse(ground, body, target, max_force, collide):
    jd = S.JointDef()
    jd.type, jd.bodyA, jd.bodyB, jd.collideConnected = S.JOINT_MOUSE, ground, body, collide
    jd.localAnchorA[0], jd.localAnchorA[1] = float(np.float32(target[0])), float(np.float32(target[1]))
    jd.length, jd.frequencyHz, jd.dampingRatio = max_force, 5.0, 0.7
    return jd


def _late_gear(joint1, joint2, ratio, body_a, body_b):
    jd = S.JointDef()
    jd.type, jd.bodyA, jd.bodyB = S.JOINT_GEAR, body_a, body_b
    jd.enableLimit, jd.enableMotor, jd.motorSpeed = joint1, joint2, ratio
    return jd


@pytest.mark.gpu
def test_gear_and_mouse_joints_on_running_worlds(oracles):
    """b2World::CreateJoint of the two joint classes whose constructors read the world (b2GearJoint.cpp:47-133: joint1 / joint2
    and their bodies' poses; b2MouseJoint.cpp:33-52: body B's transform) on worlds that have been stepping: the gears and the
    mouse joint of the vehicles scene are created at step 20 instead of with the scene, together with a brand-new revolute
    joint and a gear built on it in the same call; mouse joints grab boxes of running pyramids (collideConnected on and off);
    b2MouseJoint::SetTarget keeps moving the targets; the late gear and its joint are destroyed again.  Every float of every
    body and contact against the live reference, every step."""
    if not oracles.available("ref"):
        pytest.skip("checked against the live reference")
    scenes, late = [], []
    for seed in (1, 2, 3, 4, 5, 6):
        s = S.vehicles(seed, flags=NOCCD if seed % 2 else S.WORLD_DEFAULT)
        late.append([j for j in s.joints if int(j.type) in (S.JOINT_GEAR, S.JOINT_MOUSE)])
        s.joints = [j for j in s.joints if int(j.type) not in (S.JOINT_GEAR, S.JOINT_MOUSE)]  # the gears' joint1 / joint2 precede them
        scenes.append(s)
    scenes += [S.pyramid(4), S.pyramid(3, flags=NOCCD, jitter_seed=3), S.pyramid(4, jitter_seed=8)]
    refs = [oracles.RefWorld(s, record_events=True) for s in scenes]
    batch = B.WorldBatch(scenes, event_mask=B.EVENT_ALL, joint_headroom=8)
    rng = S.LCG(31337)
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    mice = {}      # world -> body held by a mouse joint
    late_ids = {}  # world -> (gear id, revolute id) of the pair created together
    for step in range(130):
        if step == 20:
            new = []
            for w, s in enumerate(scenes):
                rb = refs[w].bodies()
                if w < 6:
                    n0 = len(s.joints)
                    for jd in late[w]:
                        if int(jd.type) == S.JOINT_MOUSE:  # grabbed where the crate is NOW
                            b = int(jd.bodyB)
                            jd = _late_mouse(int(jd.bodyA), b, (rb[b, 4] + 0.3, rb[b, 5] + 0.3), float(jd.length), 0)
                            mice[w] = b
                        new.append((w, jd))
                    # a revolute joint between the ground and the motor joint's platform, and a gear between it and the first wheel pivot
                    plat = [int(j.bodyB) for j in s.joints if int(j.type) == S.JOINT_MOTOR][0]
                    pivot = [i for i, j in enumerate(s.joints) if int(j.type) == S.JOINT_REVOLUTE][0]
                    rev = S.JointDef()
                    rev.type, rev.bodyA, rev.bodyB = S.JOINT_REVOLUTE, 0, plat
                    rev.localAnchorA[0], rev.localAnchorA[1] = f32(rb[plat, 4]), f32(rb[plat, 5])
                    rev.referenceAngle = f32(rb[plat, 6])
                    new.append((w, rev))
                    rev_id = n0 + len(late[w])
                    new.append((w, _late_gear(rev_id, pivot, 1.5, plat, int(s.joints[pivot].bodyB))))
                    late_ids[w] = (rev_id + 1, rev_id)
                else:
                    b = s.body_count - 1 - (w - 6)
                    new.append((w, _late_mouse(0, b, (rb[b, 4] + 0.2, rb[b, 5] + 0.4), 300.0, w % 2)))
                    mice[w] = b
            new = [x for x in new if x[0] >= 4] + [x for x in new if x[0] < 4]  # worlds out of order; each world's own order is kept
            want = [refs[w].create_joint(jd) for (w, jd) in new]
            got = batch.create_joints(new)
            assert got.tolist() == want and min(want) >= 0, (got.tolist(), want)
        if step > 20 and step % 9 == 0:
            cmds = []
            for w, b in sorted(mice.items()):
                rb = refs[w].bodies()
                t = (f32(rb[b, 4] + rng.uniform(-1.0, 1.0)), f32(rb[b, 5] + rng.uniform(-0.5, 1.5)))
                refs[w].command(b, B.CMD_MOUSE_SET_TARGET, 1, t, (0.0, 0.0))
                cmds.append((w, b, B.CMD_MOUSE_SET_TARGET, 1, t, (0.0, 0.0)))
            batch.body_commands(np.array(cmds, dtype=B.BODY_COMMAND_DTYPE))
        if step == 80:  # the gear first, then the joint it was built on (b2GearJoint.h:53)
            cmds = []
            for w, ids in sorted(late_ids.items()):
                for j in ids:
                    refs[w].command(0, B.CMD_DESTROY_JOINT, j)
                    cmds.append((w, 0, B.CMD_DESTROY_JOINT, j, (0.0, 0.0), (0.0, 0.0)))
            batch.body_commands(np.array(cmds, dtype=B.BODY_COMMAND_DTYPE))
        for r in refs:
            r.step(1 / 60, 8, 3)
        batch.step(1 / 60, 8, 3)
        for w, r in enumerate(refs):
            compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} step {step}")
    assert (batch.status() == 0).all(), batch.status()
    # a gear on a joint that is gone, or that never was, is refused
    with pytest.raises(B.B2GpuError):
        batch.create_joints([(0, _late_gear(late_ids[0][1], 0, 1.0, 1, 2))])
    with pytest.raises(B.B2GpuError):
        batch.create_joints([(0, _late_gear(999, 0, 1.0, 1, 2))])
    batch.close()
