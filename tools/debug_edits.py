"""Diagnostic (GPU): the random-edit parity loop of tests/test_body_commands.py::test_world_edits_match_reference on one scene
type, with a log of the edits per world; prints the edits that preceded the first divergence.
    python tools/debug_edits.py [ops allowed, comma separated: filter,type,active,mass,reset,dfix,cfix,cjoint,djoint]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S
from oracle import worlds as O
from parity_util import NOCCD, compare_world
import test_body_commands as T

allowed = set((sys.argv[1] if len(sys.argv) > 1 else "filter,type,active,mass,reset,dfix,cfix,cjoint,djoint").split(","))
scenes = [S.vehicles(2, flags=NOCCD)] * 6
refs = [O.RefWorld(s, record_events=True) for s in scenes]
batch = B.WorldBatch(scenes, event_mask=B.EVENT_ALL, headroom=(0, 60, 90, 1500, 0), joint_headroom=64)
rng = S.LCG(int(os.environ.get("SEED", "977")))
bodies = [list(range(s.body_count)) for s in scenes]
fixtures = [{i: int(f.body) for i, f in enumerate(s.fixtures)} for s in scenes]
joints = [set(range(len(s.joints))) - {13, 14, 15} for s in scenes]  # joints a gear joint was built from stay (b2GearJoint.h:53)
log = [[] for _ in scenes]
pick = lambda seq: seq[int(rng.next_u() * len(seq)) % len(seq)]
jt = [int(j.type) for j in scenes[0].joints]
print("scene joints (type, A, B):", [(int(j.type), int(j.bodyA), int(j.bodyB)) for j in scenes[0].joints])
for step in range(300):
    if step % 5 == 2:
        cmds, attach, new_joints, filters = [], [], [], {}
        for w, r in enumerate(refs):
            u = rng.next_u()
            body = pick(bodies[w])
            if u < 0.16 and fixtures[w] and "filter" in allowed:
                fx = pick(sorted(fixtures[w]))
                filters[len(cmds)] = (1 << (int(rng.next_u() * 4) % 4), 0xffff if rng.next_u() < 0.5 else int(rng.next_u() * 16) % 16, int(rng.next_u() * 5) % 5 - 2)
                cmds.append((w, fixtures[w][fx], B.CMD_SET_FILTER, fx, (0.0, 0.0), (0.0, 0.0)))
            elif u < 0.40 and body == 0:
                continue  # the ground of this scene is bodyC / bodyD of the gear joints: see the test's note
            elif u < 0.28 and "type" in allowed:
                cmds.append((w, body, B.CMD_SET_TYPE, int(rng.next_u() * 3) % 3, (0.0, 0.0), (0.0, 0.0)))
            elif u < 0.40 and "active" in allowed:
                cmds.append((w, body, B.CMD_SET_ACTIVE, 1 if rng.next_u() < 0.6 else 0, (0.0, 0.0), (0.0, 0.0)))
            elif u < 0.50 and "mass" in allowed:
                cmds.append((w, body, B.CMD_SET_MASS_DATA, 0, (float(np.float32(rng.uniform(0.5, 5.0))), float(np.float32(rng.uniform(0.5, 3.0)))),
                             (float(np.float32(rng.uniform(-0.1, 0.1))), float(np.float32(rng.uniform(-0.1, 0.1))))))
            elif u < 0.56 and "reset" in allowed:
                cmds.append((w, body, B.CMD_RESET_MASS_DATA, 0, (0.0, 0.0), (0.0, 0.0)))
            elif u < 0.66 and fixtures[w] and "dfix" in allowed:
                fx = pick(sorted(fixtures[w]))
                cmds.append((w, fixtures[w].pop(fx), B.CMD_DESTROY_FIXTURE, fx, (0.0, 0.0), (0.0, 0.0)))
            elif u < 0.80 and "cfix" in allowed:
                fds, verts = T._edit_fixture(rng, int(rng.next_u() * 4) % 4)
                attach.append((w, body, fds, verts))
            elif u < 0.92 and "cjoint" in allowed:
                other = pick([b for b in bodies[w] if b != body])
                new_joints.append((w, T._edit_joint(rng, body, other)))
            elif joints[w] and "djoint" in allowed:
                j = pick(sorted(joints[w]))
                joints[w].discard(j)
                cmds.append((w, 0, B.CMD_DESTROY_JOINT, j, (0.0, 0.0), (0.0, 0.0)))
        for i, (w, body, op, wake, v, p) in enumerate(cmds):
            log[w].append((step, "cmd", op, body, wake, v, p, filters.get(i)))
            if op == B.CMD_SET_FILTER:
                refs[w].set_filter(wake, *filters[i])
            else:
                refs[w].command(body, op, wake, v, p)
        if cmds:
            arr = np.array(cmds, dtype=B.BODY_COMMAND_DTYPE)
            for i, f in filters.items():
                B.put_filter(arr, i, *f)
            batch.body_commands(arr)
        if attach:
            want = [refs[w].create_fixture(body, fds[0], verts) for (w, body, fds, verts) in attach]
            got = batch.create_fixtures(attach)
            for (w, body, fds, _), fx in zip(attach, want):
                fixtures[w][fx] = body
                log[w].append((step, "fixture", body, fx, int(fds[0].shapeType), float(fds[0].density)))
            assert got.tolist() == want, (got, want)
        if new_joints:
            want = [refs[w].create_joint(jd) for (w, jd) in new_joints]
            got = batch.create_joints(new_joints)
            for (w, jd), j in zip(new_joints, want):
                joints[w].add(j)
                log[w].append((step, "joint", j, int(jd.type), int(jd.bodyA), int(jd.bodyB), int(jd.collideConnected)))
            assert got.tolist() == want, (got, want)
    for phase in ("edits", "step"):
        if phase == "step":
            for r in refs:
                r.step(1 / 60, 8, 3)
            batch.step(1 / 60, 8, 3)
        elif step % 5 != 2:
            continue
        for w, r in enumerate(refs):
            try:
                compare_world(r, batch.bodies(w), batch.contacts(w), batch.proxies(w), batch.events(w), what=f"world {w} {phase} {step}")
            except AssertionError as e:
                print("FIRST DIVERGENCE:", str(e)[:400])
                for entry in log[w][-8:]:
                    print("   ", entry)
                sys.exit(1)
print("no divergence in 300 steps with ops", sorted(allowed))
