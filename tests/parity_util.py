"""Helpers shared by the parity tests: step two implementations side by side and compare bit patterns."""
import numpy as np

NOCCD = 15 & ~4  # B2G_WORLD_DEFAULT_FLAGS without continuous physics (phase-1 parity mode)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_bits(a, b):
    return a.shape == b.shape and bool((bits(a) == bits(b)).all())


BODY_COLS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 13, 14, 15]  # xf, sweep c/a, v, w, awake, localCenter, destroyed marker


def compare_world(ref, gpu_bodies, gpu_contacts, gpu_proxies=None, gpu_events=None, what=""):
    """ref: oracle world (after the same number of steps).  Raises AssertionError with a diagnosis."""
    rb = ref.bodies()
    assert rb.shape == gpu_bodies.shape, f"{what}: body count {rb.shape} vs {gpu_bodies.shape}"
    for c in BODY_COLS:
        if not same_bits(rb[:, c], gpu_bodies[:, c]):
            bad = np.argwhere(bits(rb[:, c]) != bits(gpu_bodies[:, c]))[:4, 0]
            raise AssertionError(f"{what}: body column {c} differs at bodies {bad.tolist()}: "
                                 f"ref {rb[bad, c]} gpu {gpu_bodies[bad, c]}")
    ri, rf = ref.contacts()
    gi, gf = gpu_contacts
    assert ri.shape == gi.shape, f"{what}: contact count ref {ri.shape[0]} gpu {gi.shape[0]}"
    if not (ri == gi).all():
        bad = np.argwhere(ri != gi)[:4]
        raise AssertionError(f"{what}: contact ints differ at {bad.tolist()}: ref {ri[bad[0][0]]} gpu {gi[bad[0][0]]}")
    if not same_bits(rf, gf):
        bad = np.argwhere(bits(rf) != bits(gf))[:4]
        raise AssertionError(f"{what}: contact floats differ at {bad.tolist()}: ref {rf[bad[0][0]]} gpu {gf[bad[0][0]]}")
    if gpu_proxies is not None:
        rp, _ = ref.proxies()
        assert same_bits(rp, gpu_proxies), f"{what}: fat AABBs differ"
    if gpu_events is not None:
        re_i, re_f = ref.events()
        ge_i, ge_f = gpu_events
        assert re_i.shape == ge_i.shape, f"{what}: event count ref {re_i.shape[0]} gpu {ge_i.shape[0]}"
        assert (re_i == ge_i).all(), f"{what}: event ints differ"
        assert same_bits(re_f, ge_f), f"{what}: event impulses differ"
