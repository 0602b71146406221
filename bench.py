#!/usr/bin/env python
"""bench.py — batched b2World::Step throughput on B200 (BASELINE.json metric: world-steps/s).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU Step, all host cores

Workload (config.workload): BASELINE.json configs[2] — 16,384 independent 20-row box-pyramid worlds per B200
(weak scaling: every rank owns 16,384 worlds, sharded by world index, no collective on the step path),
jittered per world so control flow diverges inside a warp, dt = 1/60, 8 velocity / 3 position iterations,
sleeping off and continuous physics off on BOTH arms, contact events off on both arms.
Worlds are pre-rolled 90 steps (untimed) so the timed steps run on the settled pyramid (~590 contacts,
~400 touching per world), which is the expensive regime.

One JSON line on stdout (rank 0).  `value` = device-timed world-steps/s with state resident in HBM;
`e2e` = same metric through the C ABI with host buffers (forces H2D + transforms D2H every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 1.0 / 60.0
ROWS = 20
BODIES_DYNAMIC = ROWS * (ROWS + 1) // 2
PREROLL = 90
# algorithmic bytes per item per step
BYTES_PER_DYNAMIC_BODY = 144
BYTES_PER_PROXY = 84
BYTES_PER_CONTACT = 188


def flags_throughput():
    from box2dpp_b200 import scene as S
    return S.WORLD_DEFAULT & ~S.WORLD_CONTINUOUS & ~S.WORLD_ALLOW_SLEEP


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.window = index, [], False, None

    def run(self):
        while not self.stop_flag:
            t = time.perf_counter()
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 8:
                    self.samples.append((t, f))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [s for s in self.samples if self.window and self.window[0] <= s[0] <= self.window[1]]
        if not inside and self.window:  # region shorter than one sample: take the closest one
            mid = 0.5 * (self.window[0] + self.window[1])
            inside = [min(self.samples, key=lambda s: abs(s[0] - mid))]
        use = inside or self.samples
        sm = sorted(float(s[1][0]) for s in use)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, f in use:
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(use[0][1][1]), "reasons": sorted(reasons),
                "samples": len(use), "power_w_max": max(float(s[1][2]) for s in use)}


def build_scenes(workload, count, first_seed, flags, jittered=True):
    """`count` worlds of the named workload, world i seeded with first_seed + i.  Returns (batch-or-list, scene(i))."""
    from box2dpp_b200 import scene as S
    if workload == "pyramid":
        sb = S.pyramid_batch(count, ROWS, flags=flags, first_seed=first_seed, jittered=jittered)
        return sb, sb.scene
    make = S.mixed if workload == "mixed" else S.jointed
    scenes = [make(first_seed + i, flags=flags) for i in range(count)]
    return scenes, scenes.__getitem__


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# measured on this pool's GPU hosts (16 cores): ~47 k pyramid / ~310 k mixed / ~600 k jointed world-steps/s; only used
# to size the CPU sample so that one timed block lasts about a second
CPU_RATE_GUESS = {"pyramid": 3000.0, "mixed": 20000.0, "jointed": 40000.0}  # world-steps/s per thread


def cpu_reference_run(steps, warmup, threads=None, worlds_per_thread=None, workload="pyramid", flags=None, repeats=3):
    """The reference's own b2World::Step (vanilla Box2D compiled from /root/reference into oracle/_ref),
    one world per thread slot on every host core.  The sample is sized so that a timed block of `steps` steps lasts
    >= ~1 s (a 0.1 s block read 76 k - 92 k on the same box in round 1), the block is timed `repeats` times back to
    back and the median is the value; the spread is reported beside it.  If oracle/_ref is absent the C port is NOT a
    fair stand-in for speed, so we say so."""
    from box2dpp_b200 import scene as S
    from oracle import worlds  # TEST INFRASTRUCTURE: used only as the CPU baseline / reference arm
    threads = threads or os.cpu_count()
    if worlds_per_thread is None:
        worlds_per_thread = int(min(512, max(16, -(-1.2 * CPU_RATE_GUESS[workload] // max(steps, 1)))))
    n = threads * worlds_per_thread
    flags = flags_throughput() if flags is None else flags
    _, get = build_scenes(workload, n, 0, flags)
    scenes = [get(i) for i in range(n)]
    if not worlds.available("ref"):
        return None
    secs = worlds.bench_run_repeats(scenes, worlds_per_thread, threads, steps, warmup, DT, 8, 3, repeats)
    med = sorted(secs)[len(secs) // 2]
    what = {"pyramid": f"jittered pyramid-{ROWS}", "mixed": "mixed-shape (chain terrain, circles, polygons)",
            "jointed": "jointed (bridge, cart-pole, walker, pendulums)"}[workload]
    return {"value": n * steps / med, "unit": "world-steps/s", "cores": threads, "kind": "reference",
            "cpu": cpu_model(),
            "sample": f"{threads} threads x {worlds_per_thread} {what} worlds, {warmup} warm-up + "
                      f"{repeats} x {steps} timed steps (median), 8/3 iterations, sleeping {'on' if flags & S.WORLD_ALLOW_SLEEP else 'off'}, "
                      f"TOI {'on' if flags & S.WORLD_CONTINUOUS else 'off'}, vanilla Box2D 2.3.2 -O2",
            "repeat_values": [n * steps / x for x in secs],
            "spread": (max(secs) - min(secs)) / med,
            "seconds": med, "worlds": n}


def pin_to_gpu_numa(gpu_index):
    """Bind this process to the CPUs of the NUMA node its GPU hangs off (before any pinned buffer is allocated, so that
    first touch puts the staging memory next to the GPU's root port).  Returns what it did, for the JSON line."""
    info = {"gpu": gpu_index, "node": None, "cpus": None}
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={gpu_index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("0000") and len(bus.split(":")[0]) == 8:
            bus = bus[4:]  # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info["node"] = node
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["cpus"] = f"{min(use)}-{max(use)} ({len(use)} of {len(allowed)} allowed)"
        else:
            info["cpus"] = f"unchanged: {len(allowed)} allowed CPUs, {'all local' if use == allowed else 'none local'}"
    except Exception as exc:  # a container without sysfs PCI nodes: leave the affinity alone
        info["error"] = str(exc)[:80]
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds-per-gpu", type=int, default=16384)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--identical", action="store_true", help="identical worlds instead of jittered ones")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--generic-solver", action="store_true", help="force the fallback solver kernel")
    ap.add_argument("--workload", default="pyramid", choices=["pyramid", "mixed", "jointed"],
                    help="pyramid = BASELINE configs[2] (the headline); mixed = configs[3] per-GPU share; "
                         "jointed = bridge + cart-pole + walker + pendulums (joints on the tile path)")
    ap.add_argument("--toi", action="store_true", help="continuous physics on (the reference's default) on both arms")
    ap.add_argument("--sleep", action="store_true", help="sleeping allowed (the reference's default) on both arms")
    ap.add_argument("--preroll", type=int, default=-1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3)

    from box2dpp_b200 import scene as S
    flags = flags_throughput()
    if args.toi:
        flags |= S.WORLD_CONTINUOUS
    if args.sleep:
        flags |= S.WORLD_ALLOW_SLEEP
    preroll = args.preroll if args.preroll >= 0 else (PREROLL if args.workload == "pyramid" else 120)
    if args.workload == "pyramid":
        nbody, nproxy = BODIES_DYNAMIC, BODIES_DYNAMIC + 1
        what = (f"{args.worlds_per_gpu} {'identical' if args.identical else 'jittered'} pyramid-{ROWS} worlds "
                f"per GPU (BASELINE configs[2])")
    elif args.workload == "mixed":
        nbody, nproxy = 40, 72
        what = (f"{args.worlds_per_gpu} mixed-shape worlds per GPU (BASELINE configs[3]'s per-GPU share: 32-edge chain "
                f"terrain, 20 circles + 20 random convex polygons, seed = world index)")
    else:
        nbody, nproxy = 26, 27
        what = (f"{args.worlds_per_gpu} jointed worlds per GPU (plank bridge, cart-pole, two-legged walker, pendulums: "
                f"22 revolute / prismatic / distance joints, seed = world index)")
    config = {"workload": what + ", dt=1/60, 8 vel / 3 pos iterations",
              "worlds_per_gpu": args.worlds_per_gpu, "dynamic_bodies_per_world": nbody,
              "sleeping": bool(args.sleep), "continuous_physics": bool(args.toi), "warm_starting": True, "events": "off",
              "preroll_steps": preroll, "sharding": f"world-index blocks x{args.gpus}, no collective",
              "l2": "working set (~0.25 MB/world x worlds) far exceeds the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.steps, preroll + warmup, workload=args.workload, flags=flags)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libb2ref.so not built on this box"}))
            return
        line = {"impl": "reference", "metric": "world_steps_per_s", "value": r["value"], "unit": "world-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "body_steps_per_s": r["value"] * nbody,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu", "repeat_values", "spread")},
                "e2e": {"value": r["value"], "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    from box2dpp_b200.batch import F_CONTACT_COUNT, F_ISLAND_CONTACT_END, F_ISLAND_COUNT, WorldBatch

    # nvidia-smi counts physical GPUs; under CUDA_VISIBLE_DEVICES (tools/sweep_multi.py runs several jobs side by side on
    # disjoint GPU sets) CUDA device `local_rank` is another one
    visible = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
    smi_index = int(visible[local_rank]) if local_rank < len(visible) else local_rank
    numa = pin_to_gpu_numa(smi_index)
    dist = None
    if world_size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from box2dpp_b200.shard import shard_range
    # weak scaling: the job has worlds_per_gpu x N worlds; rank r owns a contiguous world-index block
    first_world, W = shard_range(args.worlds_per_gpu * world_size, world_size, rank)
    sb, _ = build_scenes(args.workload, W, first_world, flags, jittered=not args.identical)
    batch = WorldBatch(sb, device=local_rank, block_size=args.block, generic_solver=args.generic_solver)
    batch.step_n(DT, 8, 3, preroll)
    batch.sync()

    sampler = ClockSampler(smi_index)
    sampler.start()
    batch.step_n(DT, 8, 3, warmup)
    batch.sync()

    # ---- timed region: K steps, CUDA events on the launching stream, barrier + sync on both sides ----
    barrier()
    launches0 = batch.launch_count()
    t0 = time.perf_counter()
    batch.step_n(DT, 8, 3, args.steps)
    batch.sync()
    t1 = time.perf_counter()
    ms = batch.last_step_ms()
    launches = batch.launch_count() - launches0
    barrier()
    sampler.window = (t0, t1)
    ms = max_over_ranks(ms)
    total_worlds = W * world_size
    value = total_worlds * args.steps / (ms * 1e-3)

    # ---- per-kernel times (live CUDA events around every kernel, same stream), taken right after the timed
    # region so that they describe the same regime (the position solver's early exit makes steps cheaper as
    # the pyramid keeps settling) ----
    prof_steps = 10
    kms = {k: v / prof_steps for k, v in batch.profile_kernels(DT, 8, 3, prof_steps).items()}

    # ---- end to end through the C ABI with host buffers: forces H2D, step, transforms D2H, every step ----
    e2e_steps = args.e2e_steps or max(10, min(args.steps, 100))
    # What crosses the bus every step: a force (fx, fy, torque) for every non-static body going in, the pose (x, y, angle)
    # of every non-static body coming out (b2gpu_batch_io_select): not the static ground / terrain, three floats per pose.
    scene0 = sb.scene(0) if hasattr(sb, "scene") else sb[0]
    io_bodies = [i for i, bd in enumerate(scene0.bodies) if bd.type != S.STATIC_BODY]
    batch.io_select(force_bodies=io_bodies, pose_bodies=io_bodies, xya=True)
    try:
        import torch
        forces_t = torch.zeros((W, len(io_bodies), 3), dtype=torch.float32).pin_memory()
        out_t = torch.empty((W, len(io_bodies), 3), dtype=torch.float32).pin_memory()
        forces, out = forces_t.numpy(), out_t.numpy()
    except Exception:
        forces = np.zeros((W, len(io_bodies), 3), np.float32)
        out = np.empty((W, len(io_bodies), 3), np.float32)
    # (a) the three separate calls, everything on the batch's stream: copy in, step, copy out
    for _ in range(3):
        batch.apply_forces(forces); batch.step(DT, 8, 3); batch.read_transforms(out)
    barrier()
    t2 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.apply_forces(forces)
        batch.step(DT, 8, 3)
        batch.read_transforms(out)  # synchronises: the caller holds this step's transforms
    t3 = time.perf_counter()
    sep_s = max_over_ranks(t3 - t2)
    # (b) the fused call (b2gpu_batch_step_io): same work and bytes, each world group's copies under the other groups'
    # kernels; synchronises every step as well.  This is the call a per-step caller makes, so it is the e2e figure.
    for _ in range(3):
        batch.step_io(forces, DT, 8, 3, out)
    barrier()
    t2 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.step_io(forces, DT, 8, 3, out)
    t3 = time.perf_counter()
    e2e_s = max_over_ranks(t3 - t2)
    # device-resident steps right after, i.e. in the regime the e2e loop ran in (the pyramid keeps settling, so steps
    # get cheaper than in the timed region above): the difference to e2e's ms_per_step is what the host I/O costs
    batch.sync()
    t2 = time.perf_counter()
    batch.step_n(DT, 8, 3, 40)
    batch.sync()
    same_regime_ms = max_over_ranks(time.perf_counter() - t2) * 1e3 / 40
    e2e = {"value": total_worlds * e2e_steps / e2e_s, "unit": "world-steps/s",
           "h2d_bytes_per_step": int(forces.nbytes) * world_size, "d2h_bytes_per_step": int(out.nbytes) * world_size,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "call": "b2gpu_batch_step_io",
           "io": f"forces in and poses (x, y, angle) out for the {len(io_bodies)} non-static bodies of every world "
                 f"(b2gpu_batch_io_select), pinned host buffers", "numa": numa,
           "device_resident_ms_per_step_same_regime": same_regime_ms,
           "separate_calls": {"value": total_worlds * e2e_steps / sep_s, "ms_per_step": 1e3 * sep_s / e2e_steps,
                              "call": "b2gpu_batch_apply_forces + b2gpu_batch_step + b2gpu_batch_read_transforms"}}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    contacts = batch.field(F_CONTACT_COUNT)[:, 0, 0].astype(np.float64)
    nisl = batch.field(F_ISLAND_COUNT)[:, 0, 0]
    iend = batch.field(F_ISLAND_CONTACT_END)[:, :, 0]
    touching = np.array([iend[w, nisl[w] - 1] if nisl[w] > 0 else 0 for w in range(W)], dtype=np.float64)
    alg_step = nbody * BYTES_PER_DYNAMIC_BODY + nproxy * BYTES_PER_PROXY + contacts.mean() * BYTES_PER_CONTACT
    # Algorithmic bytes per world for each kernel taken alone: persistent fields it must read once
    # plus persistent fields it must write once; solver temporaries (constraint records) count zero.
    tc, nc = touching.mean(), contacts.mean()
    alg_kernel = {
        "k_collide": nc * (128 + 24 + 4) + nproxy * 84 + nbody * 16,           # manifolds r/w, ids/flags, fat AABB + shape, xf
        "k_islands": nc * (4 + 8 + 32) + nbody * 8,                              # flags, body pair, list links; body flags r/w
        "k_solve_prepare": nbody * (24 + 12 + 28 + 24 + 12) + tc * (64 + 24),    # c,a,v,w, c0/a0, mass data, force; manifold + ids
        "k_solve_velocity": nbody * 24 + tc * 32,                                # b2Velocity in/out; impulses in/out
        "k_solve_integrate": nbody * (12 + 12 + 12) + tc * 32,                   # v in, v out, c/a out; impulses -> manifold
        "k_solve_position": nbody * 24 + tc * 40,                                # b2Position in/out; local manifold geometry
        "k_solve_finalize": nbody * (12 + 16 + 16),                              # c/a, xf, flags + sleep time
        "k_sync_pairs": nproxy * (16 + 68) + nbody * (16 + 12),                  # fat AABB + shape; xf, c0/a0
        "k_solve": nbody * BYTES_PER_DYNAMIC_BODY + tc * BYTES_PER_CONTACT,
        "k_toi": nc * (4 + 4 + 4) + nbody * (16 + 16 + 4) + nproxy * 84,          # flags r/w + cached toi; sweep, sweep0.alpha0, flags; shapes
        "k_toi_events": nc * (4 + 4) + nbody * 4,                                # flags + cached toi of the list walk; body flags
    }
    top = max(kms, key=kms.get)
    top_ms = kms[top]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_kernel[top] * W / (top_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture
    # (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); only quoted for the configuration it was taken on
    traffic = None
    kernels_traffic = {}
    try:
        captured = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for name, t in captured.items():
            if t.get("worlds") == W and t.get("workload") == args.workload and not args.toi and not args.sleep:
                kernels_traffic[name] = t["dram_bytes_per_launch"]
        traffic = kernels_traffic.get(top)
    except Exception:
        pass
    step_ms_sum = sum(kms.values())
    roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                # the dominant kernel's real DRAM traffic against the same peak (its constraint records are re-streamed
                # every Gauss-Seidel pass, which the algorithmic figure counts as zero)
                "traffic_gbs": (traffic / (top_ms * 1e-3) / 1e9) if traffic else None,
                "traffic_frac": (traffic / (top_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "kernels_traffic": kernels_traffic or None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": alg_kernel[top] * W, "kernel_ms": top_ms,
                "kernel_share_of_step": top_ms / step_ms_sum,
                "kernels_ms": kms,
                "kernels_frac": {k: alg_kernel[k] * W / (v * 1e-3) / 1e9 / peak for k, v in kms.items() if v > 0},
                "step_algorithmic_bytes_per_world": alg_step,
                "step_achieved_gbs": alg_step * value / world_size / 1e9,
                "step_frac": alg_step * value / world_size / 1e9 / peak,
                "mean_contacts": float(contacts.mean()), "mean_touching_contacts": float(touching.mean())}

    status_bad = int((batch.status() != 0).sum())
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_run(100, preroll + 10, workload=args.workload, flags=flags, worlds_per_thread=32, repeats=3)
            if cpu:
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu", "spread")}
        except Exception as exc:  # the baseline is a report, never a dependency of the GPU number
            cpu = {"error": str(exc)}
    batch.close()
    if rank == 0:
        line = {"metric": "world_steps_per_s", "value": value, "unit": "world-steps/s", "n_gpus": world_size,
                "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "body_steps_per_s": value * nbody, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
                "worlds_with_overflow": status_bad}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
