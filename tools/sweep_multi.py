#!/usr/bin/env python
"""BASELINE configs[3] and configs[4] on one 8-GPU box: bench.py cells {total worlds} x {GPUs}, several cells side by
side on disjoint GPU sets (CUDA_VISIBLE_DEVICES), one JSON line per cell appended to the output file.

    python tools/sweep_multi.py <out.jsonl> [--gpus 8] [--quick]

Cells whose per-GPU share does not fit (0.42 MB of device state per pyramid-20 world, 180 GB per GPU) are skipped and
said so.  The headline cells (16,384 worlds per GPU, with e2e) are run ALONE first so that nothing shares the host."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MB_PER_PYRAMID_WORLD = 0.42


def launch(job, gpus, port, out_dir):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=",".join(str(g) for g in gpus))
    n = len(gpus)
    args = ["--gpus", str(n), "--steps", str(job["steps"]), "--warmup", "3", "--worlds-per-gpu", str(job["wpg"]),
            "--workload", job["workload"], "--no-cpu-baseline", "--e2e-steps", str(job["e2e"])]
    if n == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py")] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(ROOT, "bench.py")] + args
    log = open(os.path.join(out_dir, f"sweep_{job['name']}.log"), "w")
    return subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=log, text=True, cwd=ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cells", default="", help="explicit plan: workload:total_worlds:gpus[:alone],... (replaces the default one)")
    a = ap.parse_args()
    out_dir = os.path.dirname(os.path.abspath(a.out))
    G = a.gpus
    counts = [n for n in (1, 2, 4, 8) if n <= G]
    alone, shared = [], []
    # headline cells, alone: BASELINE configs[2] weak-scaled, and configs[3] = 131,072 mixed worlds on 8 GPUs
    for n in sorted({1, counts[-1]}):
        alone.append({"name": f"pyramid_16384x{n}", "workload": "pyramid", "wpg": 16384, "n": n, "steps": 20, "e2e": 20})
    for n in sorted({1, counts[-1]}):
        alone.append({"name": f"mixed_16384x{n}", "workload": "mixed", "wpg": 16384, "n": n, "steps": 20, "e2e": 20})
    # configs[4]: total worlds x GPUs
    totals = [1 << 10, 1 << 14, 1 << 16, 1 << 18, 1 << 20]
    skipped = []
    for t in totals:
        for n in counts:
            wpg = t // n
            if wpg == 16384 and n in (1, counts[-1]):
                continue  # measured above, alone
            if t == 1 << 10 and n not in (1, counts[-1]):
                continue  # the small end: one GPU and all of them
            if wpg * MB_PER_PYRAMID_WORLD > 150e3:
                skipped.append({"total_worlds": t, "gpus": n, "why": f"{wpg} worlds per GPU need ~{wpg * MB_PER_PYRAMID_WORLD / 1e3:.0f} GB"})
                continue
            if a.quick and not (t in (1 << 10, 1 << 16, 1 << 20)):
                continue
            shared.append({"name": f"pyramid_{t}on{n}", "workload": "pyramid", "wpg": wpg, "n": n,
                           "steps": 20 if wpg <= 65536 else 10, "e2e": 3})
    if a.cells:
        alone, shared, skipped = [], [], []
        for c in a.cells.split(","):
            f = c.split(":")
            wl, t, n = f[0], int(f[1]), int(f[2])
            job = {"name": f"{wl}_{t}on{n}", "workload": wl, "wpg": t // n, "n": n, "steps": 20 if t // n <= 65536 else 10,
                   "e2e": 20 if len(f) > 3 else 3}
            (alone if len(f) > 3 else shared).append(job)
    results = open(a.out, "a")
    for s in skipped:
        results.write(json.dumps({"skipped": s}) + "\n")
    results.flush()

    def collect(job, proc):
        stdout, _ = proc.communicate()
        lines = [l for l in stdout.splitlines() if l.startswith("{")]
        rec = {"cell": job["name"], "rc": proc.returncode}
        if lines:
            rec.update(json.loads(lines[-1]))
        results.write(json.dumps(rec) + "\n")
        results.flush()
        v = rec.get("value")
        print(f"{job['name']:24s} rc={proc.returncode} " + (f"{v / 1e6:8.3f} M world-steps/s  e2e {rec['e2e']['value'] / 1e6:8.3f}  "
              f"{rec['ms_per_step']:8.3f} ms/step" if v else "no result"), flush=True)

    port = 29610
    for job in alone:
        t0 = time.time()
        p = launch(job, list(range(job["n"])), port, out_dir)
        port += 1
        collect(job, p)
        print(f"   ({time.time() - t0:.0f} s)", flush=True)
    # side by side: greedy placement, biggest jobs first
    shared.sort(key=lambda j: (-j["n"], -j["wpg"]))
    free = list(range(G))
    running = []
    while shared or running:
        for job in list(shared):
            if job["n"] <= len(free):
                gpus, free = free[:job["n"]], free[job["n"]:]
                running.append((job, gpus, launch(job, gpus, port, out_dir)))
                port += 1
                shared.remove(job)
        time.sleep(1.0)
        for item in list(running):
            job, gpus, proc = item
            if proc.poll() is not None:
                collect(job, proc)
                free += gpus
                running.remove(item)
    results.close()


if __name__ == "__main__":
    main()
