#!/bin/bash
# BASELINE configs[4]: world-count sweep on one GPU (pyramid-20), plus the mixed workload and the reference's
# default flags (TOI + sleeping on).  Usage (GPU box, repo root): bash tools/sweep.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_sweep.jsonl
: > $OUT
for W in 1024 4096 8192 16384 65536 262144; do
  S=100; [ $W -ge 65536 ] && S=30; [ $W -ge 262144 ] && S=10
  timeout 900 python bench.py --steps $S --warmup 3 --worlds-per-gpu $W --no-cpu-baseline --e2e-steps 5 >> $OUT 2>> gpurun_out/${TAG}_sweep.err
done
timeout 900 python bench.py --steps 100 --warmup 3 --toi --no-cpu-baseline --e2e-steps 5 >> $OUT 2>> gpurun_out/${TAG}_sweep.err
timeout 900 python bench.py --steps 100 --warmup 3 --toi --sleep --no-cpu-baseline --e2e-steps 5 >> $OUT 2>> gpurun_out/${TAG}_sweep.err
timeout 1200 python bench.py --steps 100 --warmup 3 --workload mixed >> $OUT 2>> gpurun_out/${TAG}_sweep.err
python - <<'PY' $OUT
import json, sys
for line in open(sys.argv[1]):
    try:
        d = json.loads(line)
    except Exception:
        continue
    c = d["config"]
    print(f'{c["worlds_per_gpu"]:>7} worlds toi={int(c["continuous_physics"])} sleep={int(c["sleeping"])} '
          f'{c["workload"][:40]:40s} {d["value"]/1e6:8.3f} M world-steps/s  e2e {d["e2e"]["value"]/1e6:8.3f}  '
          f'{d["ms_per_step"]:8.3f} ms/step  step_frac {d["roofline"]["step_frac"]:.4f} top {d["roofline"]["kernel"]}')
PY
tail -3 gpurun_out/${TAG}_sweep.err
