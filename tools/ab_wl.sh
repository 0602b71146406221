#!/bin/bash
# A/B of library variants on another workload: tools/ab_wl.sh <tag> <workload> <variant> ...
TAG=$1; WL=$2; shift; shift
for V in "$@"; do
  LIB=box2dpp_b200/libb2gpu_${V}.so
  [ "$V" = "base" ] && LIB=box2dpp_b200/libb2gpu.so
  B2GPU_LIB=$LIB timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 5 --workload $WL > gpurun_out/${TAG}_${WL}_${V}.json 2> gpurun_out/${TAG}_${WL}_${V}.err
  python - "$V" gpurun_out/${TAG}_${WL}_${V}.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().split('\n')[-1])
    k=d['roofline']['kernels_ms']
    print(sys.argv[1], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], ' '.join(f"{a.replace('k_','')}={b:.3f}" for a,b in k.items()))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
