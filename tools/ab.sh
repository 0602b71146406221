#!/bin/bash
# A/B of library variants on the GPU box: tools/ab.sh <tag> <variant> [<variant> ...]
# (variants built beforehand with tools/build_variant.sh; "base" = the default libb2gpu.so)
TAG=$1; shift
mkdir -p gpurun_out
for V in "$@"; do
  LIB=box2dpp_b200/libb2gpu_${V}.so
  [ "$V" = "base" ] && LIB=box2dpp_b200/libb2gpu.so
  B2GPU_LIB=$LIB timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${TAG}_${V}.json 2> gpurun_out/${TAG}_${V}.err
  python - "$V" gpurun_out/${TAG}_${V}.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().split('\n')[-1])
    k=d['roofline']['kernels_ms']
    print(sys.argv[1], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], ' '.join(f"{a.replace('k_','')}={b:.3f}" for a,b in k.items()))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
