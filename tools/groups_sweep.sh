#!/bin/bash
# World-group sweep: the same bench with 1, 2, 3, 4 groups of tiles stepping on separate streams (B2GPU_GROUPS).
TAG=${1:-r1}
EXTRA=${2:-}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_groups.jsonl
: > $OUT
for G in 1 2 3 4 6; do
  B2GPU_GROUPS=$G timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 10 --preroll 200 $EXTRA >> $OUT 2>> gpurun_out/${TAG}_groups.err
  tail -1 $OUT | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('groups $G', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])"
done
tail -3 gpurun_out/${TAG}_groups.err
