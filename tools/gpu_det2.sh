#!/bin/bash
# determinism statistics per library variant: tools/gpu_det2.sh <tag> <variant> ...
TAG=$1; shift
mkdir -p gpurun_out
for V in "$@"; do
  LIB=box2dpp_b200/libb2gpu_${V}.so
  [ "$V" = "base" ] && LIB=box2dpp_b200/libb2gpu.so
  echo "== $V"
  B2GPU_LIB=$LIB RUNS=${RUNS:-16} B2GPU_GROUPS=${GROUPS_:-2} timeout 400 python tools/diag_determinism.py 16384 120 ${MODE:-1} 2>&1 | tail -4
done > gpurun_out/${TAG}_determinism.log 2>&1
cat gpurun_out/${TAG}_determinism.log
