#!/bin/bash
# determinism statistics per configuration: tools/gpu_det.sh <tag>
TAG=${1:-det}
mkdir -p gpurun_out
( RUNS=${RUNS:-16} B2GPU_GROUPS=2 timeout 400 python tools/diag_determinism.py 16384 120 1
  RUNS=${RUNS:-16} B2GPU_GROUPS=1 timeout 400 python tools/diag_determinism.py 16384 120 1
  RUNS=${RUNS:-16} B2GPU_GROUPS=2 timeout 400 python tools/diag_determinism.py 16384 120 1 generic
  RUNS=${RUNS:-16} B2GPU_GROUPS=2 timeout 400 python tools/diag_determinism.py 16384 120 n ) > gpurun_out/${TAG}_determinism.log 2>&1
cat gpurun_out/${TAG}_determinism.log
