#!/bin/bash
# determinism of the default build in three regimes, A/B of early vs late stage release, parity suite, smoke, bench
TAG=${1:-r2q}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.csv
( RUNS=41 B2GPU_GROUPS=2 timeout 600 python tools/diag_determinism.py 16384 120 1
  RUNS=21 B2GPU_GROUPS=2 timeout 600 python tools/diag_determinism.py 16384 120 n
  RUNS=21 B2GPU_GROUPS=4 timeout 600 python tools/diag_determinism.py 16384 120 1
  B2GPU_LIB=box2dpp_b200/libb2gpu_early.so RUNS=21 B2GPU_GROUPS=2 timeout 600 python tools/diag_determinism.py 16384 120 1 ) > gpurun_out/${TAG}_determinism.log 2>&1
cat gpurun_out/${TAG}_determinism.log
bash tools/ab.sh ${TAG} base early
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
tail -3 gpurun_out/${TAG}_smoke.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
