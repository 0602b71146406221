#!/bin/bash
# Round-2 GPU session: determinism stress (two runs of one batch must agree bit for bit), parity suite, smoke, bench.
TAG=${1:-r2j}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.csv
( for G in 2 1 4; do
    RUNS=5 B2GPU_GROUPS=$G timeout 300 python tools/diag_determinism.py 16384 120 n
  done
  RUNS=3 timeout 300 python tools/diag_determinism.py 16384 120 1 ) > gpurun_out/${TAG}_determinism.log 2>&1
cat gpurun_out/${TAG}_determinism.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
tail -3 gpurun_out/${TAG}_smoke.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
