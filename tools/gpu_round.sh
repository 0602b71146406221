#!/bin/bash
# Standard GPU session: parity tests, bench (both arms), ncu launch list, ncu full capture of one kernel.
# Usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag> [worlds] [kernel] [skip-tests]
TAG=${1:-r1}
WORLDS=${2:-16384}
KERNEL=${3:-k_solve_velocity}
SKIPTESTS=${4:-0}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.csv
nproc > gpurun_out/${TAG}_nproc.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/${TAG}_nproc.txt
if [ "$SKIPTESTS" = "0" ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
  tail -5 gpurun_out/${TAG}_pytest.log
fi
( time python bench.py --steps 200 --warmup 5 --worlds-per-gpu $WORLDS ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
( time python bench.py --impl reference --steps 100 --warmup 5 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 600 gpurun_out/${TAG}_bench_ref.json
# The two ncu passes run the batch as ONE world group (B2GPU_GROUPS=1) so that a launch is the whole batch, like the
# per-kernel CUDA-event times of bench.py's roofline block.
# launch list: every kernel with its device time (cold cache, serialised: compare shares); the skip count passes the
# world-building uploads (~300 launches) and 290 pre-roll steps (8 launches each), so the list shows the same settled
# regime as the end of bench.py's default timed region (where its per-kernel CUDA-event times are taken)
B2GPU_GROUPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 40 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 12 --warmup 3 --preroll 290 --worlds-per-gpu $WORLDS --no-cpu-baseline --e2e-steps 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -3 gpurun_out/${TAG}_launches.csv
# full capture of one kernel (a launch on the settled pyramid, well after the pre-roll)
B2GPU_GROUPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^${KERNEL} -s 95 -c 1 -f \
    -o gpurun_out/${TAG}_${KERNEL} python bench.py --steps 6 --warmup 3 --worlds-per-gpu $WORLDS --no-cpu-baseline --e2e-steps 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
