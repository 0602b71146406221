#!/bin/bash
# Standard GPU session: parity tests, bench, ncu launch list, ncu full capture of the dominant kernel.
# Usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag> [worlds]
TAG=${1:-r1}
WORLDS=${2:-16384}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.csv
( time python bench.py --steps 200 --warmup 5 --worlds-per-gpu $WORLDS ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
# launch list: every kernel with its device time (cold cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 40 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 12 --warmup 3 --worlds-per-gpu $WORLDS --no-cpu-baseline --e2e-steps 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -3 gpurun_out/${TAG}_launches.csv
# full capture of the dominant kernel
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 95 -c 2 -f \
    -o gpurun_out/${TAG}_solve python bench.py --steps 6 --warmup 3 --worlds-per-gpu $WORLDS --no-cpu-baseline --e2e-steps 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
