#!/bin/bash
# Round-2 measurement session: tools/gpu_round.sh (parity suite, bench both arms, ncu launch list, ncu --set full of the
# velocity kernel), then full captures of the position and collide kernels and the jointed / mixed workloads with their
# reference arms.
TAG=${1:-r2x}
bash tools/gpu_round.sh $TAG 16384 k_solve_velocity
for K in k_solve_position k_collide; do
  B2GPU_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${K} -s 95 -c 1 -f \
      -o gpurun_out/${TAG}_${K} python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
done
for WL in jointed mixed; do
  ( time timeout 300 python bench.py --steps 20 --warmup 5 --workload $WL ) > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
  tail -c 400 gpurun_out/${TAG}_bench_${WL}.json
  ( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 --workload $WL ) > gpurun_out/${TAG}_bench_${WL}_ref.json 2> gpurun_out/${TAG}_bench_${WL}_ref.err
  tail -c 300 gpurun_out/${TAG}_bench_${WL}_ref.json
done
ls -la gpurun_out | tail -20
