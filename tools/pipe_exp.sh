#!/bin/bash
# Group / pipeline / shared-memory-diet sweep (GPU).  Variants are built with tools/build_variant.sh.
run() { env "$@" python tools/pipe_exp.py 2>&1 | tail -1; }
run B2GPU_GROUPS=1
run B2GPU_GROUPS=2
run B2GPU_GROUPS=2 B2GPU_PIPELINE=1
for lib in diet4 diet5; do
  L=box2dpp_b200/libb2gpu_$lib.so
  [ -f $L ] || continue
  run B2GPU_LIB=$L B2GPU_GROUPS=1
  run B2GPU_LIB=$L B2GPU_GROUPS=2
  run B2GPU_LIB=$L B2GPU_GROUPS=2 B2GPU_PIPELINE=1
  run B2GPU_LIB=$L B2GPU_GROUPS=2 B2GPU_PIPELINE=1 B2GPU_ISLAND_KB=18
  run B2GPU_LIB=$L B2GPU_GROUPS=2 B2GPU_PIPELINE=1 B2GPU_ISLAND_KB=9
  run B2GPU_LIB=$L B2GPU_GROUPS=2 B2GPU_PIPELINE=0 B2GPU_ISLAND_KB=9
  run B2GPU_LIB=$L B2GPU_GROUPS=4 B2GPU_PIPELINE=1 B2GPU_ISLAND_KB=9
  run B2GPU_LIB=$L B2GPU_GROUPS=3 B2GPU_PIPELINE=1 B2GPU_ISLAND_KB=9
done
echo "== trace diet4 groups 2 pipeline island 9"
B2GPU_LIB=box2dpp_b200/libb2gpu_diet4.so B2GPU_GROUPS=2 B2GPU_PIPELINE=1 B2GPU_ISLAND_KB=9 python tools/groups_trace.py 2>&1 | tail -12
