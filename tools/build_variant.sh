#!/bin/bash
# Build a variant of libb2gpu.so with extra -D flags: tools/build_variant.sh <name> [-DFLAG ...]
# -> box2dpp_b200/libb2gpu_<name>.so; run with B2GPU_LIB=box2dpp_b200/libb2gpu_<name>.so
NAME=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-ffp-contract=off "$@" -shared -o box2dpp_b200/libb2gpu_${NAME}.so box2dpp_b200/csrc/b2g_step.cu box2dpp_b200/csrc/b2g_host.cpp
