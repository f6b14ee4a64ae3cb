#!/bin/bash
# Builds eraytracer_b200/lib/libert_b200_<name>.so with extra -D flags (tuning experiments; select with ERT_B200_LIB).
# usage: bash tools/build_variant.sh <name> [-DFOO=1 ...]
name=$1; shift
cd "$(dirname "$0")/../eraytracer_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-fvisibility=hidden,-O2,-Wall "$@" -shared -o ../lib/libert_b200_$name.so \
  ert_api.cu bvh_build.cpp light_grid.cpp cell_grid.cpp -lpthread -Xptxas -v 2>&1 | grep -A3 "wf_shadow_shade\|wf_trace_path_refillILb0ELb1ELb1\|wf_trace_pathILb0ELb0ELb1ELb1" | grep -E "Compiling|registers|spill"
