#!/bin/bash
# Developer A/B builds of the library (tools/librt_<name>.so, git-ignored, shipped to the GPU box by gpurun): the traversal-stack and
# top-of-tree experiments of csrc/kernels.cu (bvh_closest_hit_ww) and other register budgets of the persistent kernel.
# Run tools/gpu_libs.py librt_<name>.so ... on a B200 to time them; tools/profile_lib.py <lib> is the ncu target.
set -e
cd "$(dirname "$0")/../raytracer_rs_b200/csrc"
build() {  # name, flags
  make -j8 OUT=../../tools/librt_$1.so BUILD=build_$1 EXTRA="$2" 2>&1 | grep -E "error|trace_shade_persistent_kernelILi1ELi0ELb0" -A3 | grep -E "error|Used|spill" | sed "s/^/$1: /"
}
build stk8 "-DRT_STACK_SMEM=8"
build stk16 "-DRT_STACK_SMEM=16"
build regtop "-DRT_STACK_REGTOP=1"
build top128 "-DRT_TOP_SMEM=128"
build top512 "-DRT_TOP_SMEM=512"
build blocks4 "-DRT_PERSISTENT_MIN_BLOCKS=4"
build stk8blocks4 "-DRT_STACK_SMEM=8 -DRT_PERSISTENT_MIN_BLOCKS=4"
# ray-stream kernel: how much of the traversal stack lives in shared memory (default build: 8 entries)
build streamstk0 "-DRT_STREAM_STACK_SMEM=0"
build streamstk12 "-DRT_STREAM_STACK_SMEM=12"
build streamstk16 "-DRT_STREAM_STACK_SMEM=16"
