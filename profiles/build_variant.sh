#!/bin/bash
# Developer tool: build a compile-time variant of the library next to the real one, for A/B runs on the GPU box
#   profiles/build_variant.sh NAME [-DORIE_WALK_RESIDENT=1280 ...]   ->  profiles/_variants/NAME.so
# (profiles/ab_variants.sh copies each variant over liborie_b200.so in turn on the box)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p profiles/_variants
C=edgeml-object-detection_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 -shared -cudart static "$@" \
    $C/api.cu $C/sort.cu $C/match.cu $C/index.cu $C/reward.cu $C/rank.cu $C/dcsb_fit.cu -o profiles/_variants/$name.so
echo built profiles/_variants/$name.so
