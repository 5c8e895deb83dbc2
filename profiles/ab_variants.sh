#!/bin/bash
# Developer tool (GPU box): per-kernel medians of compile-time variants built by profiles/build_variant.sh
#   profiles/ab_variants.sh [workload] NAME...
cd "$(dirname "$0")/.."
L=edgeml-object-detection_b200/liborie_b200.so
wl=$1; shift
cp $L /tmp/_default.so
echo "default: $(python profiles/ab.py $wl 2>&1 | tail -1)"
for v in "$@"; do
    cp profiles/_variants/$v.so $L
    echo "$v: $(python profiles/ab.py $wl 2>&1 | tail -1)"
done
cp /tmp/_default.so $L
