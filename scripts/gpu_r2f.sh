#!/bin/bash
# BVH kernel variants under scripts/_build/exp/ against the default build: config 2 (variant=bvh) and config 4
set -x
mkdir -p gpurun_out
L=gpurun_out/${1:-r2f}_bvh.log
: > $L
timeout 60 python scripts/exp_bvh.py --set "" >> $L 2>&1
for so in scripts/_build/exp/*.so; do timeout 60 python scripts/exp_bvh.py --so $so --set "" >> $L 2>&1; done
grep -v "^+" $L
