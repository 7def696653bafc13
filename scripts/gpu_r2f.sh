#!/bin/bash
# config 4 with the default build and the unit-size variants, three times each (run-to-run noise)
set -x
mkdir -p gpurun_out
L=gpurun_out/r2f_bvh.log
: > $L
for i in 1 2 3; do
  timeout 60 python scripts/exp_bvh.py --only 4 --set "" >> $L 2>&1
  for so in scripts/_build/exp/u512.so scripts/_build/exp/u64.so; do timeout 60 python scripts/exp_bvh.py --only 4 --so $so --set "" >> $L 2>&1; done
done
grep -v "^+" $L
