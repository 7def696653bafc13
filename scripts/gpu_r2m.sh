#!/bin/bash
set -x
mkdir -p gpurun_out
L=gpurun_out/${1:-r2m}_diag.log
: > $L
for so in scripts/_build/exp/*.so; do timeout 100 python scripts/c4_diag.py --so $so 64 >> $L 2>&1; done
grep -v "^+" $L
