#!/bin/bash
# Round 2, last session, second call: where the e2e call's time goes; the BVH kernel variants of scripts/_build/exp/ on configs 2 / 4
# (exp_bvh.py); BVH parity tests with the short-listed variants in place of the product library (on the box's scratch copy only).
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2c.sh r2c "m8 m8p8 m8l8"'
set -x
mkdir -p gpurun_out
T=${1:-r2c}
L=gpurun_out/${T}_bvh.log
: > $L
timeout 120 python scripts/e2e_probe.py > gpurun_out/${T}_e2e.log 2>&1
timeout 60 python scripts/exp_bvh.py --set "" >> $L 2>&1
for so in scripts/_build/exp/*.so; do [ -f $so ] && timeout 90 python scripts/exp_bvh.py --so $so --set "" >> $L 2>&1; done
grep -v "^+" $L | tail -40
cp rayz_b200/lib/librayz_cuda.so /tmp/librayz_cuda.product.so
for c in $2; do
  cp scripts/_build/exp/$c.so rayz_b200/lib/librayz_cuda.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "bvh or lbvh or config4 or staged_megakernel or capacity_overflow" > gpurun_out/${T}_pytest_$c.log 2>&1
  echo "pytest rc $?" >> gpurun_out/${T}_pytest_$c.log
  tail -2 gpurun_out/${T}_pytest_$c.log
done
cp /tmp/librayz_cuda.product.so rayz_b200/lib/librayz_cuda.so
cat gpurun_out/${T}_e2e.log | tail -3
