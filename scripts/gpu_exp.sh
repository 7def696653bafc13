#!/bin/bash
# One GPU call: the env-knob sweep on the in-tree library, then each experimental build of scripts/exp_build.sh swapped in.
mkdir -p gpurun_out
timeout 600 python scripts/exp_probe.py default > gpurun_out/x_default.log 2>&1
cp rayz_b200/lib/librayz_cuda.so /tmp/librayz_keep.so
for f in scripts/_build/exp/s7.so; do
  n=$(basename $f .so)
  cp $f rayz_b200/lib/librayz_cuda.so
  timeout 300 python scripts/exp_probe.py $n --quick > gpurun_out/x_$n.log 2>&1
done
cp /tmp/librayz_keep.so rayz_b200/lib/librayz_cuda.so
cat gpurun_out/x_*.log | cut -c1-250
