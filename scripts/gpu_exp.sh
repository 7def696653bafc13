#!/bin/bash
# Experiment run: stage timings of the staged K1 under tuning settings and alternative builds (scripts/exp_probe.py).
set -x
mkdir -p gpurun_out
L=gpurun_out/${1:-exp}.log
: > $L
python scripts/exp_probe.py --set "" --set unit_entries=1024 --set unit_entries=256 --set queue_log2=28 --set queue_log2=28,unit_entries=1024 --set second_stages=4 --set second_stages=3 >> $L 2>&1
for so in scripts/_build/exp/*.so; do python scripts/exp_probe.py --so $so --set "" >> $L 2>&1; done
cat $L
