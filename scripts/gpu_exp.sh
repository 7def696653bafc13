#!/bin/bash
# Experiment run: a few parity tests, then stage timings of the staged K1 under tuning settings and alternative builds.
#   scripts/gpu_exp.sh NAME ["set-group" ...]      each set-group is one exp_probe.py measurement (field=value[,field=value])
set -x
mkdir -p gpurun_out
L=gpurun_out/${1:-exp}.log
shift
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "sort or staged or hostile or capacity or sharding" >> $L 2>&1
SETS=(--set "")
for g in "$@"; do SETS+=(--set "$g"); done
python scripts/exp_probe.py "${SETS[@]}" >> $L 2>&1
python scripts/exp_probe.py --glass "${SETS[@]}" >> $L 2>&1
for so in scripts/_build/exp/*.so; do [ -f $so ] && python scripts/exp_probe.py --so $so --set "" >> $L 2>&1; done
grep -v "^+" $L | tail -40
