#!/bin/bash
set -x
mkdir -p gpurun_out
L=gpurun_out/${1:-expb}.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "bvh or lbvh or config4 or hostile or staged or edge or small_scene or capacity" >> $L 2>&1
python scripts/exp_bvh.py --set "" --set bvh_active_min=16 --set bvh_descend_min=16 >> $L 2>&1
grep -v "^+" $L | tail -30
