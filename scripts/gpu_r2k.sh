#!/bin/bash
# Camera stage of the BVH family with tile lists on 8 x 4 pixel blocks: BVH parity tests, then config 4's camera-stage time and
# counters for the current build and the variants under scripts/_build/exp/ (list capacity; cl0 = per-ray walks).
set -x
T=${1:-r2k}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "bvh or lbvh or config4 or big_job or capacity_overflow or sharding or multi" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -n 3 gpurun_out/${T}_pytest.log
L=gpurun_out/${T}_diag.log
timeout 200 python scripts/c4_diag.py 64 128 > $L 2>&1
for so in scripts/_build/exp/*.so; do timeout 100 python scripts/c4_diag.py --so $so 64 >> $L 2>&1; done
timeout 100 python scripts/exp_bvh.py --set "" >> $L 2>&1
grep -v "^+" $L
