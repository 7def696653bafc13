#!/bin/bash
# Camera stage of the BVH family with per-tile sphere lists (RZ_BVH_CAMERA_LISTS): BVH parity tests on the current build, then
# configs 2 / 4 under chunk sizes and against the variants under scripts/_build/exp/ (cl0 = per-ray walks, the previous form).
set -x
T=${1:-r2j}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "bvh or lbvh or config4 or big_job or capacity_overflow" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -n 3 gpurun_out/${T}_pytest.log
L=gpurun_out/${T}_bvh.log
timeout 120 python scripts/exp_bvh.py --only 4 --set "" --set chunk=32 --set chunk=64 --set chunk=128 > $L 2>&1
for so in scripts/_build/exp/*.so; do timeout 60 python scripts/exp_bvh.py --only 4 --so $so --set chunk=64 >> $L 2>&1; done
timeout 60 python scripts/exp_bvh.py --only 2 --set "" --set chunk=64 >> $L 2>&1
grep -v "^+" $L
