#!/bin/bash
# Round 2, last session: (1) GPU tests + the bench line of the current build, (2) BVH kernel A/B on config 4 (RzTuning sweeps and
# the compile-time variants scripts/exp_build.sh left under scripts/_build/exp/), (3) ncu captures of the BVH kernels ON the
# 99,856-sphere scene (camera stage + persistent queue kernel).
#   gpurun --timeout 1100 -- 'bash scripts/gpu_r2b.sh'
set -x
mkdir -p gpurun_out
T=${1:-r2b}
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
L=gpurun_out/${T}_bvh.log
: > $L
timeout 240 python scripts/exp_bvh.py --only 4 --set "" --set bvh_active_min=4 --set bvh_active_min=12 --set bvh_active_min=16 --set bvh_active_min=24 \
    --set bvh_descend_min=16 --set bvh_descend_min=20 --set bvh_descend_min=28 --set bvh_descend_min=32 --set bvh_stages=1 --set bvh_stages=2 --set bvh_staged=0 >> $L 2>&1
for so in scripts/_build/exp/*.so; do [ -f $so ] && timeout 120 python scripts/exp_bvh.py --so $so --set "" >> $L 2>&1; done
grep -v "^+" $L | tail -40
# ncu on config 4 (64 spp: one pass of the staged BVH pipeline), production instances (STATS = false)
CMD="python scripts/render_once.py --variant auto --width 1920 --spp 64 --grid 158"
$CMD > gpurun_out/${T}_c4_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${T}_c4_launches.csv $CMD > gpurun_out/${T}_c4_launches.log 2>&1
for k in rz_bvh_stage_kernel rz_bvh_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${k}<\(bool\)0" -c 1 -o gpurun_out/${T}_c4_prof_${k} $CMD > gpurun_out/${T}_c4_full_${k}.log 2>&1
done
ls -la gpurun_out | grep ${T}_
