#!/bin/bash
set -x
mkdir -p gpurun_out
CMD1="python scripts/sweep.py --width 1200 --spp 16 --reps 1 --configs bvh:1:16"
CMD2="python scripts/sweep.py --width 1920 --spp 8 --reps 1 --grid 158 --configs bvh:1:16"
$CMD1 > gpurun_out/i_bvh_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rz_bvh_kernel -s 1 -c 1 -o gpurun_out/k_prof_bvh_small $CMD1 > gpurun_out/i_ncu1.log 2>&1
$CMD2 > gpurun_out/i_bvh_100k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rz_bvh_kernel -s 1 -c 1 -o gpurun_out/k_prof_bvh_100k $CMD2 > gpurun_out/i_ncu2.log 2>&1
tail -2 gpurun_out/i_bvh_small.log gpurun_out/i_bvh_100k.log
