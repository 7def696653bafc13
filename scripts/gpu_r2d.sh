#!/bin/bash
# Round 2, last session, validation of the build with 8 resident CTAs for the BVH kernels and the faster host SAH build:
# GPU tests, both bench arms, BVH tuning re-sweep at the new occupancy, e2e probe, occupancy variants of the staged K1's kernels
# (scripts/_build/exp/), ncu launch list of the bench command and one --set full capture of the BVH kernel on config 4.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2d.sh r2d'
set -x
T=${1:-r2d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -n 3 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc $?"
timeout 60 python scripts/e2e_probe.py > gpurun_out/${T}_e2e.log 2>&1
L=gpurun_out/${T}_bvh.log
timeout 200 python scripts/exp_bvh.py --only 4 --set "" --set bvh_active_min=6 --set bvh_active_min=12 --set bvh_descend_min=20 --set bvh_descend_min=28 > $L 2>&1
timeout 60 python scripts/exp_bvh.py --only 2 --set "" >> $L 2>&1
P=gpurun_out/${T}_probe.log
timeout 60 python scripts/exp_probe.py --set "" > $P 2>&1
for so in scripts/_build/exp/*.so; do [ -f $so ] && timeout 60 python scripts/exp_probe.py --so $so --set "" >> $P 2>&1; done
grep -v "^+" $L $P | tail -20
if [ "$2" != "noncu" ]; then
CMD="python bench.py --steps 2 --warmup 3 --spp 40 --variant mega --no-cpu-baseline --no-e2e --no-variants --no-other-configs"
$CMD > gpurun_out/${T}_ncu_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1
CMD4="python scripts/render_once.py --variant auto --width 1920 --spp 64 --grid 158"
$CMD4 > gpurun_out/${T}_c4_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rz_bvh_kernel<\(bool\)0" -c 1 -o gpurun_out/${T}_c4_prof_rz_bvh_kernel $CMD4 > gpurun_out/${T}_c4_full.log 2>&1
fi
cut -c1-300 gpurun_out/${T}_bench_n1.json; tail -n 3 gpurun_out/${T}_bench_n1.err; cat gpurun_out/${T}_e2e.log | tail -n 1
