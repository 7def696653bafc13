#!/bin/bash
# First GPU validation: smoke, parity tests, sweep, bench (both arms), ncu launch list + full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python scripts/sweep.py > gpurun_out/sweep.log 2>&1
timeout 300 python scripts/sweep.py --glass --configs mega:2:16,bvh:1:16 > gpurun_out/sweep_glass.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# ncu: launch list, then one full capture of the path kernel (same command, run plain first)
CMD="python bench.py --steps 2 --warmup 3 --spp 40 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rz_path_kernel -s 2 -c 1 -o gpurun_out/prof_path $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/smoke.log gpurun_out/pytest.log gpurun_out/sweep.log gpurun_out/bench_n1.json
