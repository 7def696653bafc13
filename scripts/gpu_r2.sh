#!/bin/bash
# Round-2 validation run: GPU parity tests, N=1 bench (both arms), ncu launch list + full captures of the staged K1's kernels.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r2.sh <tag> [pytest -k expression]'
set -x
T=${1:-r2}
mkdir -p gpurun_out
if [ -n "$2" ]; then K=(-k "$2"); else K=(); fi
timeout 1500 python -m pytest tests -m gpu -x -q "${K[@]}" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
if [ "$3" != "noncu" ]; then
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
CMD="python bench.py --steps 2 --warmup 3 --spp 40 --variant mega --no-cpu-baseline --no-e2e --no-variants --no-other-configs"
$CMD > gpurun_out/${T}_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rz_second_kernel -s 7 -c 1 -o gpurun_out/${T}_prof_second $CMD > gpurun_out/${T}_ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rz_primary_kernel -s 2 -c 1 -o gpurun_out/${T}_prof_primary $CMD > gpurun_out/${T}_ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:rz_bin_(count|scatter)_kernel" -s 20 -c 2 -o gpurun_out/${T}_prof_sort $CMD > gpurun_out/${T}_ncu_full3.log 2>&1
fi
tail -n 3 gpurun_out/${T}_pytest.log; cut -c1-400 gpurun_out/${T}_bench_n1.json; tail -n 5 gpurun_out/${T}_bench_n1.err
