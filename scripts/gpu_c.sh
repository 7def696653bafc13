#!/bin/bash
# Validation run after a kernel change: parity tests, N=1 bench (both arms), host stand-in, ncu launch list + full capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/e_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err
( time host/_build/rayz_host 1200 gpurun_out/e_out_1200.ppm --spp 500 --seed 42 ) > gpurun_out/e_host_1gpu.log 2>&1
md5sum gpurun_out/e_out_1200.ppm >> gpurun_out/e_host_1gpu.log; rm -f gpurun_out/e_out_1200.ppm
CMD="python bench.py --steps 2 --warmup 3 --spp 40 --variant mega --no-cpu-baseline --no-e2e --no-variants"
$CMD > gpurun_out/e_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/e_launches.csv $CMD > gpurun_out/e_ncu_launches.log 2>&1
$CMD > gpurun_out/e_ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rz_bvh_kernel -s 2 -c 1 -o gpurun_out/e_prof_tail $CMD > gpurun_out/e_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rz_primary_kernel -s 2 -c 1 -o gpurun_out/e_prof_primary $CMD > gpurun_out/e_ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rz_second_kernel -s 10 -c 1 -o gpurun_out/e_prof_second $CMD > gpurun_out/e_ncu_full3.log 2>&1
tail -n 3 gpurun_out/e_pytest.log gpurun_out/e_host_1gpu.log; cut -c1-300 gpurun_out/e_bench_n1.json
