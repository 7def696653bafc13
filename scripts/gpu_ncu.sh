#!/bin/bash
# ncu captures of the staged K1's kernels on a 40-spp config-2 render (one pass): launch list + --set full of one launch each.
#   gpurun -- 'bash scripts/gpu_ncu.sh <tag> [kernel-regex ...]'
set -x
T=${1:-ncu}; shift
SPP=${SPP:-40}            # SPP=500: the launches of the benchmark configuration itself (four passes)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --spp $SPP --variant mega --no-cpu-baseline --no-e2e --no-variants --no-other-configs"
$CMD > gpurun_out/${T}_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_launches.log 2>&1
for K in "${@:-rz_second_kernel rz_primary_kernel}"; do
  for k in $K; do
    # the production instances only (template argument STATS = false; the bench's stats render launches the <true> ones);
    # rz_second_kernel: seven launches per 40-spp render, so 14 skipped = the FIRST sorted stage of the third render
    S=2; [ "$k" == "rz_second_kernel" ] && S=${SKIP_SECOND:-14}
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${k}<\(bool\)0" -s $S -c 1 -o gpurun_out/${T}_prof_${k} $CMD > gpurun_out/${T}_full_${k}.log 2>&1
  done
done
ls -la gpurun_out | grep ${T}_
