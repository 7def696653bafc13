#!/bin/bash
# sanitizer pass on small renders + multi-GPU checks (run with --gpus N)
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sweep.py --width 160 --spp 4 --reps 1 --configs mega:2:16,mega:1:16,bvh:1:16,wavefront:2:16 > gpurun_out/f_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/f_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sweep.py --width 96 --spp 2 --reps 1 --configs mega:2:16,wavefront:2:16 > gpurun_out/f_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/f_racecheck.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_n$n.json 2> gpurun_out/f_bench_n$n.err
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/f_bench_n$n.json 2> gpurun_out/f_bench_n$n.err; fi
    echo "bench n=$n exit $?" >> gpurun_out/f_bench_n$n.err
  fi
done
tail -n 4 gpurun_out/f_memcheck.log gpurun_out/f_racecheck.log gpurun_out/f_pytest.log
for n in 1 2 4 8; do [ -f gpurun_out/f_bench_n$n.json ] && python -c "
import json,sys
for l in open('gpurun_out/f_bench_n$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print($n, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'])
"; done
