#!/bin/bash
# Scaling run on one box: bench.py at N = 1, 2, 4, 8 (as the driver launches it) + the reference arm at N=8.
set -x
mkdir -p gpurun_out
NMAX=${1:-8}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.log
for n in 1 2 4 8; do
  [ $n -le $NMAX ] || continue
  if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/g_bench_n$n.json 2> gpurun_out/g_bench_n$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/g_bench_n$n.json 2> gpurun_out/g_bench_n$n.err; fi
  echo "bench n=$n exit $?" >> gpurun_out/g_bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29540 bench.py --impl reference --gpus $NMAX --steps 1 --warmup 0 > gpurun_out/g_bench_ref_n$NMAX.json 2> gpurun_out/g_bench_ref.err
( time host/_build/rayz_host 3840 /tmp/out4k.ppm --spp 100 --seed 42 --gpus $NMAX ) > gpurun_out/g_host_ngpu.log 2>&1
tail -n 2 gpurun_out/g_pytest.log gpurun_out/g_host_ngpu.log
