#!/bin/bash
# Multi-GPU validation: the multi-device parity test, then bench.py launched the way the driver launches it.
#   gpurun --gpus N -- 'bash scripts/gpu_multi.sh <tag> N'
set -x
T=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_gpu.py -x -q -k "multi_device or cpp_host" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
( time host/_build/rayz_host 3840 gpurun_out/${T}_out.ppm --spp 100 --seed 42 --gpus $N --ppm-bench ) > gpurun_out/${T}_host.log 2>&1; rm -f gpurun_out/${T}_out.ppm
tail -n 4 gpurun_out/${T}_pytest.log; cut -c1-600 gpurun_out/${T}_bench_n$N.json; tail -n 5 gpurun_out/${T}_bench_n$N.err; tail -n 8 gpurun_out/${T}_host.log
