#!/bin/bash
# Second GPU run (2 GPUs): multi-GPU tests, N=2 bench under torchrun, rayz_host, other configs.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/b_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err; echo "exit $?" >> gpurun_out/b_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/b_bench_ref_n2.json 2> gpurun_out/b_bench_ref_n2.err
make -C host > gpurun_out/b_host_build.log 2>&1
( time host/_build/rayz_host 1200 gpurun_out/b_out_1200.ppm --spp 500 --seed 42 ) > gpurun_out/b_host_1gpu.log 2>&1
( time host/_build/rayz_host 3840 /tmp/out4k.ppm --spp 100 --seed 42 --gpus 2 ) > gpurun_out/b_host_2gpu.log 2>&1
head -c 200 gpurun_out/b_out_1200.ppm > gpurun_out/b_out_head.txt; md5sum gpurun_out/b_out_1200.ppm >> gpurun_out/b_out_head.txt; rm -f gpurun_out/b_out_1200.ppm
# config 3 on one GPU (4K, fewer spp), config 4 (100k spheres, BVH), config 5 (glass)
timeout 300 python scripts/sweep.py --width 3840 --spp 50 --reps 2 --configs mega:2:16,bvh:1:16 > gpurun_out/b_sweep_4k.log 2>&1
timeout 600 python scripts/sweep.py --width 1920 --spp 32 --reps 2 --grid 158 --configs bvh:1:16 > gpurun_out/b_sweep_100k.log 2>&1
timeout 300 python scripts/sweep.py --width 1200 --spp 100 --reps 2 --glass --configs mega:2:16,bvh:1:16,wavefront:2:16 > gpurun_out/b_sweep_glass.log 2>&1
tail -n 3 gpurun_out/b_pytest.log gpurun_out/b_bench_n2.json gpurun_out/b_host_1gpu.log gpurun_out/b_host_2gpu.log
