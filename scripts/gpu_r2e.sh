#!/bin/bash
# Round 2, last session: validation of the current build (GPU tests, bench) + stage timings of the library variants under
# scripts/_build/exp/ (exp_probe.py: the staged K1 on config 2, standard and glass-heavy scene).
#   gpurun --timeout 600 -- 'bash scripts/gpu_r2e.sh r2e'
set -x
T=${1:-r2e}
mkdir -p gpurun_out
md5sum rayz_b200/lib/librayz_cuda.so > gpurun_out/${T}_lib.md5
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -n 3 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc $?"
P=gpurun_out/${T}_probe.log
timeout 60 python scripts/exp_probe.py --set "" > $P 2>&1
timeout 60 python scripts/exp_probe.py --glass --set "" >> $P 2>&1
for so in scripts/_build/exp/*.so; do [ -f $so ] && timeout 60 python scripts/exp_probe.py --so $so --set "" >> $P 2>&1; done
grep -v "^+" $P | tail -20
cut -c1-300 gpurun_out/${T}_bench_n1.json; tail -n 3 gpurun_out/${T}_bench_n1.err
