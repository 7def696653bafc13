#!/bin/bash
# Round 2, last session: final validation of the shipped build — GPU tests, smoke, both bench arms, stage timings, config 4
# (+ config 4's camera stage under the library variants left in scripts/_build/exp/).
#   gpurun --timeout 600 -- 'bash scripts/gpu_r2g.sh r2g'
set -x
T=${1:-r2g}
mkdir -p gpurun_out
md5sum rayz_b200/lib/librayz_cuda.so > gpurun_out/${T}_lib.md5
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -n 3 gpurun_out/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/${T}_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc $?"
P=gpurun_out/${T}_probe.log
timeout 60 python scripts/exp_probe.py --set "" > $P 2>&1
timeout 60 python scripts/exp_bvh.py --set "" >> $P 2>&1
timeout 60 python scripts/c4_diag.py 64 >> $P 2>&1
for so in scripts/_build/exp/*.so; do [ -f $so ] && timeout 100 python scripts/c4_diag.py --so $so 64 >> $P 2>&1; done
timeout 60 python scripts/e2e_probe.py >> $P 2>&1
grep -v "^+" $P | grep "^\[" | tail -12; tail -n 2 gpurun_out/${T}_smoke.log
cut -c1-200 gpurun_out/${T}_bench_n1.json; tail -n 3 gpurun_out/${T}_bench_n1.err
