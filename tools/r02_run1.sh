#!/bin/bash
# round 2, GPU call 2: new tests, then the bench workloads at N=1 (c2 full; c3; c1; reduced c4 / c5 for a first look)
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_a_pytest_gpu.log
tail -5 $O/r02_a_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02_a_bench_c2.json 2> $O/r02_a_bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --workload c3 > $O/r02_a_bench_c3.json 2> $O/r02_a_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload c1 > $O/r02_a_bench_c1.json 2> $O/r02_a_bench_c1.err; echo "c1 rc=$?"
timeout 600 python bench.py --workload c4 --frames 240 --steps 2 --no-cpu-baseline > $O/r02_a_bench_c4_240.json 2> $O/r02_a_bench_c4_240.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c5 --clips 4 --frames 200 --steps 2 > $O/r02_a_bench_c5_small.json 2> $O/r02_a_bench_c5_small.err; echo "c5 rc=$?"
for f in $O/r02_a_bench_*.json; do echo "== $f"; cut -c1-400 $f; done
for f in $O/r02_a_bench_*.err; do echo "== $f"; tail -5 $f; done
