#!/bin/bash
# Round-end refresh after the late kernel changes (transposed-gather UpdateMatrices, list-based Canny CCL):
# full GPU test suite, bench line, ncu launch list of a short bench run, `ncu --set full` of the two changed kernels.
set -u
TAG=${1:-r01_final5}
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err || { echo "bench failed"; tail -5 gpurun_out/${TAG}_bench_n1.err; exit 1; }
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n1.json').read()); r=d['roofline']
print('bench', d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel'], r['frac'], d['clocks']['reasons'])"
python bench.py --frames 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench64_plain.json 2> gpurun_out/${TAG}_bench64_plain.err || { echo "bench64 failed"; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench64.csv \
    python bench.py --frames 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/profile_step.py 24 > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
cap() {  # name regex skip count
  timeout 150 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 \
      -f -o gpurun_out/ncu_${TAG}_$1 python tools/profile_step.py 24 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$1.log
}
cap matrices_t4 k_fb_matrices_t4 6 1
cap canny_nms k_canny_nms 0 1
ls -la gpurun_out/*${TAG}*.ncu-rep | awk '{s+=$5} END {print "ncu-rep bytes", s}'
