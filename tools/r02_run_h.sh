#!/bin/bash
# product build after the fused yuv420p ingest, histogram moments and the cp.async decimating pyramid: full GPU test suite, bench, ncu of the new kernels
set -u
O=gpurun_out
T=${1:-h}
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_gpu.log
tail -6 $O/r02_${T}_pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/r02_${T}_bench_c2.json 2> $O/r02_${T}_bench_c2.err; echo "c2 rc=$?"
python - "$O/r02_${T}_bench_c2.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'res', d['result']['scene_complexity'])
for k, v in d['roofline']['kernels'].items(): print('   ', k, v['ms'], v['launches'], v['frac_of_hbm_peak'])
PY
if [ "${2:-}" = "ncu" ]; then
  bash tools/ncu_capture_one.sh r02$T ing "k_yuv420_gray_hist|k_hist_moments|k_fb_pyramid_dec" 0 4
  python tools/ncu_summary.py $O/ncu_r02${T}_ing.ncu-rep > $O/r02_${T}_ncu_ing_summary.txt 2>&1; cat $O/r02_${T}_ncu_ing_summary.txt
fi
