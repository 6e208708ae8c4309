#!/bin/bash
# development build A/B: rows per thread of UpdateMatrices (1 / 2 / 4), parity with 2 and 4
set -u
O=gpurun_out
T=${1:-o}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
for r in 2 4; do
VQA_MAT_ROWS=$r timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "farneback or golden or 4k or halo or small_and_ragged or config1" > $O/r02_${T}_pytest_rows$r.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_rows$r.log
tail -2 $O/r02_${T}_pytest_rows$r.log
done
leg() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_$name.json 2> $O/r02_${T}_ab_$name.err
  python - "$O/r02_${T}_ab_$name.json" "$name" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
k = d['roofline']['kernels']
print(sys.argv[2], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], d['result']['scene_complexity'][0],
      {n: v['ms'] for n, v in k.items() if n.startswith(('k_fb_blur', 'k_fb_mat'))})
PY
}
leg rows1_1 VQA_MAT_ROWS=1
leg rows2_1 VQA_MAT_ROWS=2
leg rows4_1 VQA_MAT_ROWS=4
leg rows1_2 VQA_MAT_ROWS=1
leg rows2_2 VQA_MAT_ROWS=2
leg rows4_2 VQA_MAT_ROWS=4
