#!/bin/bash
# development build A/B: blur that evaluates the next UpdateMatrices (VQA_FB_NEXT) and the two-stream schedule (VQA_FB_DUAL), parity with both on
set -u
O=gpurun_out
T=${1:-n}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
VQA_FB_NEXT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_yuv_gpu.py -m gpu -x -q > $O/r02_${T}_pytest_next.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_next.log
tail -4 $O/r02_${T}_pytest_next.log
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
leg base_1 VQA_FB_NEXT=0
leg next_1 VQA_FB_NEXT=1
leg dual_1 VQA_FB_DUAL=1
leg nextdual_1 VQA_FB_NEXT=1 VQA_FB_DUAL=1
leg base_2 VQA_FB_NEXT=0
leg next_2 VQA_FB_NEXT=1
leg dual_2 VQA_FB_DUAL=1
