#!/bin/bash
# development build A/B: DCT GEMM 2 on 128 x 192 tiles with 32-element K stages (VQA_DCT_G2W)
set -u
O=gpurun_out
T=${1:-t2}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
VQA_DCT_G2W=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dct or golden or config1 or full_size or 4k or small_and_ragged" > $O/r02_${T}_pytest_g2w.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_g2w.log
tail -12 $O/r02_${T}_pytest_g2w.log
leg() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_$name.json 2> $O/r02_${T}_ab_$name.err
  python - "$O/r02_${T}_ab_$name.json" "$name" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
k = d['roofline']['kernels']
print(sys.argv[2], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], d['result']['scene_complexity'][1], d['result']['scene_complexity'][6],
      {n: (v['ms'], v['TFLOPs']) for n, v in k.items() if n.startswith('k_dct')})
PY
}
leg g2w0_1 VQA_DCT_G2W=0
leg g2w1_1 VQA_DCT_G2W=1
leg g2w0_2 VQA_DCT_G2W=0
leg g2w1_2 VQA_DCT_G2W=1
