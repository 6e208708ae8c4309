#!/bin/bash
# development build A/B: frames per chunk (VQA_CHUNK)
set -u
O=gpurun_out
T=${1:-q}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
leg() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_$name.json 2> $O/r02_${T}_ab_$name.err
  python - "$O/r02_${T}_ab_$name.json" "$name" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'launches/step', d['gpu_launches'] / d['steps'], d['clocks']['sm_mhz'], d['result']['scene_complexity'][0])
PY
}
leg chunk24 VQA_CHUNK=24
leg chunk48 VQA_CHUNK=48
leg chunk75 VQA_CHUNK=75
leg chunk100 VQA_CHUNK=100
leg chunk150 VQA_CHUNK=150
leg chunk48b VQA_CHUNK=48
leg chunk100b VQA_CHUNK=100
