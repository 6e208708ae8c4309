#!/bin/bash
# development build A/B: side chain (Canny / ORB / DCT) launched at level 0 of the Farneback chain, with and without the higher stream priority
set -u
O=gpurun_out
T=${1:-r}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
VQA_SIDE_AT_L0=1 VQA_SIDE_PRIO=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_yuv_gpu.py -m gpu -x -q > $O/r02_${T}_pytest_l0.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_l0.log
tail -3 $O/r02_${T}_pytest_l0.log
leg() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_$name.json 2> $O/r02_${T}_ab_$name.err
  python - "$O/r02_${T}_ab_$name.json" "$name" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_mean'), d['result']['scene_complexity'][0])
PY
}
leg base_1 VQA_SIDE_AT_L0=0
leg l0_1 VQA_SIDE_AT_L0=1
leg l0prio_1 VQA_SIDE_AT_L0=1 VQA_SIDE_PRIO=1
leg base_2 VQA_SIDE_AT_L0=0
leg l0_2 VQA_SIDE_AT_L0=1
leg l0prio_2 VQA_SIDE_AT_L0=1 VQA_SIDE_PRIO=1
