#!/bin/bash
# development build A/B: side stream (Canny / ORB / DCT) at the higher stream priority
set -u
O=gpurun_out
T=${1:-p}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
VQA_SIDE_PRIO=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_yuv_gpu.py -m gpu -x -q > $O/r02_${T}_pytest_prio.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_prio.log
tail -3 $O/r02_${T}_pytest_prio.log
leg() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_$name.json 2> $O/r02_${T}_ab_$name.err
  python - "$O/r02_${T}_ab_$name.json" "$name" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_mean'), d['result']['scene_complexity'][0])
PY
}
leg prio0_1 VQA_SIDE_PRIO=0
leg prio1_1 VQA_SIDE_PRIO=1
leg prio0_2 VQA_SIDE_PRIO=0
leg prio1_2 VQA_SIDE_PRIO=1
leg prio1_dual VQA_SIDE_PRIO=1 VQA_FB_DUAL=1
VQA_SIDE_PRIO=1 timeout 300 python bench.py --workload c4 --frames 240 --steps 2 --no-cpu-baseline > $O/r02_${T}_ab_prio1_c4.json 2>/dev/null
VQA_SIDE_PRIO=0 timeout 300 python bench.py --workload c4 --frames 240 --steps 2 --no-cpu-baseline > $O/r02_${T}_ab_prio0_c4.json 2>/dev/null
python - <<PY
import json
for leg in (0, 1):
    d = json.load(open('$O/r02_${T}_ab_prio%d_c4.json' % leg))
    print('c4 240 frames prio=%d' % leg, round(d['value'], 1), round(d['ms_per_step'], 1), d['clocks']['sm_mhz'])
PY
