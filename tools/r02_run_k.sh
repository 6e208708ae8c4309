#!/bin/bash
# two-stream staggered Farneback schedule (development build): parity tests with it on, then A/B in one process each
set -u
O=gpurun_out
T=${1:-k}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
VQA_FB_DUAL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_yuv_gpu.py -m gpu -x -q > $O/r02_${T}_pytest_dual.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_dual.log
tail -4 $O/r02_${T}_pytest_dual.log
for rep in 1 2; do for leg in 0 1; do
  VQA_FB_DUAL=$leg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_dual${leg}_$rep.json 2> $O/r02_${T}_ab_dual${leg}_$rep.err
  python - "$O/r02_${T}_ab_dual${leg}_$rep.json" "dual=$leg rep=$rep" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_mean'), d['result']['scene_complexity'][0])
PY
done; done
VQA_FB_DUAL=1 timeout 300 python bench.py --workload c4 --frames 240 --steps 2 --no-cpu-baseline > $O/r02_${T}_ab_dual1_c4.json 2>/dev/null
VQA_FB_DUAL=0 timeout 300 python bench.py --workload c4 --frames 240 --steps 2 --no-cpu-baseline > $O/r02_${T}_ab_dual0_c4.json 2>/dev/null
python - <<PY
import json
for leg in (0, 1):
    d = json.load(open('$O/r02_${T}_ab_dual%d_c4.json' % leg))
    print('c4 240 frames dual=%d' % leg, round(d['value'], 1), round(d['ms_per_step'], 1), d['clocks']['sm_mhz'])
PY
