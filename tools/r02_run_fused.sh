#!/bin/bash
# fused Farneback iteration v2 (development build -DVQA_AB): parity tests, A/B against the chain in one process each, ncu of one level-0 launch
set -u
O=gpurun_out
T=${1:-f}
mkdir -p $O
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build.log 2>&1 || { tail -20 $O/r02_${T}_build.log; exit 1; }
VQA_FB_FUSED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "farneback or golden or 4k or config1 or halo or small_and_ragged or device_resident or full_size" > $O/r02_${T}_pytest_fused.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_fused.log
tail -5 $O/r02_${T}_pytest_fused.log
for leg in 0 1; do
  VQA_FB_FUSED=$leg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_fused$leg.json 2>$O/r02_${T}_ab_fused$leg.err
  python - <<PY
import json
d=json.load(open('$O/r02_${T}_ab_fused$leg.json'))
print('fused=$leg value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'motion',d['result']['scene_complexity'][0])
for k,v in list(d['roofline']['kernels'].items())[:6]: print('   ',k,v)
PY
done
if [ "${2:-}" = "ncu" ]; then
  VQA_FB_FUSED=1 bash tools/ncu_capture_one.sh r02$T fi "k_fb_iter" 9 3
  python tools/ncu_summary.py $O/ncu_r02${T}_fi.ncu-rep > $O/r02_${T}_ncu_fi_summary.txt 2>&1; cat $O/r02_${T}_ncu_fi_summary.txt
  python tools/ncu_lines.py $O/ncu_r02${T}_fi.ncu-rep "k_fb_iter<(int)0>" 30 > $O/r02_${T}_ncu_fi_lines.txt 2>&1
fi
