#!/bin/bash
# blur v2 (240-column strips, one barrier per row) + decimating pyramid kernel: parity tests on the product build, then the
# development build's A/B legs (strip width, strip height cap, pair groups for L2 residency of M), then ncu of the changed kernels
set -u
O=gpurun_out
T=${1:-g}
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_gpu.log
tail -4 $O/r02_${T}_pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/r02_${T}_bench_c2.json 2> $O/r02_${T}_bench_c2.err; echo "c2 rc=$?"
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
k = d['roofline']['kernels']
pick = {n: v['ms'] for n, v in k.items() if n.startswith('k_fb_')}
print(sys.argv[2], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'motion', d['result']['scene_complexity'][0])
print('   ', pick)
PY
}
show $O/r02_${T}_bench_c2.json product
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > $O/r02_${T}_build_ab.log 2>&1 || { tail -20 $O/r02_${T}_build_ab.log; exit 1; }
leg() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_${T}_ab_$name.json 2> $O/r02_${T}_ab_$name.err
  show $O/r02_${T}_ab_$name.json $name
}
leg w128 VQA_BLUR_W=128
leg h360 VQA_MS_H=360
leg h180 VQA_MS_H=180
leg grp1 VQA_FB_GROUP=1
leg grp2 VQA_FB_GROUP=2
leg grp4 VQA_FB_GROUP=4
leg grp8 VQA_FB_GROUP=8
python real-time-video-quality-analysis_b200/build.py --force > /dev/null 2>&1
if [ "${2:-}" = "ncu" ]; then
  bash tools/ncu_capture_one.sh r02$T blur "k_fb_blur_solve" 9 2
  python tools/ncu_summary.py $O/ncu_r02${T}_blur.ncu-rep > $O/r02_${T}_ncu_blur_summary.txt 2>&1; cat $O/r02_${T}_ncu_blur_summary.txt
  bash tools/ncu_capture_one.sh r02$T pyr "k_fb_pyramid" 0 4
  python tools/ncu_summary.py $O/ncu_r02${T}_pyr.ncu-rep > $O/r02_${T}_ncu_pyr_summary.txt 2>&1; cat $O/r02_${T}_ncu_pyr_summary.txt
fi
