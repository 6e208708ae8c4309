#!/bin/bash
# ncu --set full of the Farneback kernels after the float4 + float layout (one level-0 launch each), bench line with the 256-bit polyexp stores
set -u
O=gpurun_out
T=${1:-m}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "farneback or golden or 4k or halo or small_and_ragged" > $O/r02_${T}_pytest_fb.log 2>&1; echo "pytest rc=$?" >> $O/r02_${T}_pytest_fb.log; tail -3 $O/r02_${T}_pytest_fb.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_${T}_bench_c2.json 2> $O/r02_${T}_bench_c2.err; echo "c2 rc=$?"
python - "$O/r02_${T}_bench_c2.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d['roofline']
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
for k, v in list(r['kernels'].items())[:8]: print('   ', k, v['ms'], v['launches'], v['frac_of_hbm_peak'])
PY
# level 0 is the last level: of the 12 blur / 8 matrices<0> / 3 matrices<1> / 4 polyexp launches of a step take the last ones
bash tools/ncu_capture_one.sh r02$T blur "k_fb_blur_solve" 10 1
bash tools/ncu_capture_one.sh r02$T mat "k_fb_matrices" 9 2
bash tools/ncu_capture_one.sh r02$T pe "k_fb_polyexp" 3 1
for k in blur mat pe; do python tools/ncu_summary.py $O/ncu_r02${T}_$k.ncu-rep; done > $O/r02_${T}_ncu_fb_summary.txt 2>&1; cat $O/r02_${T}_ncu_fb_summary.txt
python tools/ncu_lines.py $O/ncu_r02${T}_blur.ncu-rep "k_fb_blur_solve" 45 > $O/r02_${T}_ncu_blur_lines.txt 2>&1
python tools/ncu_lines.py $O/ncu_r02${T}_mat.ncu-rep "k_fb_matrices" 30 > $O/r02_${T}_ncu_mat_lines.txt 2>&1
python tools/ncu_lines.py $O/ncu_r02${T}_pe.ncu-rep "k_fb_polyexp" 30 > $O/r02_${T}_ncu_pe_lines.txt 2>&1
