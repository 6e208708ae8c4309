#!/bin/bash
# fused Farneback iteration: parity tests, bench, and an A/B against the round-1 chain (development build with -DVQA_AB)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "farneback or golden or 4k or config1 or halo or small_and_ragged or device_resident or full_size" > $O/r02_b_pytest_fb.log 2>&1; echo "pytest rc=$?" >> $O/r02_b_pytest_fb.log
tail -5 $O/r02_b_pytest_fb.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_b_bench_c2.json 2> $O/r02_b_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_b_bench_c2.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"])
for k,v in list(d["roofline"]["kernels"].items())[:12]: print(k,v)
print(d["result"]["scene_complexity"][0])
PY
# A/B in one process: development build
VQA_NVCC_EXTRA="-DVQA_AB" python real-time-video-quality-analysis_b200/build.py --force > /dev/null 2>&1
for leg in 0 1; do
  VQA_FB_LEGACY=$leg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_b_ab_legacy$leg.json 2>/dev/null
  python -c "
import json; d=json.load(open('$O/r02_b_ab_legacy$leg.json')); print('legacy=$leg value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'motion',d['result']['scene_complexity'][0])"
done
for hcap in 96 180 540; do
  VQA_FI_H=$hcap timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_b_ab_fih$hcap.json 2>/dev/null
  python -c "
import json; d=json.load(open('$O/r02_b_ab_fih$hcap.json')); print('FI_H=$hcap value',round(d['value'],1),'ms',round(d['ms_per_step'],2))"
done
