#!/bin/bash
# ORB general-size kernels at 1080p: tests, bench leg with --orb-size, one `ncu --set full` capture of the
# level-0 k_orb_fast launch.  usage: tools/ncu_capture_orb.sh <tag>
set -u
tag=$1
timeout 200 python -m pytest tests/test_orb_gpu.py -x -q > gpurun_out/orb_test_$tag.log 2>&1; tail -2 gpurun_out/orb_test_$tag.log
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --orb-size 1920x1080 > gpurun_out/${tag}_bench_orb1080.json 2> gpurun_out/${tag}_bench_orb1080.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_orb1080.json").read())
print(d["value"], d["e2e"]["value"], d["ms_per_step"])
print({k: v for k, v in d["roofline"]["kernels"].items() if "orb" in k})
PY
export VQA_PROF_ORB=1920x1080
python tools/profile_step.py 24 > gpurun_out/prof_plain_orb.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_orb.log; exit 1; }
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_orb_fast -s 0 -c 1 \
    -f -o gpurun_out/ncu_${tag}_orb_fast python tools/profile_step.py 24 > gpurun_out/ncu_${tag}_orb_fast.log 2>&1
tail -1 gpurun_out/ncu_${tag}_orb_fast.log
ls -la gpurun_out/ncu_${tag}_orb_fast.ncu-rep
