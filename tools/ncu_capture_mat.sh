#!/bin/bash
set -u
TAG=${1:-r01_final}
python tools/profile_step.py 24 > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
cap() {
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 \
      -f -o gpurun_out/ncu_${TAG}_$1 python tools/profile_step.py 24 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$1.log
}
cap matrices k_fb_matrices 9 2
