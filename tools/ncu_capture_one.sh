#!/bin/bash
# one `ncu --set full` capture: tools/ncu_capture_one.sh <tag> <name> <kernel regex> <skip> <count>
set -u
python tools/profile_step.py 24 > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$3" -s $4 -c $5 \
    -f -o gpurun_out/ncu_$1_$2 python tools/profile_step.py 24 > gpurun_out/ncu_$1_$2.log 2>&1
tail -1 gpurun_out/ncu_$1_$2.log
ls -la gpurun_out/ncu_$1_$2.ncu-rep
