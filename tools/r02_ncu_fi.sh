#!/bin/bash
# ncu --set full of the fused Farneback iteration at level 0 (launches 9..11 of the 12 k_fb_iter launches of one step)
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "farneback or golden or full_size or halo" > $O/r02_c_pytest_fb.log 2>&1; echo "pytest rc=$?" >> $O/r02_c_pytest_fb.log; tail -3 $O/r02_c_pytest_fb.log
bash tools/ncu_capture_one.sh r02c fi "k_fb_iter" 9 3
python tools/ncu_summary.py $O/ncu_r02c_fi.ncu-rep > $O/r02_c_ncu_fi_summary.txt 2>&1; cat $O/r02_c_ncu_fi_summary.txt
python tools/ncu_phases.py $O/ncu_r02c_fi.ncu-rep 47692800 > $O/r02_c_ncu_fi_phases.txt 2>&1; cat $O/r02_c_ncu_fi_phases.txt
ncu -i $O/ncu_r02c_fi.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Issue Slots|Eligible|No Eligible|Theoretical Occ|Achieved Occ|L1/TEX Hit|Mem Busy|Max Bandwidth|Bank" | head -60 > $O/r02_c_ncu_fi_details.txt
