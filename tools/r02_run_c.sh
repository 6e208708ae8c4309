#!/bin/bash
# ncu --set full of the Canny and ingest kernels of the final build (per-line instruction / stall shares)
set -u
O=gpurun_out
T=${1:-cn}
mkdir -p $O
bash tools/ncu_capture_one.sh r02$T canny "k_canny_nms|k_ccl_merge|k_yuv420_gray_hist|k_fb_pyramid3" 0 5
python tools/ncu_summary.py $O/ncu_r02${T}_canny.ncu-rep > $O/r02_${T}_ncu_summary.txt 2>&1; cat $O/r02_${T}_ncu_summary.txt
python tools/ncu_lines.py $O/ncu_r02${T}_canny.ncu-rep "k_canny_nms" 40 > $O/r02_${T}_ncu_canny_lines.txt 2>&1
python tools/ncu_lines.py $O/ncu_r02${T}_canny.ncu-rep "k_fb_pyramid3<(int)0>" 20 > $O/r02_${T}_ncu_pyr3_lines.txt 2>&1
python tools/ncu_phases.py $O/ncu_r02${T}_canny.ncu-rep > $O/r02_${T}_ncu_phases.txt 2>&1
