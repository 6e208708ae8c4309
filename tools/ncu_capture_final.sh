#!/bin/bash
# Final round captures: one `ncu --set full` per hot kernel (level-0 launch for the Farneback kernels)
set -u
TAG=${1:-r01_final}
python tools/profile_step.py 24 > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 \
      -f -o gpurun_out/ncu_${TAG}_$1 python tools/profile_step.py 24 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$1.log
}
cap blur_solve k_fb_blur_solve 9 1
cap matrices k_fb_matrices 9 2
cap polyexp k_fb_polyexp 3 1
cap dct_umma k_dct_umma 0 2
cap canny_nms k_canny_nms 0 1
cap gray_hist k_gray_hist 0 1
cap psnr_ssim k_psnr_ssim 0 1
ls -la gpurun_out/*${TAG}*.ncu-rep
