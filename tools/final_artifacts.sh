#!/bin/bash
# Round-end artifact refresh on one B200: bench lines (own arm + reference arm), ncu full captures of the
# hot kernels, and the ncu launch list of a short bench run.  Everything lands in gpurun_out/.
set -u
TAG=${1:-r01_final}
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err || { echo "bench failed"; tail -5 gpurun_out/${TAG}_bench_n1.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py --frames 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench64_plain.json 2> gpurun_out/${TAG}_bench64_plain.err || { echo "bench64 failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench64.csv \
    python bench.py --frames 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/profile_step.py 24 > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 \
      -f -o gpurun_out/ncu_${TAG}_$1 python tools/profile_step.py 24 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$1.log
}
cap blur_solve k_fb_blur_solve 9 1
cap matrices k_fb_matrices 9 2
cap polyexp k_fb_polyexp 3 1
cap pyramid k_fb_pyramid 0 4
cap dct_umma k_dct_umma 0 2
cap canny_nms k_canny_nms 0 1
cap ccl k_ccl 0 3
cap gray_hist k_gray_hist 0 1
cap psnr_ssim k_psnr_ssim 0 1
ls -la gpurun_out/*${TAG}*.ncu-rep | awk '{s+=$5} END {print "ncu-rep bytes", s}'
