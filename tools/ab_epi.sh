#!/bin/bash
# A/B of the fused blur+UpdateMatrices epilogue (VQA_FB_EPI) on one box, alternating runs
for i in 1 2; do
  for e in 1 0; do
    VQA_FB_EPI=$e python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_epi_${e}_$i.json 2> gpurun_out/ab_epi_${e}_$i.err
  done
done
