#!/bin/bash
# ncu --set full of the large-n kernels (BASELINE config 5, n = 10 000): Gram formation and the three un-fused tail kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LG1="python bench.py --workload large-n --n 10000 --trials 4096 --steps 1 --warmup 1"
timeout 600 ncu --set full --clock-control none -k regex:'tft_moments_large|votes_kernel|scale_large|final_large' -s 4 -c 4 -o gpurun_out/prof_large -f $LG1 > gpurun_out/prof_large.log 2>&1
ls -la gpurun_out/prof_large*
# summary: python tools/ncu_summary.py --rep gpurun_out/prof_large.ncu-rep --out profiles/rNN_large_n --problems-per-launch 4096
