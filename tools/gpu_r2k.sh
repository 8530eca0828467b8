#!/bin/bash
# ncu --set full of the two-problem stage-1 solver and of the fused tail, shipping build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --legs headline --no-cpu-baseline --trials 524288"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tft_stage1_solve -s 2 -c 1 -o gpurun_out/r2k_solve -f $CMD > gpurun_out/r2k_ncu_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pose_tail_fused -s 2 -c 1 -o gpurun_out/r2k_tail -f $CMD > gpurun_out/r2k_ncu_tail.log 2>&1
ls -la gpurun_out/r2k*
