#!/bin/bash
# ncu --set full of all six kernels of the TFT step, current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
HEAD="python bench.py --steps 2 --warmup 1 --legs headline --no-cpu-baseline --trials 524288"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tft_|candidates|pose_tail' -s 18 -c 6 -o gpurun_out/ncu_step_prof -f $HEAD > gpurun_out/ncu_step_prof.log 2>&1
ls -la gpurun_out/r2l*
