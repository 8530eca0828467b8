#!/bin/bash
# round 2, call C: ncu --set full of the dual stage-1 kernel (and the old one) with source correlation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --legs headline --no-cpu-baseline --trials 524288"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tft_stage1 -s 2 -c 1 -o gpurun_out/r2c_s1dual -f $CMD > gpurun_out/r2c_ncu_dual.log 2>&1
TVF_LIBPATH=tools/_build/variants/libtvf_s1old.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:tft_stage1 -s 2 -c 1 -o gpurun_out/r2c_s1old -f $CMD > gpurun_out/r2c_ncu_old.log 2>&1
ls -la gpurun_out/r2c*
tail -3 gpurun_out/r2c_ncu_dual.log
