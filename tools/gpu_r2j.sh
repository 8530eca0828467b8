#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --legs headline --no-cpu-baseline --trials 524288"
for v in dual0 dual1; do
TVF_LIBPATH=tools/_build/variants/libtvf_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:tft_stage1_solve -s 2 -c 1 -o gpurun_out/r2j_$v -f $CMD > gpurun_out/r2j_ncu_$v.log 2>&1
done
ls -la gpurun_out/r2j*
