#!/bin/bash
# ncu --set full of the linear-F / optimal-F kernels (one 524 288-problem launch each): f_stage1_kernel, f_finish_kernel, optimf_gh_kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ALL="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --trials 524288 --large-n-scenes 0"
timeout 400 ncu --set full --clock-control none -k regex:'f_stage1_kernel|f_finish_kernel' -s 2 -c 2 -o gpurun_out/fpath_prof -f $ALL > gpurun_out/fpath_prof.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'optimf_gh' -s 1 -c 1 -o gpurun_out/fpath_prof_gh -f $ALL > gpurun_out/fpath_prof_gh.log 2>&1
ls -la gpurun_out/fpath_prof*
