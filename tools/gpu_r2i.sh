#!/bin/bash
# round 2, call I: compute-sanitizer (memcheck, racecheck, synccheck) on the smoke script
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitizer_smoke.py > gpurun_out/r2i_sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/r2i_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitizer_smoke.py > gpurun_out/r2i_sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/r2i_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitizer_smoke.py > gpurun_out/r2i_sanitizer_synccheck.log 2>&1; echo "synccheck rc $?" >> gpurun_out/r2i_sanitizer_synccheck.log
tail -4 gpurun_out/r2i_sanitizer_memcheck.log gpurun_out/r2i_sanitizer_racecheck.log gpurun_out/r2i_sanitizer_synccheck.log
