set -x
T=${1:-r1g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/${T}_tests.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/${T}_bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/${T}_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"tft_stage1_kernel|tft_stage2_kernel|pose_tail_fused|candidates_kernel|tft_epipoles" -s 5 -c 5 -f -o gpurun_out/${T}_prof python bench.py --trials 524288 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_prof.log 2>&1; echo "ncu full exit $?"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/${T}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitizer_smoke.py > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/${T}_sanitizer_racecheck.log
