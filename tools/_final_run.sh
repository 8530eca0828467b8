set -x
T=${1:-r1g}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/${T}_tests.log | cut -c1-300
# ncu --set full of one launch of every kernel of the TFT step (524 288 problems per launch), summarised on the box so
# that the bench line below carries the counters of THIS build
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"tft_stage1_kernel|tft_stage2|pose_tail_fused|candidates_kernel|tft_epipoles" -s 5 -c 5 -f -o gpurun_out/${T}_prof python bench.py --trials 524288 --steps 1 --warmup 1 --no-cpu-baseline --legs headline > gpurun_out/${T}_prof.log 2>&1; echo "ncu full exit $?"
python tools/ncu_summary.py --rep gpurun_out/${T}_prof.ncu-rep --out profiles/r01 --problems-per-launch 524288; echo "summary exit $?"
cp profiles/r01_kernels.md gpurun_out/${T}_kernels.md; cp profiles/r01_counters.json gpurun_out/${T}_counters.json
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/${T}_bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/${T}_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --legs headline > gpurun_out/${T}_launches.log 2>&1; echo "launch list exit $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sweep_seeds_warp -s 2 -c 1 -f -o gpurun_out/${T}_prof_gen python tools/gen_only.py > gpurun_out/${T}_prof_gen.log 2>&1; echo "ncu gen exit $?"
python tools/ncu_summary.py --rep gpurun_out/${T}_prof_gen.ncu-rep --out gpurun_out/${T}_gen --problems-per-launch 1000000; echo "gen summary exit $?"
timeout 300 compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/${T}_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python tools/sanitizer_smoke.py > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/${T}_sanitizer_racecheck.log
ls -la gpurun_out
