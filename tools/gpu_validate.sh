#!/bin/bash
# final validation of a build: sanitizer on the late-round kernels (bulk-copy staging in moments / stage 2 / fused tail), full suite, default
# bench line, reference arm, ncu launch list + counters of the shipping kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitizer_smoke.py > gpurun_out/val_sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/val_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitizer_smoke.py > gpurun_out/val_sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/val_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitizer_smoke.py > gpurun_out/val_sanitizer_synccheck.log 2>&1; echo "synccheck rc $?" >> gpurun_out/val_sanitizer_synccheck.log
tail -n 4 gpurun_out/val_sanitizer_memcheck.log gpurun_out/val_sanitizer_racecheck.log gpurun_out/val_sanitizer_synccheck.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 8 > gpurun_out/val_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/val_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/val_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/val_bench_n1.json 2> gpurun_out/val_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/val_bench_reference.json 2> gpurun_out/val_bench_reference.err
HEAD="python bench.py --steps 2 --warmup 1 --legs headline --no-cpu-baseline --trials 524288"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/val_launches.csv $HEAD > gpurun_out/val_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tft_|candidates|pose_tail' -s 18 -c 6 -o gpurun_out/val_prof -f $HEAD > gpurun_out/val_prof.log 2>&1
cat gpurun_out/val_tests.log; tail -n 2 gpurun_out/val_smoke.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/val_bench_n1.json"))
print("value %.4g"%d["value"], "e2e", d["e2e"]["value"], {k:round(v["ms_total"],2) for k,v in d["kernels"].items()})
print("   roofline", {k:d["roofline"].get(k) for k in ("kernel","achieved","frac","fp64_pipe_active_pct")}, d["roofline_step"]["frac"])
print("   e2e variants", {k:(round(v["value"]/1e6,2), round(v["frac_of_link_ceiling"],3)) for k,v in d["e2e_variants"].items()}, d["host_link"])
print("   large_n", d["large_n"].get("value"), d["large_n"].get("roofline",{}).get("frac"), d["large_n"].get("full_pipeline"))
print("   cpu", d["cpu_baseline"], d["clocks"])
print(open("gpurun_out/val_bench_reference.json").read()[:600])
PY
tail -n 3 gpurun_out/val_bench_n1.err
