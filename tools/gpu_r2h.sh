#!/bin/bash
# round 2, call H: staged pageable path, full suite, default bench line, ncu counters of the current kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2h_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
HEAD="python bench.py --steps 2 --warmup 1 --legs headline --no-cpu-baseline --trials 524288"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv $HEAD > gpurun_out/r2h_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tft_|candidates|pose_tail' -s 18 -c 6 -o gpurun_out/r2h_prof -f $HEAD > gpurun_out/r2h_prof.log 2>&1
LG1="python bench.py --workload large-n --n 10000 --trials 4096 --steps 1 --warmup 1"
timeout 600 ncu --set full --clock-control none -k regex:'tft_moments_large|votes_kernel|scale_large|final_large' -s 4 -c 4 -o gpurun_out/r2h_prof_large -f $LG1 > gpurun_out/r2h_prof_large.log 2>&1
cat gpurun_out/r2h_tests.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_n1.json"))
print("value %.4g"%d["value"], "e2e", d["e2e"]["value"], {k:round(v["ms_total"],2) for k,v in d["kernels"].items()})
print("   roofline", {k:d["roofline"].get(k) for k in ("kernel","achieved","frac","fp64_pipe_active_pct")}, d["roofline_step"]["frac"])
print("   e2e variants", {k:(round(v["value"]/1e6,2), round(v["frac_of_link_ceiling"],3)) for k,v in d["e2e_variants"].items()}, d["host_link"])
print("   large_n", d["large_n"].get("value"), d["large_n"].get("roofline",{}).get("frac"), d["large_n"].get("full_pipeline"))
print("   cpu", d["cpu_baseline"], d["clocks"])
print(open("gpurun_out/r2h_bench_reference.json").read()[:600])
PY
tail -3 gpurun_out/r2h_bench_n1.err
