#!/bin/bash
# round 2, call D: full GPU suite, full bench line, ncu launch list + full counters of the headline kernels,
# large-n (config 5) DFMA vs DMMA variants with ncu counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
HEAD="python bench.py --steps 2 --warmup 1 --legs headline --no-cpu-baseline --trials 524288"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches.csv $HEAD > gpurun_out/r2d_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tft_|candidates|pose_tail' -s 18 -c 6 -o gpurun_out/r2d_prof -f $HEAD > gpurun_out/r2d_prof.log 2>&1
# config 5: DFMA (shipping) and DMMA (tensor-core) Gram kernels
LG="python bench.py --workload large-n --n 10000 --trials 8192 --steps 3 --warmup 1"
timeout 600 $LG > gpurun_out/r2d_large_dfma.json 2> gpurun_out/r2d_large_dfma.err
TVF_LIBPATH=tools/_build/variants/libtvf_lgdmma.so timeout 600 $LG > gpurun_out/r2d_large_dmma.json 2> gpurun_out/r2d_large_dmma.err
TVF_LIBPATH=tools/_build/variants/libtvf_lgdmma.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "large_n" 2>&1 | tail -4 > gpurun_out/r2d_tests_dmma.log
LG1="python bench.py --workload large-n --n 10000 --trials 4096 --steps 1 --warmup 1"
timeout 600 ncu --set full --clock-control none -k regex:tft_moments_large -s 1 -c 1 -o gpurun_out/r2d_prof_large_dfma -f $LG1 > gpurun_out/r2d_prof_large_dfma.log 2>&1
TVF_LIBPATH=tools/_build/variants/libtvf_lgdmma.so timeout 600 ncu --set full --clock-control none -k regex:tft_moments_large -s 1 -c 1 -o gpurun_out/r2d_prof_large_dmma -f $LG1 > gpurun_out/r2d_prof_large_dmma.log 2>&1
cat gpurun_out/r2d_tests.log gpurun_out/r2d_tests_dmma.log
python - <<'PY'
import json
for f in ("r2d_bench_n1","r2d_large_dfma","r2d_large_dmma"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.4g"%d["value"], "e2e", (d.get("e2e") or {}).get("value"), {k:round(v["ms_total"],2) for k,v in d["kernels"].items()})
        if "roofline" in d: print("   roofline", {k:d["roofline"].get(k) for k in ("kernel","achieved","frac","fp64_pipe_active_pct")})
        if d.get("e2e_variants"): print("   e2e variants", {k:(round(v["value"]/1e6,2), round(v["frac_of_link_ceiling"],3)) for k,v in d["e2e_variants"].items()}, d["host_link"])
        if d.get("large_n"): print("   large_n", d["large_n"].get("value"), d["large_n"].get("roofline",{}).get("frac"), d["large_n"].get("full_pipeline"))
        if d.get("full_pipeline"): print("   full", d["full_pipeline"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2d_bench_n1.err
