#!/bin/bash
# round 2, call G: certified vote signs + large-n tail kernels: full parity suite, headline, large-n, against the previous build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2g_tests.log
for v in base nofast fastonly; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --legs headline --no-cpu-baseline > gpurun_out/r2g_bench_$v.json 2> gpurun_out/r2g_bench_$v.err
  timeout 300 python bench.py --workload large-n --n 10000 --trials 8192 --steps 3 --warmup 1 > gpurun_out/r2g_large_$v.json 2> gpurun_out/r2g_large_$v.err
done
unset TVF_LIBPATH
cat gpurun_out/r2g_tests.log
python - <<'PY'
import json
for f in ("base","nofast","fastonly"):
    for w in ("bench","large"):
        try:
            d=json.load(open("gpurun_out/r2g_%s_%s.json"%(w,f)))
            print(w, f, "value %.4g"%d["value"], {k:round(v["ms_total"],2) for k,v in d["kernels"].items()}, d.get("full_pipeline",{}).get("value"), "flagged", d["flagged_problems"])
        except Exception as e:
            print(w, f, "failed", e)
PY
