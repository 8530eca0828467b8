#!/bin/bash
# round 2, call B: dual stage 1 -- parity + timing against the old kernel on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2b_tests.log
for v in base nopid; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --legs headline --no-cpu-baseline > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err
done
unset TVF_LIBPATH
cat gpurun_out/r2b_tests.log; ./tools/_build/probe_clusters > gpurun_out/r2b_probe_clusters.txt 2>&1
python - <<'PY'
import json
for f in ("base","nopid"):
    try:
        d=json.load(open("gpurun_out/r2b_bench_%s.json"%f))
        print(f, "value %.4g"%d["value"], {k:round(v["ms_total"],2) for k,v in d["kernels"].items()}, "flagged", d["flagged_problems"])
    except Exception as e:
        print(f, "failed", e)
PY
