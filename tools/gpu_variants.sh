#!/bin/bash
# Compile-time variants against the shipping build on ONE box: the full GPU suite on the shipping library, then the headline leg
# of the bench for each variant.  Usage (under gpurun): bash tools/gpu_variants.sh name1 name2 ...   where
# tools/_build/variants/libtvf_<name>.so was built by `python tools/build_variants.py name=-DFLAG=1,-DOTHER=2 ...`
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 8 > gpurun_out/var_tests.log
for v in base "$@"; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --legs headline --no-cpu-baseline > gpurun_out/var_bench_$v.json 2> gpurun_out/var_bench_$v.err
done
unset TVF_LIBPATH
cat gpurun_out/var_tests.log
python - base "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open("gpurun_out/var_bench_%s.json" % f))
        print(f, "value %.4g" % d["value"], {k: round(v["ms_total"], 2) for k, v in d["kernels"].items()}, "flagged", d["flagged_problems"])
    except Exception as e:
        print(f, "failed", e)
PY
