#!/bin/bash
# Headline leg of the bench for the in-tree library and every named compile-time variant (tools/_build/variants/libtvf_<name>.so);
# no tests (use gpu_tail_variants.sh / gpu_variants.sh when the variant changes results).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-hv}
for v in base "$@"; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --legs headline --no-cpu-baseline > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
done
unset TVF_LIBPATH
python - $T base "$@" <<'PY'
import json, sys
T = sys.argv[1]
for f in sys.argv[2:]:
    try:
        d = json.load(open("gpurun_out/%s_bench_%s.json" % (T, f)))
        print(f, "value %.4g" % d["value"], {k: round(v["ms_total"], 2) for k, v in d["kernels"].items()}, "flagged", d["flagged_problems"])
    except Exception as e:
        print(f, "failed", e)
PY
