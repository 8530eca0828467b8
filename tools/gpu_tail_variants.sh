#!/bin/bash
# Pose-tail variants (certified ray test for the cheirality votes) on ONE box: full GPU suite on the in-tree library, then the
# headline leg and the large-n bench for the in-tree library and every named variant (tools/_build/variants/libtvf_<name>.so).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-tailv}
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 8 > gpurun_out/${T}_tests.log
for v in base "$@"; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --legs headline --no-cpu-baseline > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
  timeout 300 python bench.py --workload large-n --n 10000 --trials 8192 --steps 3 --warmup 1 > gpurun_out/${T}_large_$v.json 2> gpurun_out/${T}_large_$v.err
done
unset TVF_LIBPATH
cat gpurun_out/${T}_tests.log
python - $T base "$@" <<'PY'
import json, sys
T = sys.argv[1]
for f in sys.argv[2:]:
    try:
        d = json.load(open("gpurun_out/%s_bench_%s.json" % (T, f)))
        print(f, "value %.4g" % d["value"], {k: round(v["ms_total"], 2) for k, v in d["kernels"].items()}, "flagged", d["flagged_problems"])
        d = json.load(open("gpurun_out/%s_large_%s.json" % (T, f)))
        print("   large-n: gram %.4g scenes/s" % d["value"], "full %.4g" % d["full_pipeline"]["value"],
              {k: round(v["ms_total"], 2) for k, v in d["kernels"].items()}, "flagged", d["flagged_problems"])
    except Exception as e:
        print(f, "failed", e)
PY
