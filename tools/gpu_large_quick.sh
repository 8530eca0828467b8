cd "$(dirname "$0")/.."; mkdir -p gpurun_out
for v in base "$@"; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 300 python bench.py --workload large-n --n 10000 --trials 8192 --steps 3 --warmup 1 > gpurun_out/lq_large_$v.json 2> gpurun_out/lq_large_$v.err
  python - $v <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.load(open("gpurun_out/lq_large_%s.json" % f))
    print(f, "gram %.4g" % d["value"], "full %.4g" % d["full_pipeline"]["value"], {k: round(v["ms_total"], 2) for k, v in d["kernels"].items() if v["ms_total"] > 1}, "flagged", d["flagged_problems"])
except Exception as e:
    print(f, "failed", e)
PY
done
