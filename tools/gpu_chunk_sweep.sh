#!/bin/bash
# chunk-size sweep of the device-pointer path (problems per launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 262144 524288 1000000; do
  timeout 300 python bench.py --steps 10 --warmup 3 --legs headline --no-cpu-baseline --chunk $c > gpurun_out/chunk_bench_c$c.json 2> gpurun_out/chunk_bench_c$c.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --legs headline --no-cpu-baseline --trials 4000000 --chunk 2000000 > gpurun_out/chunk_bench_c2000000.json 2> gpurun_out/chunk_bench_c2000000.err
python - <<'PY'
import json
for f in ("262144","524288","1000000","2000000"):
    try:
        d=json.load(open("gpurun_out/chunk_bench_c%s.json"%f))
        print(f, "value %.4g"%d["value"], {k:round(v["ms_total"]/d["steps"]/ (d["config"]["trials_per_gpu"]/1e6),3) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
