mkdir -p gpurun_out
for c in 16384 24576 32768 49152; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $c > gpurun_out/chunk_$c.json 2> gpurun_out/chunk_$c.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/chunk_$c.json").read().strip().splitlines()[-1])
print("chunk $c value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]))
PY
done
