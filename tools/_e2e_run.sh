mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "chunking or batch_of_one or sweep_first_260 or full_size" > gpurun_out/e2e_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/e2e_tests.log | cut -c1-200
for c in 0 49152; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $c > gpurun_out/e2e_$c.json 2> gpurun_out/e2e_$c.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/e2e_$c.json").read().strip().splitlines()[-1])
print("chunk $c value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), d["e2e"]["max_abs_diff_vs_device_path"])
PY
done
