set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "device_scene_generator or device_resident_sweep or experiments_sweep_table" > gpurun_out/gen_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/gen_tests.log | cut -c1-300
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/gen_bench_new.json 2> gpurun_out/gen_bench_new.err; echo "bench exit $?"
python - <<'PY'
import json
for t in ("new",):
    try:
        d = json.loads(open("gpurun_out/gen_bench_%s.json" % t).read().strip().splitlines()[-1])
        print(t, "value %.4g" % d["value"], "sweep %.4g" % d["device_resident_sweep"]["value"], d["input_generation"])
    except Exception as e:
        print(t, "parse fail", e)
PY
bash tools/_gen_prof.sh
