mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "device_scene_generator or device_resident_sweep or experiments_sweep_table" > gpurun_out/gen_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/gen_tests.log | cut -c1-300
timeout 100 python tools/gen_only.py
timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v_memo2.json 2> gpurun_out/v_memo2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/v_memo2.json").read().strip().splitlines()[-1])
print("variant memo2 value %.4g e2e %.4g sweep %.4g gen_warm_ms %.3f" % (d["value"], d["e2e"]["value"], d["device_resident_sweep"]["value"], d["input_generation"]["seconds_warm"]*1e3), d["input_generation"]["check"])
PY
