mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "device_scene_generator or device_resident_sweep or experiments_sweep_table" > gpurun_out/gen_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/gen_tests.log | cut -c1-300
for v in base swminb12; do
  TVF_LIBPATH=$PWD/tools/_build/variants/libtvf_$v.so timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_$v.json").read().strip().splitlines()[-1])
    print("variant $v value %.4g e2e %.4g sweep %.4g gen_warm_ms %.3f" % (d["value"], d["e2e"]["value"], d["device_resident_sweep"]["value"], d["input_generation"]["seconds_warm"]*1e3))
except Exception as e: print("variant $v parse fail", e)
PY
done
