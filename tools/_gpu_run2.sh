set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all tests exit $?"; tail -12 gpurun_out/t_all.log | cut -c1-400
timeout 240 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/v_fdual.json 2> gpurun_out/v_fdual.err
python - <<PY
import json
d=json.loads(open("gpurun_out/v_fdual.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g F %.4g optF %.4g sweep %.4g" % (d["value"], d["e2e"]["value"], d["f_method"]["value"], d["optimf_method"]["value"], d["device_resident_sweep"]["value"]))
PY
