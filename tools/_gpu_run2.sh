mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "golden or live_oracle or batch_of_one or chunking or linearTFT_direct or linearF_direct or full_size" > gpurun_out/pipe_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/pipe_tests.log | cut -c1-300
for v in base nopipe; do
  TVF_LIBPATH=$PWD/tools/_build/variants/libtvf_$v.so timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_$v.json").read().strip().splitlines()[-1])
    print("variant $v value %.4g F %.4g optF %.4g sweep %.4g" % (d["value"], d["f_method"]["value"], d["optimf_method"]["value"], d["device_resident_sweep"]["value"]), {k: round(x["ms_total"],2) for k,x in d["kernels"].items()})
except Exception as e: print("variant $v parse fail", e)
PY
done
