set -x
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_$1.json").read().strip().splitlines()[-1])
    print("variant $1 value %.4g" % (d["value"]), {k: round(x["ms_total"],2) for k,x in d["kernels"].items()}, "flagged", d["flagged_problems"])
except Exception as e: print("variant $1 parse fail", e)
PY
}
for v in base cheir2 cheir2_t1; do
  TVF_LIBPATH=$PWD/tools/_build/variants/libtvf_$v.so timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --legs headline > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  show $v
done
