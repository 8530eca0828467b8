set -x
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_$1.json").read().strip().splitlines()[-1])
    print("variant $1 value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]))
except Exception as e: print("variant $1 parse fail", e)
PY
}
for v in base nslot4 nslot6 ramp_nslot4; do
  for c in 0 32768; do
  TVF_LIBPATH=$PWD/tools/_build/variants/libtvf_$v.so timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $c > gpurun_out/v_${v}_$c.json 2> gpurun_out/v_${v}_$c.err
  show ${v}_$c
  done
done
