set -x
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_$1.json").read().strip().splitlines()[-1])
    print("variant $1 value %.4g e2e %.4g F %.4g" % (d["value"], d["e2e"]["value"], d["f_method"]["value"]), {k: round(x["ms_total"],2) for k,x in d["kernels"].items()})
except Exception as e: print("variant $1 parse fail", e)
PY
}
for v in base st64 cand64 cand256; do
  TVF_LIBPATH=$PWD/tools/_build/variants/libtvf_$v.so timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  show $v
done
for c in 59200 62160 125000 131072 250000 500000 1000000; do
  TVF_LIBPATH=$PWD/tools/_build/variants/libtvf_base.so timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $c > gpurun_out/v_chunk$c.json 2> gpurun_out/v_chunk$c.err
  show chunk$c
done
