#!/bin/bash
# Large-n Gram kernel (BASELINE config 5) compile-time variants on ONE box: parity tests of the large-n route on the in-tree
# library (and on the variants named in $PARITY), a racecheck pass over a small large-n batch, then the large-n bench of
# every variant.  Usage (under gpurun): PARITY="a b" bash tools/gpu_large_n_variants.sh name1 name2 ...  where
# tools/_build/variants/libtvf_<name>.so was built by tools/build_variants.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-lgv}
if [ -z "$SKIP_BASE_CHECKS" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_n" 2>&1 | tail -n 4 > gpurun_out/${T}_tests_base.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 1 python - > gpurun_out/${T}_racecheck.log 2>&1 <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import tft_vs_fund_b200 as tvf
from tft_vs_fund_b200 import scene
base = np.stack([scene.generateSyntheticScene(1100, 1.0, s, 50, 0)[2] for s in (1, 2, 3)])
CalM = scene.generateSyntheticScene(20, 1.0, 1, 50, 0)[0]
Cs = base[np.arange(120) % 3]
r = tvf.LinearTFTPoseEstimation(Cs, CalM)
print("large ok", int(np.count_nonzero(r.status)), all(np.array_equal(r[3][k::3], np.broadcast_to(r[3][k], r[3][k::3].shape)) for k in range(3)))
PY
echo "racecheck rc $?" >> gpurun_out/${T}_racecheck.log
fi
for v in $PARITY; do
  TVF_LIBPATH=tools/_build/variants/libtvf_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_n" 2>&1 | tail -n 4 > gpurun_out/${T}_tests_$v.log
done
LG="python bench.py --workload large-n --n 10000 --trials 8192 --steps 3 --warmup 1"
for v in base "$@"; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 300 $LG > gpurun_out/${T}_large_$v.json 2> gpurun_out/${T}_large_$v.err
done
unset TVF_LIBPATH
LG1="python bench.py --workload large-n --n 10000 --trials 4096 --steps 1 --warmup 1"
for v in $NCU; do    # full ncu capture (with source) of the Gram kernel of the named variants
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tft_moments_large -s 1 -c 1 -o gpurun_out/${T}_prof_$v -f $LG1 > gpurun_out/${T}_prof_$v.log 2>&1
done
unset TVF_LIBPATH
for f in gpurun_out/${T}_tests_*.log; do echo $f; cat $f; done
[ -f gpurun_out/${T}_racecheck.log ] && tail -n 3 gpurun_out/${T}_racecheck.log
python - $T base "$@" <<'PY'
import json, sys
T = sys.argv[1]
for f in sys.argv[2:]:
    try:
        d = json.load(open("gpurun_out/%s_large_%s.json" % (T, f)))
        print(f, "gram %.4g scenes/s" % d["value"], "hbm frac %.3f" % d["roofline"]["frac"], "full %.4g" % d["full_pipeline"]["value"],
              {k: round(v["ms_total"], 2) for k, v in d["kernels"].items()}, "flagged", d["flagged_problems"])
    except Exception as e:
        print(f, "failed", e)
PY
