#!/bin/bash
# round 2, call A: new parity tests + full GPU suite + baseline bench + raw-column variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2a_tests_r2.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_r2.py 2>&1 | tail -25 > gpurun_out/r2a_tests_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_base.json 2> gpurun_out/r2a_bench_base.err
TVF_LIBPATH=tools/_build/variants/libtvf_rawcol.so timeout 300 python bench.py --steps 10 --warmup 3 --legs headline --no-cpu-baseline > gpurun_out/r2a_bench_rawcol.json 2> gpurun_out/r2a_bench_rawcol.err
TVF_LIBPATH=tools/_build/variants/libtvf_rawcol.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or live_oracle or linearTFT or linearF" 2>&1 | tail -5 > gpurun_out/r2a_tests_rawcol.log
tail -3 gpurun_out/r2a_tests_r2.log gpurun_out/r2a_tests_all.log gpurun_out/r2a_tests_rawcol.log
python - <<'PY'
import json
for f in ("r2a_bench_base","r2a_bench_rawcol"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.4g"%d["value"], "e2e", d.get("e2e") and "%.4g"%d["e2e"]["value"], {k:round(v["ms_total"],2) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
