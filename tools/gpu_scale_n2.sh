#!/bin/bash
# round 2, call P (final build) (2 GPUs): multi-device C ABI == single device on real hardware; 2-rank bench with the shard check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "sharded or group" 2>&1 | tail -6 > gpurun_out/n2_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench_n2.json 2> gpurun_out/n2_bench_n2.err
cat gpurun_out/n2_gpus.txt gpurun_out/n2_tests_2gpu.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/n2_bench_n2.json"))
print("value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["shard_check"], d["host_link"], {k:(round(v["value"]/1e6,2), round(v["frac_of_link_ceiling"],3)) for k,v in d["e2e_variants"].items()}, d["device_resident_sweep"]["value"])
PY
tail -3 gpurun_out/n2_bench_n2.err
