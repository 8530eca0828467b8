#!/bin/bash
# round 2, call Q (4 GPUs, final build): the scaling bench at N = 4 (host-link ceiling with all ranks at once, shard check)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/n4_bench_n4.json 2> gpurun_out/n4_bench_n4.err
nvidia-smi topo -m > gpurun_out/n4_topo.txt 2>&1; nproc >> gpurun_out/n4_topo.txt; numactl -H >> gpurun_out/n4_topo.txt 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/n4_bench_n4.json"))
print("value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["shard_check"], d["host_link"], {k:(round(v["value"]/1e6,2), round(v["frac_of_link_ceiling"],3)) for k,v in d["e2e_variants"].items()}, d["device_resident_sweep"]["value"])
PY
tail -3 gpurun_out/n4_bench_n4.err
