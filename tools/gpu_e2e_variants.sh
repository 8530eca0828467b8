#!/bin/bash
# end-to-end (host-pointer, pinned) pipeline variants: slots and ramped chunk schedule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base nslot4 ramp nslot4ramp nslot6; do
  if [ $v = base ]; then unset TVF_LIBPATH; else export TVF_LIBPATH=tools/_build/variants/libtvf_$v.so; fi
  echo "== $v" >> gpurun_out/e2e_var_e2e.txt
  timeout 300 python tools/e2e_chunk_sweep.py 32768 65536 >> gpurun_out/e2e_var_e2e.txt 2>&1
done
cat gpurun_out/e2e_var_e2e.txt
