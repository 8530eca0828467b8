set -x
mkdir -p gpurun_out
cat > /tmp/gen_only.py <<'PY'
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from tft_vs_fund_b200 import scene
buf = torch.empty(1000000 * 120, dtype=torch.float64, device="cuda")
for _ in range(3):
    scene.sweep_batch_device(1000000, 20, out_ptr=buf.data_ptr())
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    scene.sweep_batch_device(1000000, 20, out_ptr=buf.data_ptr())
torch.cuda.synchronize()
print("gen ms per 1M trials", (time.perf_counter() - t0) / 5 * 1e3)
PY
timeout 120 python /tmp/gen_only.py
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sweep_seeds_warp -s 2 -c 1 -f -o gpurun_out/gen_prof python /tmp/gen_only.py > gpurun_out/gen_prof.log 2>&1; echo "ncu exit $?"
