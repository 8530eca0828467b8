"""Development tool: time / profile the device scene generator alone (1 M trials of the n = 20 sweep)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from tft_vs_fund_b200 import scene  # noqa: E402

buf = torch.empty(1000000 * 120, dtype=torch.float64, device="cuda")
for _ in range(3):
    scene.sweep_batch_device(1000000, 20, out_ptr=buf.data_ptr(), meta=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    scene.sweep_batch_device(1000000, 20, out_ptr=buf.data_ptr(), meta=False)
torch.cuda.synchronize()
print("gen ms per 1M trials (wall, incl. launch + sync)", (time.perf_counter() - t0) / 5 * 1e3)
