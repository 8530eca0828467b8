import numpy as np, sys
sys.path.insert(0, "/root/repo")
import tft_vs_fund_b200 as tvf
from tft_vs_fund_b200 import scene, experiments
# large-n cluster kernel (small batch), GH kernel, per-seed generator, device sweep
Cs = np.stack([scene.generateSyntheticScene(1100, 1.0, s, 50, 0)[2] for s in (1, 2, 3)])
CalM = scene.generateSyntheticScene(20, 1.0, 1, 50, 0)[0]
r = tvf.LinearTFTPoseEstimation(Cs, CalM); print("large ok", r.status)
d = scene.sweep_batch(60, 20)
r = tvf.OptimFPoseEstimation(d["Corresp"], d["CalM"]); print("optf ok", r[4][:8])
dev = scene.sweep_batch_device(100, 20, first_trial=7)
host = scene.sweep_batch(100, 20, first_trial=7)
print("gen equal", np.array_equal(dev["Corresp"], host["Corresp"]))
t = experiments.run_sweep_device(13 * 4, 20, methods=(1, 7)); print("sweep ok", t[1][:2, 0])
dev = scene.sweep_batch_device(30, 20, noise_levels=[0.0, 2.5], image=(1100.0, 800.0)); print("small image ok", dev["Corresp"].shape)
