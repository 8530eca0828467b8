import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
# round 2: split stage 1 (moments + solve kernels), certified vote signs, fused and un-fused tails, stage-2 dual kernel with an
# odd batch, refine path (n < 12), the four device-resident experiments, group handle, staged pageable path
d = scene.sweep_batch(77, 20, first_trial=5)
r = tvf.LinearTFTPoseEstimation(d["Corresp"], d["CalM"]); print("tft n=20 ok", int(np.count_nonzero(r.status)), r.votes[0])
r = tvf.LinearFPoseEstimation(d["Corresp"], d["CalM"]); print("f n=20 ok", int(np.count_nonzero(r.status)))
d9 = scene.sweep_batch(13, 9); r = tvf.LinearTFTPoseEstimation(d9["Corresp"], d9["CalM"]); print("refine ok", r.status)
d3 = scene.sweep_batch(5, 300 - 100 if False else 40); r = tvf.LinearTFTPoseEstimation(d3["Corresp"], d3["CalM"]); print("n=40 ok")
C300 = np.stack([scene.generateSyntheticScene(300, 1.0, s, 50, 0)[2] for s in (1, 2)])
r = tvf.LinearTFTPoseEstimation(C300, CalM); print("unfused tail ok", r.status, r.votes[0])
for opt in ("focal", "points", "angle"):
    iv, t = experiments.run_experiment(opt, n_sim=2, methods=(1, 7)); print(opt, "ok", t[1][:2, 0])
r2 = tvf.LinearTFTPoseEstimation(d["Corresp"], d["CalM"], device=(0, 0)); print("group ok", np.array_equal(r2[3], tvf.LinearTFTPoseEstimation(d["Corresp"], d["CalM"])[3]))
big = scene.sweep_batch(6000, 20, first_trial=3)
r = tvf.LinearTFTPoseEstimation(big["Corresp"], big["CalM"]); print("staged ok", int(np.count_nonzero(r.status)))
# round 2, late: TMA-staged moments (n <= 64) against the plain moments kernel (n = 65), both CTA sizes of the fused tail
# (n = 48 -> 256 threads, n = 100 -> 128 threads), odd batches
for nn, bb in ((48, 11), (64, 9), (65, 7), (100, 5)):
    dn = scene.sweep_batch(bb, nn, first_trial=11)
    r = tvf.LinearTFTPoseEstimation(dn["Corresp"], dn["CalM"]); print("n=%d ok" % nn, int(np.count_nonzero(r.status)))
