"""The synthetic sweep of experiments.m, batched: every (noise level, seed) trial of
experiments.m:74-144 for the two linear methods in one (sharded) call.

experiments.m loops `interval` x `n_sim` x methods, calling `methods{m}(Corresp,CalM)` once per trial
(:107-109) and accumulating ReprError / AngError means per noise level (:112-124).  Here the trial
loop collapses into one batched call per method and rank; the accumulation is a per-level sum that is
gathered and added in rank order (tft_vs_fund_b200.sharding)."""
import numpy as np

from . import _lib, api, scene, sharding

NOISE_LEVELS = np.arange(0.0, 3.0 + 1e-9, 0.25)          # experiments.m:40  interval=0:0.25:3
METHODS = {1: ("Linear TFT", api.LinearTFTPoseEstimation), 7: ("Linear F", api.LinearFPoseEstimation),
           8: ("Optimal F", api.OptimFPoseEstimation)}          # numbering of experiments.m:51-59


def evaluate(res, R_t0, device=None):
    """Per-trial errors as experiments.m:112-120: (repr_err, rot_err, t_err) with the two views averaged."""
    B = res[0].shape[0]
    r2, t2 = api.AngError(R_t0[0], res[0], device=device)
    r3, t3 = api.AngError(R_t0[1], res[1], device=device)
    return np.asarray(res.repr_err).reshape(B), (r2 + r3) / 2.0, (t2 + t3) / 2.0


def level_sums(level_idx, n_levels, *columns, bad=None):
    """Sum each per-trial column per level -> (n_levels, len(columns)+2): sums, trials counted, trials skipped.
    `bad` (bool per trial): trials without a pose (status NO_POSE / NONFINITE) -- the reference would stop there on
    the undefined R_f; both sweep drivers leave them out of the sums and report their number per level (the same
    policy as sweep_eval_accumulate_kernel)."""
    level_idx = np.asarray(level_idx)
    good = np.ones(level_idx.shape, dtype=bool) if bad is None else ~np.asarray(bad, dtype=bool)
    out = np.zeros((n_levels, len(columns) + 2))
    for k, col in enumerate(columns):
        out[:, k] = np.bincount(level_idx[good], weights=np.asarray(col)[good], minlength=n_levels)
    out[:, -2] = np.bincount(level_idx[good], minlength=n_levels)
    out[:, -1] = np.bincount(level_idx[~good], minlength=n_levels)
    return out


def run_sweep(total_trials, n=20, methods=(1, 7), noise_levels=NOISE_LEVELS, focalL=50, angle=0, device=None,
              solver=None, workers=1):
    """Sharded sweep.  Returns on rank 0 a dict method -> (n_levels, 3) array of mean [repr_err, rot_err,
    t_err] per noise level (the `(:,m,1)` slices of experiments.m:68-70), None on the other ranks.
    `solver(method_id, Corresp, CalM)` may replace the GPU call (used by the CPU plumbing tests)."""
    rank, size = sharding.world()
    lo, hi = sharding.shard_range(total_trials, rank, size)
    d = scene.sweep_batch(hi - lo, n, first_trial=lo, noise_levels=noise_levels, focalL=focalL, angle=angle,
                          workers=workers)
    L = len(noise_levels)
    level_idx = (np.arange(lo, hi) % L).astype(np.int64)
    out = {}
    skipped = {}
    for m in methods:
        bad = None
        if solver is not None:
            repr_err, rot_err, t_err = solver(m, d["Corresp"], d["CalM"], d["R_t0"])
        else:
            res = METHODS[m][1](d["Corresp"], d["CalM"], device=device)
            repr_err, rot_err, t_err = evaluate(res, d["R_t0"], device=device)
            bad = (np.asarray(res.status) & (_lib.ST_NO_POSE_2 | _lib.ST_NO_POSE_3 | _lib.ST_NONFINITE)) != 0
        total = sharding.sum_in_rank_order(level_sums(level_idx, L, repr_err, rot_err, t_err, bad=bad))
        if total is not None:
            out[m] = total[:, :3] / total[:, 3:4]
            skipped[m] = total[:, 4].astype(np.int64)
    if rank == 0:
        run_sweep.last_skipped = skipped        # trials without a pose per (method, level); all zero in the reference's sweeps
    return out if rank == 0 else None


def run_sweep_device(total_trials, n=20, methods=(1, 7), noise_levels=NOISE_LEVELS, focalL=50, angle=0, device=None):
    """run_sweep with everything on the device (tvf_sweep_run): trials are generated, solved and reduced per
    noise level in HBM; each rank moves one L x 5 table.  Per-rank sums are added in rank order on rank 0."""
    rank, size = sharding.world()
    lo, hi = sharding.shard_range(total_trials, rank, size)
    noise_levels = np.ascontiguousarray(noise_levels, dtype=np.float64)
    L = noise_levels.size
    K, Ps, R_t0 = scene.scene_cameras(focalL, angle)
    P = np.ascontiguousarray(np.stack(Ps), dtype=np.float64)
    calm = np.ascontiguousarray(np.tile(K, (3, 1)).T)
    g2 = np.ascontiguousarray(R_t0[0].T); g3 = np.ascontiguousarray(R_t0[1].T)
    h = _lib.handle(device)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
    out = {}
    skipped = {}
    for m in methods:
        table = np.zeros((L, 5))
        h.call("tvf_sweep_run", int(m), lo, hi - lo, n, dp(noise_levels), L, dp(P), 36 * scene.PIX, 24 * scene.PIX,
               dp(calm), dp(g2), dp(g3), dp(table))
        total = sharding.sum_in_rank_order(table)
        if total is not None:
            out[m] = total[:, :3] / total[:, 3:4]
            skipped[m] = total[:, 4].astype(np.int64)
    if rank == 0:
        run_sweep_device.last_skipped = skipped
    return out if rank == 0 else None


# ---- the four experiments of experiments.m:23-47 -------------------------------------------------------------
INTERVALS = {                                                                     # experiments.m:38-47
    "noise": [0.25 * k for k in range(13)],                                      # 0:0.25:3
    "focal": list(range(20, 301, 20)),                                           # 20:20:300
    "points": [7, 8, 9, 10, 15, 20, 25],                                         # [7:9,10:5:25]
    "angle": [166, 168, 170, 172, 174, 175, 176, 177, 178, 179, 179.5, 180],     # [166:2:174,175:179,179.5,180]
}


def experiment_levels(option, N=12, noise=1.0, f=50, angle=0, interval=None):
    """The per-level parameters (N, noise, f, angle) of experiments.m:74-89 for `option`; defaults are :30-33."""
    interval = INTERVALS[option] if interval is None else list(interval)
    out = []
    for v in interval:
        p = dict(N=int(N), noise=float(noise), f=f, angle=angle)
        p[{"noise": "noise", "focal": "f", "points": "N", "angle": "angle"}[option]] = int(v) if option == "points" else v
        out.append(p)
    return interval, out


def run_experiment(option, n_sim=20, methods=(1, 7), N=12, noise=1.0, f=50, angle=0, interval=None, device=None):
    """experiments.m for one `option`, device-resident (tvf_sweep_run_levels): level i x seeds 1..n_sim x methods.  Sharded
    like run_sweep (contiguous ranges of the global trial index j = (seed-1)*L + level).  Returns on rank 0
    (interval, dict method -> (L, 3) mean [repr_err, rot_err, t_err]; inf where the reference skips the method for lack
    of matches, experiments.m:99-104), None elsewhere."""
    import ctypes as C
    interval, params = experiment_levels(option, N, noise, f, angle, interval)
    L = len(params)
    levels = (_lib.SweepLevel * L)()
    for lv, p in zip(levels, params):
        K, Ps, R_t0 = scene.scene_cameras(p["f"], p["angle"])
        lv.noise, lv.n = p["noise"], p["N"]
        lv.P[:] = np.ascontiguousarray(np.stack(Ps)).ravel()
        lv.calm[:] = np.tile(K, (3, 1)).T.ravel()
        lv.Rt0_2[:] = R_t0[0].T.ravel(); lv.Rt0_3[:] = R_t0[1].T.ravel()
    rank, size = sharding.world()
    lo, hi = sharding.shard_range(L * n_sim, rank, size)
    h = _lib.handle(device)
    out, skipped = {}, {}
    for m in methods:
        table = np.zeros((L, 5))
        h.call("tvf_sweep_run_levels", int(m), lo, hi - lo, levels, L, 36 * scene.PIX, 24 * scene.PIX,
               table.ctypes.data_as(_lib.c_double_p))
        unavailable = np.isinf(table[:, 0])
        table[unavailable, :3] = 0.0
        total = sharding.sum_in_rank_order(table)
        if total is not None:
            with np.errstate(divide="ignore", invalid="ignore"):
                out[m] = total[:, :3] / total[:, 3:4]
            out[m][unavailable] = np.inf
            skipped[m] = total[:, 4].astype(np.int64)
    if rank == 0:
        run_experiment.last_skipped = skipped
    return (interval, out) if rank == 0 else None


def run_experiment_host(option, n_sim=20, methods=(1, 7), N=12, noise=1.0, f=50, angle=0, interval=None, device=None):
    """The same experiment driven from the host (inputs from scene.sweep_batch, one batched solver call per level and
    method, NumPy reduction): the twin the tests hold run_experiment against."""
    interval, params = experiment_levels(option, N, noise, f, angle, interval)
    out = {m: np.zeros((len(params), 3)) for m in methods}
    for i, p in enumerate(params):
        d = scene.sweep_batch(n_sim, p["N"], noise_levels=[p["noise"]], focalL=p["f"], angle=p["angle"])
        for m in methods:
            if (m > 6 and p["N"] < 8) or p["N"] < 7:                                                 # experiments.m:99-104
                out[m][i] = np.inf
                continue
            res = METHODS[m][1](d["Corresp"], d["CalM"], device=device)
            repr_err, rot_err, t_err = evaluate(res, d["R_t0"], device=device)
            bad = (np.asarray(res.status) & (_lib.ST_NO_POSE_2 | _lib.ST_NO_POSE_3 | _lib.ST_NONFINITE)) != 0
            sums = level_sums(np.zeros(n_sim, dtype=np.int64), 1, repr_err, rot_err, t_err, bad=bad)
            out[m][i] = sums[0, :3] / sums[0, 3]
    return interval, out
