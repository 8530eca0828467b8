"""Parity of the CUDA path, part 2 (the corners VERDICT round 1 found untested): the integer cheirality votes
themselves, a random 10 000-trial sample of the 1 M-trial sweep against the live oracle, every EPFL triplet of
experiments_real.m (70 + 50) including the real-data driver `epfl.run_real`, the 'focal' / 'points' / 'angle' scenes
of experiments.m:38-47 (long focal lengths, minimal point counts, collinear centres), and shards == single call on the
hardware.  Tolerances are BASELINE.json's (conftest.py)."""
import os

import numpy as np
import pytest

import oracle as o
import parity_workers as pw
from conftest import ROOT, TOL_MODEL, assert_pose_close, library_votes_as_float, rel_frob_up_to_sign, votes8_equal

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")

def votes_equal(got10, ref8):
    """got10: 10 int32 from the library (4 + 4 votes, 2 NaN masks); ref8: the oracle's 8 votes (NaN where the sum is NaN).
    True when both pairs agree exactly under an admissible relabeling of the candidates (conftest.VOTE_RELABELINGS)."""
    return votes8_equal(library_votes_as_float(got10), ref8)


def vote_tie(ref8):
    ref8 = np.asarray(ref8, dtype=np.float64)
    for p in range(2):
        v = ref8[4 * p:4 * p + 4]
        if np.any(np.isnan(v)) or np.sum(v == np.max(v)) > 1:
            return True
    return False


def _golden(name):
    return np.load(os.path.join(GOLDEN, name))


# ---------------------------------------------------------------------------------------------- (a) votes
@pytest.mark.parametrize("name", ["sweep_n20.npz", "example_n100.npz", "epfl_triplets.npz"])
@pytest.mark.parametrize("method", ["tft", "f"])
def test_cheirality_votes_equal_the_oracles(tvf, name, method):
    """R_t_from_TFT.m:91-104 / LinearFPoseEstimation.m:94-107: the vote of every candidate, integer for integer."""
    g = _golden(name)
    assert "tft_votes" in g.files, "regenerate with tests/golden/make_golden.py votes"
    fn = tvf.LinearTFTPoseEstimation if method == "tft" else tvf.LinearFPoseEstimation
    res = fn(g["Corresp"], g["CalM"])
    ref = g["%s_votes" % method]
    n = g["Corresp"].shape[2]
    for b in range(ref.shape[0]):
        assert votes_equal(res.votes[b], ref[b]), (name, method, b, res.votes[b], ref[b])
        # every candidate sees all n points in front of both cameras, behind both, or split: |vote| <= 2n
        assert np.max(np.abs(res.votes[b][:8])) <= 2 * n
    # the stand-alone R_t_from_TFT on the oracle's tensor reports the same votes as the oracle's R_t_from_TFT
    if method == "tft":
        for b in range(0, ref.shape[0], 9):
            T = g["tft_T"][b]
            _, _, v2, v3 = o.R_t_from_TFT(T, g["CalM"][b], g["Corresp"][b], return_votes=True)
            _, _, votes = tvf.R_t_from_TFT(T, g["CalM"][b], g["Corresp"][b], return_votes=True)
            assert votes_equal(votes, list(v2) + list(v3)), (name, b)


def test_votes_of_the_unfused_tail_and_nan_masks(tvf):
    """n > 256 takes votes_kernel (not the fused tail); a point at infinity (X1(4) == 0) makes the vote NaN in MATLAB
    and must be reported through the mask, not counted."""
    CalM, R_t0, C, _ = o.generateSyntheticScene(300, 1.0, 3, 50, 0)
    res = tvf.LinearTFTPoseEstimation(C[None], CalM)
    T = o.LinearTFTPoseEstimation(C, CalM)[3]
    _, _, v2, v3 = o.R_t_from_TFT(T, CalM, C, return_votes=True)
    assert votes_equal(res.votes[0], list(v2) + list(v3))
    g = _golden("sweep_n20.npz")
    a = tvf.LinearTFTPoseEstimation(g["Corresp"][:16], g["CalM"][0])
    assert np.all(a.votes[:, 8:] == 0)                      # no NaN votes on regular data


# ------------------------------------------------------------------- (b) 10 k random sample of the 1 M sweep
def test_random_10k_sample_of_the_1M_sweep_against_live_oracle(tvf):
    """BASELINE config 3: 'parity checked on the first 13 x 20 trials plus a random 10 k sample'.  The
    sample's inputs come from the oracle's own scene generator (by global trial index); the device generator's 1 M
    trials are compared with them bit for bit at those indices, then both linear methods are solved on the GPU and
    compared with the live oracle, votes included."""
    from tft_vs_fund_b200 import scene
    B, S, n = 1_000_000, 10_000, 20
    levels = np.arange(0.0, 3.0 + 1e-9, 0.25)
    idx = np.sort(np.random.RandomState(20260101).choice(B, S, replace=False))
    with pw.pool() as pool:
        ref = pool.map(pw.oracle_trial, [(int(j), n, levels, 50, 0) for j in idx], chunksize=16)
    C = np.stack([r["Corresp"] for r in ref])
    CalM = ref[0]["CalM"]
    # the device-resident generator at the same global indices (whole 1 M trials generated, sample read back)
    import torch
    d_c = torch.empty((B, n, 6), dtype=torch.float64, device="cuda")
    scene.sweep_batch_device(B, n, device=0, out_ptr=d_c.data_ptr(), meta=False)
    torch.cuda.synchronize()
    dev_sample = d_c[torch.from_numpy(idx).cuda()].cpu().numpy().transpose(0, 2, 1)
    del d_c
    assert np.array_equal(dev_sample, C), "device generator differs from the oracle's generator inside the 1 M sweep"
    worst = {"tft": 0.0, "f": 0.0}
    ties = 0
    for method, fn in (("tft", tvf.LinearTFTPoseEstimation), ("f", tvf.LinearFPoseEstimation)):
        res = fn(C, CalM)
        assert np.count_nonzero(res.status) == 0
        for b in range(S):
            r = ref[b][method]
            assert not isinstance(r, str), r
            worst[method] = max(worst[method], rel_frob_up_to_sign(r[3], res[3][b]))
            assert votes_equal(res.votes[b], r[5]), (method, int(idx[b]))
            if vote_tie(r[5]):
                ties += 1
                continue
            assert_pose_close(r[:5], (res[0][b], res[1][b], res[2][b], res[3][b], res.repr_err[b]),
                              "%s trial %d" % (method, int(idx[b])))
            if method == "f":
                assert rel_frob_up_to_sign(r[6], res.F21[b]) < TOL_MODEL and rel_frob_up_to_sign(r[7], res.F31[b]) < TOL_MODEL
    print("10k sample: worst rel. Frobenius difference of T: tft %.2e, f %.2e; vote ties skipped: %d" % (worst["tft"], worst["f"], ties))


# ---------------------------------------------------------------------------------- (c) all EPFL triplets
@pytest.mark.parametrize("method", ["tft", "f"])
def test_epfl_all_120_triplets_golden(tvf, method):
    """Config 2 in full: rows 1-70 of fountain-P11 and 1-50 of Herz-Jesu-P8 (experiments_real.m:31-36), the sampled
    inliers of every triplet: T against the oracle AND against the reference's own .m files (ref_*), poses, votes."""
    g = _golden("epfl_all_triplets.npz")
    assert g["Corresp"].shape[0] == 120
    fn = tvf.LinearTFTPoseEstimation if method == "tft" else tvf.LinearFPoseEstimation
    for ns in np.unique(g["n_sample"]):
        rows = np.flatnonzero(g["n_sample"] == ns)
        res = fn(g["Corresp"][rows][:, :, :ns], g["CalM"][rows])
        assert np.all(res.status == 0)
        for k, b in enumerate(rows):
            assert rel_frob_up_to_sign(g["%s_T" % method][b], res[3][k]) < TOL_MODEL, (method, b)
            assert rel_frob_up_to_sign(g["ref_%s_T" % method][b], res[3][k]) < TOL_MODEL, (method, b)
            assert votes_equal(res.votes[k], g["%s_votes" % method][b]), (method, b)
            if vote_tie(g["%s_votes" % method][b]):
                continue
            from conftest import rot_angle, vec_angle, TOL_ANGLE
            for a, c in ((g["%s_Rt2" % method][b], res[0][k]), (g["%s_Rt3" % method][b], res[1][k])):
                assert rot_angle(a[:, :3], c[:, :3]) < TOL_ANGLE and vec_angle(a[:, 3], c[:, 3]) < TOL_ANGLE, (method, b)


def test_run_real_executes_experiments_real_protocol(tvf):
    """epfl.run_real (experiments_real.m:75-138, methods 1 and 7) on the committed raw inputs of all 120 triplets:
    inlier counts and sample indices exact, ground-truth RMS, and the per-triplet [ReprError over all inliers, rot_err,
    t_err] rows against the oracle's and against the reference's own .m files."""
    from tft_vs_fund_b200 import epfl
    g = _golden("epfl_all_triplets.npz")
    inp = _golden("epfl_inputs.npz")
    row0 = 0
    for dsi, (key, ntrip) in enumerate((("fountain_P11", 70), ("Herz_Jesu_P8", 50))):
        data = epfl.dataset_from_arrays(inp[key + "_indexes_sorted"], inp[key + "_matches"], inp[key + "_offsets"],
                                        inp[key + "_K"], inp[key + "_R"], inp[key + "_t"])
        details = []
        table = epfl.run_real(data, range(1, ntrip + 1), details=details)
        rows = slice(row0, row0 + ntrip)
        assert np.all(g["dataset"][rows] == dsi)
        assert [d["n_inliers"] for d in details] == list(g["n_inliers"][rows])              # integer indexing: exact
        for k, d in enumerate(details):
            ns = int(g["n_sample"][row0 + k])
            assert np.array_equal(d["sample"], g["sample"][row0 + k][:ns])
            assert abs(d["REr"] - float(g["gt_repr"][row0 + k])) < 1e-8
        for m, name in ((1, "tft"), (7, "f")):
            for src in ("real_", "ref_real_"):
                ref = g[src + name][rows]
                ok = np.array([not vote_tie(v) for v in g[name + "_votes"][rows]])
                assert np.max(np.abs(table[m][ok, 0] - ref[ok, 0])) < 1e-8, (key, name, src)            # px
                assert np.max(np.abs(table[m][ok, 1:] - ref[ok, 1:])) < 1e-4, (key, name, src)          # degrees
        row0 += ntrip
    # means_all of experiments_real.m:168-174 for the linear methods is then a plain column mean
    assert np.all(np.isfinite(table[1].mean(axis=0)))


# --------------------------------------------------- (d) the other three experiments of experiments.m:38-47
EXPERIMENTS = {
    "focal": [dict(focalL=f, angle=0, n=12) for f in range(20, 301, 20)],
    "points": [dict(focalL=50, angle=0, n=n) for n in (7, 8, 9, 10, 15, 20, 25)],
    "angle": [dict(focalL=50, angle=a, n=12) for a in (166, 168, 170, 172, 174, 175, 176, 177, 178, 179, 179.5, 180)],
}


@pytest.mark.parametrize("option", ["focal", "points", "angle"])
def test_other_experiment_scenes_against_live_oracle(tvf, option):
    """experiments.m with option = 'focal' (20:20:300 mm), 'points' (7:9, 10:5:25) and 'angle' (166 ... 180 degrees,
    collinear centres), N = 12, noise = 1 px, n_sim = 20 seeds (experiments.m:30-34,38-47): the ill-conditioned
    geometries every round-1 test avoided.  Trial by trial against the live oracle for methods 1 and 7; reports how many
    trials end with TVF_ST_EIG_NOCONV."""
    from tft_vs_fund_b200 import scene, _lib
    n_sim = 20
    jobs, meta = [], []
    for lv, cfg in enumerate(EXPERIMENTS[option]):
        for it in range(1, n_sim + 1):
            jobs.append((it - 1, cfg["n"], [1.0], cfg["focalL"], cfg["angle"]))        # L = 1: seed = j + 1
            meta.append((lv, it))
    with pw.pool() as pool:
        ref = pool.map(pw.oracle_trial, jobs, chunksize=4)
    noconv = {"tft": 0, "f": 0}
    compared = {"tft": 0, "f": 0}
    worst = {"tft": 0.0, "f": 0.0}
    for lv, cfg in enumerate(EXPERIMENTS[option]):
        sel = [k for k, m in enumerate(meta) if m[0] == lv]
        C = np.stack([ref[k]["Corresp"] for k in sel]); CalM = ref[sel[0]]["CalM"]
        # the product-side generator (which feeds tvf_sweep_run's host twin) produces the same trials bit for bit
        mine = scene.sweep_batch(n_sim, cfg["n"], noise_levels=[1.0], focalL=cfg["focalL"], angle=cfg["angle"])
        assert np.array_equal(mine["Corresp"], C) and np.array_equal(mine["CalM"], CalM), (option, cfg)
        for method, fn in (("tft", tvf.LinearTFTPoseEstimation), ("f", tvf.LinearFPoseEstimation)):
            if method == "f" and cfg["n"] < 8:
                with pytest.raises(ValueError, match="At least 8 correspondences"):     # experiments.m:99-104 skips these
                    fn(C, CalM)
                continue
            res = fn(C, CalM)
            noconv[method] += int(np.count_nonzero(res.status & _lib.ST_EIG_NOCONV))
            for k, b in enumerate(sel):
                r = ref[b][method]
                if isinstance(r, str):          # the reference itself stops here (undefined R_f): we must flag it
                    assert res.status[k] & (_lib.ST_NO_POSE_2 | _lib.ST_NO_POSE_3), (option, cfg, meta[b])
                    continue
                worst[method] = max(worst[method], rel_frob_up_to_sign(r[3], res[3][k])) if method == "tft" else worst[method]
                assert votes_equal(res.votes[k], r[5]), (option, cfg, meta[b], res.votes[k], r[5])
                if vote_tie(r[5]):
                    continue
                if method == "f":
                    worst["f"] = max(worst["f"], rel_frob_up_to_sign(r[6], res.F21[k]), rel_frob_up_to_sign(r[7], res.F31[k]))
                    assert rel_frob_up_to_sign(r[6], res.F21[k]) < TOL_MODEL and rel_frob_up_to_sign(r[7], res.F31[k]) < TOL_MODEL
                assert_pose_close(r[:5], (res[0][k], res[1][k], res[2][k], res[3][k], res.repr_err[k]),
                                  "%s %s %s trial %s" % (option, cfg, method, meta[b]))
                compared[method] += 1
    print("experiments.m option=%s: compared %s trials, EIG_NOCONV %s, worst rel. model difference %s"
          % (option, compared, noconv, {k: "%.2e" % v for k, v in worst.items()}))
    assert compared["tft"] >= 0.9 * len(jobs)


def test_device_generator_bit_exact_on_collinear_and_long_focal_scenes(tvf):
    """generateSyntheticScene.m:60-67 (p_coll for angle >= 70: centres pushed onto a line) and :53-57 (focal length
    scaling K and the centres): the device generator against the host generator, bit for bit."""
    from tft_vs_fund_b200 import scene
    for focalL, angle, n in ((50, 70, 12), (50, 166, 12), (50, 179.5, 12), (50, 180, 20), (20, 0, 12), (300, 0, 12), (160, 175, 25)):
        host = scene.sweep_batch(13 * 12, n, first_trial=26, focalL=focalL, angle=angle)
        dev = scene.sweep_batch_device(13 * 12, n, first_trial=26, focalL=focalL, angle=angle)
        assert np.array_equal(host["Corresp"], dev["Corresp"]), (focalL, angle, n)
        CalM, R_t0, C, _ = o.experiments_subsample(n, 0.5, 3, focalL, angle)          # trial j = 26 + 2: level 2, seed 3
        assert np.array_equal(C, host["Corresp"][2]) and np.array_equal(CalM, host["CalM"])
        assert np.array_equal(R_t0[0], host["R_t0"][0]) and np.array_equal(R_t0[1], host["R_t0"][1])


# ------------------------------------------------------------------------------ (e) shards == single call
def _outputs(res):
    return [np.asarray(x) for x in res[:4]] + [np.asarray(res.repr_err), np.asarray(res.status), np.asarray(res.votes)] + \
           ([np.asarray(res.F21), np.asarray(res.F31)] if res.F21 is not None else [])


@pytest.mark.parametrize("method", ["tft", "f", "optf"])
def test_sharded_call_equals_single_call_bit_for_bit(tvf, method):
    """contiguous trial ranges per device, no exchange.  A group handle (tvf_create_multi) with two members on
    one GPU -- and with two and all GPUs when the box has them -- returns exactly the bits of the one-device call,
    votes and statuses included; so do two explicit half-range calls (what two torchrun ranks do)."""
    from tft_vs_fund_b200 import scene, _lib
    fn = {"tft": tvf.LinearTFTPoseEstimation, "f": tvf.LinearFPoseEstimation, "optf": tvf.OptimFPoseEstimation}[method]
    B = 13 * 701 + 5                                              # odd, does not divide by 2 or 3
    d = scene.sweep_batch(B, 20, first_trial=13 * 77)
    C, CalM = d["Corresp"], d["CalM"]
    single = _outputs(fn(C, CalM, device=0))
    ndev = _lib.load().tvf_device_count()
    groups = [(0, 0), (0, 0, 0)]
    if ndev >= 2:
        groups += [(0, 1), tuple(range(ndev))]
    for devs in groups:
        h = _lib.handle(devs)
        assert h.lib.tvf_num_devices(h._h) == len(devs)
        got = _outputs(fn(C, CalM, device=devs))
        for a, b in zip(single, got):
            assert np.array_equal(a, b, equal_nan=True), (method, devs)
    # two ranks, each solving its own contiguous range (sharding.shard_range), concatenated in rank order
    from tft_vs_fund_b200.sharding import shard_range
    parts = []
    for rank in range(2):
        lo, hi = shard_range(B, rank, 2)
        parts.append(_outputs(fn(C[lo:hi], CalM, device=(rank % ndev))))
    for a, p0, p1 in zip(single, parts[0], parts[1]):
        assert np.array_equal(a, np.concatenate([p0, p1]), equal_nan=True), method
    # per-problem CalM through the group handle
    got = _outputs(fn(C[:999], np.broadcast_to(CalM, (999, 9, 3)).copy(), device=(0, 0)))
    for a, b in zip(single, got):
        assert np.array_equal(a[:999], b, equal_nan=True)


def test_sharded_device_resident_sweep_equals_unsharded(tvf):
    """tvf_sweep_run over [0, B) == the rank-ordered sum of its two halves: per-level counts exact, sums to rounding
    (the per-trial values are bit-identical; only the order of the per-level additions differs)."""
    from tft_vs_fund_b200 import _lib, scene
    h = _lib.handle(0)
    K, Ps, R_t0 = scene.scene_cameras(50, 0)
    lv = np.ascontiguousarray(np.arange(0.0, 3.0 + 1e-9, 0.25)); P = np.ascontiguousarray(np.stack(Ps))
    calm = np.ascontiguousarray(np.tile(K, (3, 1)).T); g2 = np.ascontiguousarray(R_t0[0].T); g3 = np.ascontiguousarray(R_t0[1].T)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
    B = 13 * 4000 + 7

    def run(first, count):
        t = np.zeros((13, 5))
        h.call("tvf_sweep_run", 1, first, count, 20, dp(lv), 13, dp(P), 1800.0, 1200.0, dp(calm), dp(g2), dp(g3), dp(t))
        return t
    whole = run(0, B)
    half = B // 2 + 3
    parts = run(0, half) + run(half, B - half)
    assert np.array_equal(whole[:, 3:], parts[:, 3:])
    assert np.max(np.abs(whole[:, :3] - parts[:, :3]) / np.abs(whole[:, :3])) < 1e-12


# -------------------------------------------------------- the other experiments, device-resident (tvf_sweep_run_levels)
@pytest.mark.parametrize("option", ["noise", "focal", "points", "angle"])
def test_device_resident_experiment_equals_host_driven_and_oracle(tvf, option):
    """experiments.m:23-47 with every `option`, generated + solved + reduced on the device (tvf_sweep_run_levels: per
    level its own noise, N, cameras, CalM, ground truth) == the host-driven twin (inputs from the host generator, one
    batched call per level, NumPy reduction), and == the same loop written with the oracle on two levels."""
    from tft_vs_fund_b200 import experiments
    n_sim = 6
    interval, dev = experiments.run_experiment(option, n_sim=n_sim, methods=(1, 7, 8))
    interval_h, host = experiments.run_experiment_host(option, n_sim=n_sim, methods=(1, 7, 8))
    assert list(interval) == list(interval_h) == list(experiments.INTERVALS[option])
    for m in (1, 7, 8):
        assert dev[m].shape == (len(interval), 3)
        fin = np.isfinite(host[m][:, 0])
        assert np.array_equal(fin, np.isfinite(dev[m][:, 0]))            # the same levels are skipped (N < 8 for methods 7, 8)
        assert np.max(np.abs(dev[m][fin, 0] - host[m][fin, 0])) < 1e-7                                    # px
        assert np.max(np.abs(dev[m][fin, 1:] - host[m][fin, 1:])) < 1e-5                                  # degrees
    if option == "points":
        assert np.all(np.isinf(dev[7][0])) and np.all(np.isinf(dev[8][0])) and np.all(np.isfinite(dev[1][0]))   # N = 7
    assert all(int(v.sum()) == 0 for v in experiments.run_experiment.last_skipped.values())
    # the oracle's version of the loop (experiments.m:91-124) on the first and the last level
    _, params = experiments.experiment_levels(option)
    for lv in (0, len(params) - 1):
        p = params[lv]
        acc = {1: np.zeros(3), 7: np.zeros(3)}
        for it in range(1, n_sim + 1):
            CalM, R_t0, C, _ = o.experiments_subsample(p["N"], p["noise"], it, p["f"], p["angle"])
            K = CalM[:3]
            for m, fn in ((1, o.LinearTFTPoseEstimation), (7, o.LinearFPoseEstimation)):
                if m == 7 and p["N"] < 8:
                    continue
                R2, R3, Rec, _, _ = fn(C, CalM)
                r2, t2 = o.AngError(R_t0[0], R2); r3, t3 = o.AngError(R_t0[1], R3)
                acc[m] += np.array([o.ReprError([K @ np.eye(3, 4), K @ R2, K @ R3], C, Rec), (r2 + r3) / 2, (t2 + t3) / 2]) / n_sim
        for m in (1, 7):
            if m == 7 and p["N"] < 8:
                continue
            assert abs(dev[m][lv, 0] - acc[m][0]) < 1e-7 and np.max(np.abs(dev[m][lv, 1:] - acc[m][1:])) < 1e-4, (option, lv, m)


def test_device_resident_experiment_sharded_ranges(tvf):
    """tvf_sweep_run_levels over [0, B) == the sum of two ranges that cut through a seed's levels."""
    from tft_vs_fund_b200 import experiments, _lib, scene
    import ctypes as C
    _, params = experiments.experiment_levels("focal")
    L = len(params)
    levels = (_lib.SweepLevel * L)()
    for lv, p in zip(levels, params):
        K, Ps, R_t0 = scene.scene_cameras(p["f"], p["angle"])
        lv.noise, lv.n = p["noise"], p["N"]
        lv.P[:] = np.ascontiguousarray(np.stack(Ps)).ravel(); lv.calm[:] = np.tile(K, (3, 1)).T.ravel()
        lv.Rt0_2[:] = R_t0[0].T.ravel(); lv.Rt0_3[:] = R_t0[1].T.ravel()
    h = _lib.handle(0)

    def run(first, count):
        t = np.zeros((L, 5))
        h.call("tvf_sweep_run_levels", 1, first, count, levels, L, 1800.0, 1200.0, t.ctypes.data_as(_lib.c_double_p))
        return t
    B = L * 40 + 7
    whole = run(0, B)
    cut = L * 17 + 4
    parts = run(0, cut) + run(cut, B - cut)
    assert np.array_equal(whole[:, 3:], parts[:, 3:]) and int(whole[:, 3].sum()) == B
    assert np.max(np.abs(whole[:, :3] - parts[:, :3]) / np.abs(whole[:, :3])) < 1e-12


def test_pageable_staged_path_equals_pinned_path(tvf):
    """Pageable caller buffers (NumPy arrays, an mxArray's data) go through pinned per-slot staging buffers filled and
    drained by host threads; pinned buffers (tvf_host_alloc) are used in place.  Same bits either way, for any thread
    count, also through a group handle."""
    import ctypes as C
    from tft_vs_fund_b200 import scene, _lib
    B, n = 150_001, 20                                        # > 4 MB of input: staged; not a multiple of the chunk
    d = scene.sweep_batch(B, n, first_trial=7, workers=min(16, os.cpu_count() or 1))
    ref = _outputs(tvf.LinearTFTPoseEstimation(d["Corresp"], d["CalM"]))
    h = _lib.handle(0)
    lib = h.lib
    c = np.ascontiguousarray(d["Corresp"].transpose(0, 2, 1))
    calm = np.ascontiguousarray(d["CalM"].T)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)

    def pinned(shape, dtype=np.float64):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = lib.tvf_host_alloc(nbytes)
        assert p
        return np.frombuffer((C.c_char * nbytes).from_address(p), dtype=dtype).reshape(shape), p
    bufs = [pinned((B, n, 6)), pinned((B, 12)), pinned((B, 12)), pinned((B, 3 * n)), pinned((B, 27)), pinned((B,)), pinned((B,), np.int32)]
    try:
        bufs[0][0][...] = c
        h.call("tvf_linear_tft_pose", dp(bufs[0][0]), dp(calm), 0, n, B, dp(bufs[1][0]), dp(bufs[2][0]), dp(bufs[3][0]), dp(bufs[4][0]),
               dp(bufs[5][0]), bufs[6][0].ctypes.data_as(_lib.c_int32_p))
        assert np.array_equal(bufs[5][0], ref[4]) and np.array_equal(bufs[6][0], ref[5])
        assert np.array_equal(bufs[4][0].reshape(B, 3, 3, 3).transpose(0, 3, 2, 1), ref[3])
        for threads in (1, 3):
            h.call("tvf_set_host_threads", threads)
            got = _outputs(tvf.LinearTFTPoseEstimation(d["Corresp"], d["CalM"]))
            for a, b in zip(ref, got):
                assert np.array_equal(a, b, equal_nan=True), threads
    finally:
        h.call("tvf_set_host_threads", 0)
        for _, p in bufs:
            lib.tvf_host_free(C.c_void_p(p))
    got = _outputs(tvf.LinearFPoseEstimation(d["Corresp"][:120_000], d["CalM"], device=(0, 0)))
    one = _outputs(tvf.LinearFPoseEstimation(d["Corresp"][:120_000], d["CalM"], device=0))
    for a, b in zip(one, got):
        assert np.array_equal(a, b, equal_nan=True)
    # staged path with a per-problem CalM (9 x 3 x B travels through the staging buffer too)
    Bc = 70_003
    got = _outputs(tvf.LinearTFTPoseEstimation(d["Corresp"][:Bc], np.broadcast_to(d["CalM"], (Bc, 9, 3)).copy()))
    for a, b in zip(ref, got):
        assert np.array_equal(a[:Bc], b, equal_nan=True)


def test_group_handle_edge_cases(tvf):
    """A batch smaller than the group (B < number of members), B = 1 through a group handle, B = 0, outputs not requested."""
    import ctypes as C
    from tft_vs_fund_b200 import scene, _lib
    d = scene.sweep_batch(5, 20, first_trial=2)
    one = tvf.LinearTFTPoseEstimation(d["Corresp"], d["CalM"])
    grp = tvf.LinearTFTPoseEstimation(d["Corresp"][:2], d["CalM"], device=(0, 0, 0))
    assert np.array_equal(grp[3], one[3][:2]) and np.array_equal(grp.votes, one.votes[:2])
    single = tvf.LinearTFTPoseEstimation(d["Corresp"][3], d["CalM"], device=(0, 0))
    assert np.array_equal(single[3], one[3][3]) and single.repr_err == one.repr_err[3]
    h = _lib.handle((0, 0))
    out = _lib.PoseOut()                                     # every member NULL: nothing is copied back, the call still runs
    c = np.ascontiguousarray(d["Corresp"].transpose(0, 2, 1)); calm = np.ascontiguousarray(d["CalM"].T)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
    assert h.call("tvf_pose", 1, dp(c), dp(calm), 0, 20, 5, C.byref(out)) == 0
    assert h.call("tvf_pose", 7, dp(c), dp(calm), 0, 20, 0, C.byref(out)) == 0        # B = 0
    with pytest.raises(_lib.TvfError, match="method must be"):
        h.call("tvf_pose", 3, dp(c), dp(calm), 0, 20, 5, C.byref(out))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [33, 48, 64, 65, 100, 128, 129, 256, 257])
def test_point_counts_across_kernel_selection_boundaries_against_live_oracle(tvf, n):
    """The estimator picks its kernels by n: bulk-copy-staged moments up to n = 64 and the plain moments kernel above,
    the fused tail on 128-thread CTAs (n = 33, 64, 100, 128) or 256-thread CTAs (n = 48, 129, 256) and the three un-fused
    tail kernels from n = 257.  Seven trials each (an odd batch: partial problem groups in every kernel), noise 1 px,
    trial by trial against the live oracle for both linear methods, votes integer for integer."""
    from tft_vs_fund_b200 import scene
    B = 7
    jobs = [(j, n, [1.0], 50, 0) for j in range(B)]
    with pw.pool(min(B, 8)) as pool:
        ref = pool.map(pw.oracle_trial, jobs)
    C = np.stack([r["Corresp"] for r in ref]); CalM = ref[0]["CalM"]
    mine = scene.sweep_batch(B, n, noise_levels=[1.0])
    assert np.array_equal(mine["Corresp"], C)
    for method, fn in (("tft", tvf.LinearTFTPoseEstimation), ("f", tvf.LinearFPoseEstimation)):
        res = fn(C, CalM)
        assert not np.any(res.status), (n, method, res.status)
        for k in range(B):
            r = ref[k][method]
            assert votes_equal(res.votes[k], r[5]), (n, method, k, res.votes[k], r[5])
            assert_pose_close(r[:5], (res[0][k], res[1][k], res[2][k], res[3][k], res.repr_err[k]), "n=%d %s trial %d" % (n, method, k))


@pytest.mark.gpu
def test_device_pointer_entry_points_alignment_and_interior_offsets(tvf):
    """`*_dev` entry points: a correspondence pointer at a problem boundary INSIDE a buffer (16-byte aligned, not 256) gives
    the same results as the batch it is part of; a pointer that is only 8-byte aligned is refused with TVF_ERR_ARG before any
    launch (the kernels read 128-bit words and issue bulk copies)."""
    import ctypes as C
    import torch
    from tft_vs_fund_b200 import scene, _lib
    B, n = 37, 21                                   # 48 n = 1008 bytes per problem: boundaries are 16- but not 32-byte aligned
    d = scene.sweep_batch(B, n, noise_levels=[1.0])
    calm = np.ascontiguousarray(d["CalM"].T)
    dev = torch.device("cuda", 0)
    corresp = torch.from_numpy(np.ascontiguousarray(d["Corresp"].transpose(0, 2, 1))).to(dev)       # (B, n, 6): 6 x n x B column-major
    h = _lib.Handle(0)
    ptr = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
    d_calm = torch.from_numpy(calm).to(dev)
    def run(first):
        k = B - first
        T = torch.empty((k, 27), dtype=torch.float64, device=dev); rep = torch.empty(k, dtype=torch.float64, device=dev)
        st = torch.zeros(k, dtype=torch.int32, device=dev)
        r2 = torch.empty((k, 12), dtype=torch.float64, device=dev); r3 = torch.empty((k, 12), dtype=torch.float64, device=dev)
        rc = h.lib.tvf_linear_tft_pose_dev(h._h, ptr(corresp, first * n * 48), ptr(d_calm), 0, n, k, ptr(r2), ptr(r3), None, ptr(T), ptr(rep), ptr(st))
        assert rc == 0, h.error()
        h.call("tvf_synchronize")
        return T.cpu().numpy(), rep.cpu().numpy(), st.cpu().numpy()
    T0, rep0, st0 = run(0)
    T5, rep5, st5 = run(5)
    assert (corresp.data_ptr() + 5 * n * 48) % 32 != 0
    assert np.array_equal(T0[5:], T5) and np.array_equal(rep0[5:], rep5) and not st0.any()
    res = tvf.LinearTFTPoseEstimation(d["Corresp"], d["CalM"])
    assert np.array_equal(res.repr_err, rep0)
    k = B - 1
    out = torch.empty((k, 27), dtype=torch.float64, device=dev)
    rc = h.lib.tvf_linear_tft_pose_dev(h._h, ptr(corresp, 8), ptr(d_calm), 0, n, k, None, None, None, ptr(out), None, None)
    assert rc == _lib.TVF_ERR_ARG and "16-byte aligned" in h.error()
    h.close()
