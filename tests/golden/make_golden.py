"""Generate the committed golden vectors under tests/golden/ (run in the build container, where
/root/reference is mounted):   python tests/golden/make_golden.py

Inputs come from the documented scene generator / the reference's EPFL data files.  Expected outputs:
  tft_*/f_*          from the oracle (oracle/reference_port.py, the NumPy restatement);
  ref_tft_*/ref_f_*  from the reference's OWN unmodified .m files under /root/reference, executed by
                     oracle/mini_matlab.py (NumPy/LAPACK built-ins) -- these pin the restatement.  The GPU box has no /root/reference, so the `-m gpu`
tests read only these .npz files (plus the live oracle on the same seeded inputs)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle as o  # noqa: E402

REFERENCE = "/root/reference"


_interp = None


def reference_run(C, CalM):
    """The reference's own .m files (unmodified, from /root/reference) executed by oracle/mini_matlab.py."""
    global _interp
    if not os.path.isdir(REFERENCE):
        return None
    from oracle.mini_matlab import reference_interpreter, Cell
    if _interp is None:
        _interp = reference_interpreter(REFERENCE, rng_factory=o.SceneRNG)
    K = [CalM[0:3], CalM[3:6], CalM[6:9]]
    out = {}
    for m, fn in (("tft", "LinearTFTPoseEstimation"), ("f", "LinearFPoseEstimation")):
        R2, R3, Rec, T, it = _interp.call(fn, [C.copy(), CalM.copy()], 5)
        assert float(np.asarray(it).item()) == 0.0
        rep = _interp.call("ReprError", [Cell([K[0] @ np.eye(3, 4), K[1] @ R2, K[2] @ R3]), C.copy(), Rec], 1)[0]
        out[m] = (R2, R3, Rec, T, float(np.asarray(rep).item()))
    return out


def run_both(C, CalM):
    K = [CalM[0:3], CalM[3:6], CalM[6:9]]
    out = {}
    ref = reference_run(C, CalM)
    if ref is not None:
        out["ref_tft"], out["ref_f"] = ref["tft"], ref["f"]
    R2, R3, Rec, T, _ = o.LinearTFTPoseEstimation(C, CalM)
    out["tft"] = (R2, R3, Rec, T, o.ReprError([K[0] @ np.eye(3, 4), K[1] @ R2, K[2] @ R3], C, Rec))
    R2, R3, Rec, T, _, F21, F31 = o.LinearFPoseEstimation(C, CalM, return_F=True)
    out["f"] = (R2, R3, Rec, T, o.ReprError([K[0] @ np.eye(3, 4), K[1] @ R2, K[2] @ R3], C, Rec), F21, F31)
    return out


def pack(cases):
    """cases: list of dict(Corresp, CalM, res=run_both(...)) with equal n -> arrays with a leading case axis."""
    d = dict(Corresp=np.stack([c["Corresp"] for c in cases]), CalM=np.stack([c["CalM"] for c in cases]))
    for m in ("tft", "f", "ref_tft", "ref_f"):
        if m not in cases[0]["res"]:
            continue
        names = ["Rt2", "Rt3", "Reconst", "T", "repr"] + (["F21", "F31"] if m == "f" else [])
        for k, name in enumerate(names):
            d["%s_%s" % (m, name)] = np.stack([np.asarray(c["res"][m][k]) for c in cases])
    for key in cases[0]:
        if key not in ("Corresp", "CalM", "res"):
            d[key] = np.stack([np.asarray(c[key]) for c in cases])
    return d


def sweep():
    """experiments.m defaults at the benchmark shape: 13 noise levels x n_sim=20 seeds, n=20."""
    cases = []
    for j in range(13 * 20):
        noise, seed = 0.25 * (j % 13), j // 13 + 1
        CalM, R_t0, C, _ = o.experiments_subsample(20, noise, seed)
        cases.append(dict(Corresp=C, CalM=CalM, res=run_both(C, CalM), noise=noise, seed=seed,
                          Rt0_2=R_t0[0], Rt0_3=R_t0[1]))
    np.savez_compressed(os.path.join(HERE, "sweep_n20.npz"), **pack(cases))


def example():
    """example.m:22-28: N=100, noise=1, seed=1, f=50, angle=0."""
    CalM, R_t0, C, _ = o.generateSyntheticScene(100, 1, 1, 50, 0)
    np.savez_compressed(os.path.join(HERE, "example_n100.npz"),
                        **pack([dict(Corresp=C, CalM=CalM, res=run_both(C, CalM), Rt0_2=R_t0[0], Rt0_3=R_t0[1])]))


def epfl(ntrip=6):
    """experiments_real.m:78-109 on the first triplets of both datasets (100 sampled inliers each)."""
    cases = []
    for ds in ("fountain-P11", "Herz-Jesu-P8"):
        path = os.path.join(REFERENCE, "Data", ds)
        idx, cor, names = o.load_corresp_triplets(path)
        for it in range(1, ntrip + 1):
            d = o.epfl_triplet(path, idx, cor, names, it)
            inl = d["Corresp_inliers"]
            sample = o.SceneRNG(it).randsample(inl.shape[1], min(100, inl.shape[1]))
            C = inl[:, sample]
            cases.append(dict(Corresp=C, CalM=d["CalM"], res=run_both(C, d["CalM"]), Rt0_2=d["R_t0"][0],
                              Rt0_3=d["R_t0"][1], n_inliers=inl.shape[1], gt_repr=d["REr"],
                              triplet=np.array(d["triplet"])))
    np.savez_compressed(os.path.join(HERE, "epfl_triplets.npz"), **pack(cases))


def optimf():
    """OptimFPoseEstimation / optimF (Gauss-Helmert refinement of F) on the first 13 x 8 trials of the
    sweep (n = 20) and on N = 12 points (experiments.m's default N).  optf_* from the oracle port
    (oracle/gauss_helmert_port.py), ref_optf_* from the reference's unmodified .m files run by mini_matlab."""
    have_ref = os.path.isdir(REFERENCE)
    if have_ref:
        from oracle.mini_matlab import reference_interpreter
        interp = reference_interpreter(REFERENCE, rng_factory=o.SceneRNG)
    for name, n, ntr in (("optimf_n20.npz", 20, 13 * 8), ("optimf_n12.npz", 12, 13 * 3)):
        d = {}
        def put(k, v):
            d.setdefault(k, []).append(np.asarray(v))
        for j in range(ntr):
            noise, seed = 0.25 * (j % 13), j // 13 + 1
            CalM, R_t0, C, _ = o.experiments_subsample(n, noise, seed)
            K = CalM[:3]
            R2, R3, Rec, T, it, F21, F31 = o.OptimFPoseEstimation(C, CalM, return_F=True)
            rep = o.ReprError([K @ np.eye(3, 4), K @ R2, K @ R3], C, Rec)
            F, it1 = o.optimF(C[0:2], C[2:4])
            for k, v in (("Corresp", C), ("CalM", CalM), ("noise", noise), ("seed", seed), ("optf_Rt2", R2), ("optf_Rt3", R3),
                         ("optf_Reconst", Rec), ("optf_T", T), ("optf_iter", it), ("optf_F21", F21), ("optf_F31", F31),
                         ("optf_repr", rep), ("optf_F_single", F), ("optf_iter_single", it1)):
                put(k, v)
            if have_ref:
                r = interp.call("OptimFPoseEstimation", [C.copy(), CalM.copy()], 5)
                rF, rit = interp.call("optimF", [C[0:2].copy(), C[2:4].copy()], 2)
                for k, v in (("ref_optf_Rt2", r[0]), ("ref_optf_Rt3", r[1]), ("ref_optf_Reconst", r[2]), ("ref_optf_T", r[3]),
                             ("ref_optf_iter", float(np.asarray(r[4]).item())), ("ref_optf_F_single", rF),
                             ("ref_optf_iter_single", float(np.asarray(rit).item()))):
                    put(k, v)
        np.savez_compressed(os.path.join(HERE, name), **{k: np.stack(v) for k, v in d.items()})


def large_n(n=10000, seeds=(1, 2)):
    """BASELINE config 5 shape: n = 10 000 correspondences per scene, 1 px noise.  The oracle needs ~90 s per
    scene, so its outputs are stored; the inputs are regenerated from the seed by the test (the first points are
    stored as a generator check) and Reconst is stored for every 50th point only."""
    d = {k: [] for k in ("seed", "Corresp_head", "CalM", "Rt2", "Rt3", "T", "repr", "Reconst_sub")}
    for seed in seeds:
        CalM, R_t0, C, _ = o.generateSyntheticScene(n, 1.0, seed, 50, 0)
        K = CalM[:3]
        R2, R3, Rec, T, _ = o.LinearTFTPoseEstimation(C, CalM)
        rep = o.ReprError([K @ np.eye(3, 4), K @ R2, K @ R3], C, Rec)
        for k, v in (("seed", seed), ("Corresp_head", C[:, :8]), ("CalM", CalM), ("Rt2", R2), ("Rt3", R3), ("T", T),
                     ("repr", rep), ("Reconst_sub", Rec[:, ::50])):
            d[k].append(np.asarray(v))
    np.savez_compressed(os.path.join(HERE, "large_n10000.npz"), **{k: np.stack(v) for k, v in d.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        large_n(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "optimf":
        optimf(); sys.exit(0)
    sweep(); example(); large_n(); optimf()
    if os.path.isdir(REFERENCE):
        epfl()
    print("golden vectors written to", HERE)
