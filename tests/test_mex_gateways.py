"""The MEX gateways of mex/ built against the stub mx* API (mex/mex_stub): they compile, export
mexFunction, reproduce linearF's error text without touching the device, and -- on the GPU box --
return what the reference signatures promise (checked against the golden vectors)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_pose_close, rel_frob_up_to_sign

MEX = os.path.join(ROOT, "mex")
BUILD = os.path.join(MEX, "_build")
GATEWAYS = ["LinearTFTPoseEstimation", "LinearFPoseEstimation", "OptimFPoseEstimation", "linearTFT", "linearF", "optimF"]


@pytest.fixture(scope="module")
def built(libtvf_path):
    subprocess.check_call(["make", "-s", "-C", MEX, "stub"])
    return BUILD


def _run(built, gateway, nlhs, inputs, tmp):
    """inputs: list of column-major-ready arrays given as (array, dims tuple).  Returns list of outputs."""
    args = [os.path.join(built, "stub_driver"), os.path.join(built, gateway + ".mexstub.so"), str(nlhs),
            os.path.join(tmp, "out.bin")]
    for k, (a, dims) in enumerate(inputs):
        f = os.path.join(tmp, "in%d.bin" % k)
        np.ascontiguousarray(a, dtype=np.float64).tofile(f)
        args.append("%s:%s" % (f, "x".join(str(d) for d in dims)))
    p = subprocess.run(args, capture_output=True, text=True)
    if p.returncode == 3:
        return p.stdout.strip()
    assert p.returncode == 0, p.stdout + p.stderr
    shapes = [tuple(int(x) for x in line.split()[2:]) for line in p.stdout.splitlines() if line.startswith("OUT")]
    data = np.fromfile(os.path.join(tmp, "out.bin"))
    outs, off = [], 0
    for s in shapes:
        n = int(np.prod(s))
        outs.append(data[off:off + n].reshape(s, order="F")); off += n
    return outs


def test_gateways_build_and_export_mexfunction(built):
    for g in GATEWAYS:
        lib = ctypes.CDLL(os.path.join(built, g + ".mexstub.so"))
        assert hasattr(lib, "mexFunction")


def test_linearF_gateway_error_text_without_device(built, tmp_path):
    p = np.zeros((2, 7))
    out = _run(built, "linearF", 1, [(p.T, (2, 7)), (p.T, (2, 7))], str(tmp_path))
    assert isinstance(out, str) and out.startswith("MEXERROR TFT_vs_Fund:linearF|At least 8 correspondences are necessary")
    out = _run(built, "optimF", 2, [(p.T, (2, 7)), (p.T, (2, 7))], str(tmp_path))                   # optimF.m:36-38
    assert isinstance(out, str) and out.startswith("MEXERROR TFT_vs_Fund:linearF|At least 8 correspondences are necessary")
    out = _run(built, "LinearTFTPoseEstimation", 5, [(np.zeros((5, 20)).T, (5, 20)), (np.zeros((9, 3)).T, (9, 3))], str(tmp_path))
    assert isinstance(out, str) and "Corresp must be 6xN" in out


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["tft", "f"])
def test_pose_gateways_match_golden(built, tmp_path, method):
    g = np.load(os.path.join(ROOT, "tests", "golden", "example_n100.npz"))
    C, CalM = g["Corresp"][0], g["CalM"][0]
    gw = "LinearTFTPoseEstimation" if method == "tft" else "LinearFPoseEstimation"
    outs = _run(built, gw, 5, [(C.T, (6, 100)), (CalM.T, (9, 3))], str(tmp_path))     # .T -> column-major bytes
    assert [o.shape for o in outs] == [(3, 4), (3, 4), (3, 100), (3, 3, 3), (1, 1)] and outs[4][0, 0] == 0.0
    ref = tuple(g["%s_%s" % (method, k)][0] for k in ("Rt2", "Rt3", "Reconst", "T", "repr"))
    assert_pose_close(ref, (outs[0], outs[1], outs[2], outs[3], ref[4]))
    # one output only, as LinearTFTPoseEstimation is sometimes called
    assert len(_run(built, gw, 1, [(C.T, (6, 100)), (CalM.T, (9, 3))], str(tmp_path))) == 1
    # batched superset: 6xNxB in, 3x4xB out
    s = np.load(os.path.join(ROOT, "tests", "golden", "sweep_n20.npz"))
    Cb = s["Corresp"][:5]                                    # (B,6,n) -> bytes of 6 x n x B column-major
    outs = _run(built, gw, 5, [(Cb.transpose(0, 2, 1), (6, 20, 5)), (s["CalM"][0].T, (9, 3))], str(tmp_path))
    assert outs[0].shape == (3, 4, 5) and outs[3].shape == (3, 3, 3, 5) and outs[4].shape == (5, 1)
    for b in range(5):
        assert rel_frob_up_to_sign(outs[3][..., b], s["%s_T" % method][b]) < 1e-9


@pytest.mark.gpu
def test_estimator_gateways(built, tmp_path):
    import oracle as o
    CalM, _, C, _ = o.experiments_subsample(20, 1.0, 4)
    xs = [o.Normalize2Ddata(C[2 * v:2 * v + 2])[0] for v in range(3)]
    outs = _run(built, "linearTFT", 4, [(x.T, (2, 20)) for x in xs], str(tmp_path))
    T, P1, P2, P3 = o.linearTFT(*xs)
    assert rel_frob_up_to_sign(outs[0], T) < 1e-9 and np.array_equal(outs[1], np.eye(3, 4))
    assert outs[2].shape == (3, 4) and outs[3].shape == (3, 4)
    F = _run(built, "linearF", 1, [(C[0:2].T, (2, 20)), (C[2:4].T, (2, 20))], str(tmp_path))[0]
    assert rel_frob_up_to_sign(F, o.linearF(C[0:2], C[2:4])) < 1e-9


@pytest.mark.gpu
def test_optimf_gateways_match_golden(built, tmp_path):
    """[R_t_2,R_t_3,Reconst,T,iter]=OptimFPoseEstimation(Corresp,CalM) and [F,iter]=optimF(p1,p2) through mexFunction."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "optimf_n20.npz"))
    b = 17
    C, CalM = g["Corresp"][b], g["CalM"][b]
    outs = _run(built, "OptimFPoseEstimation", 5, [(C.T, (6, 20)), (CalM.T, (9, 3))], str(tmp_path))
    assert [o.shape for o in outs] == [(3, 4), (3, 4), (3, 20), (3, 3, 3), (1, 1)]
    assert outs[4][0, 0] == float(g["optf_iter"][b])
    ref = (g["optf_Rt2"][b], g["optf_Rt3"][b], g["optf_Reconst"][b], g["optf_T"][b], 0.0)
    assert_pose_close(ref, (outs[0], outs[1], outs[2], outs[3], 0.0))
    F, it = _run(built, "optimF", 2, [(C[0:2].T, (2, 20)), (C[2:4].T, (2, 20))], str(tmp_path))
    assert rel_frob_up_to_sign(F, g["optf_F_single"][b]) < 1e-9 and it[0, 0] == float(g["optf_iter_single"][b])
    Cb = g["Corresp"][:4]
    outs = _run(built, "OptimFPoseEstimation", 5, [(Cb.transpose(0, 2, 1), (6, 20, 4)), (CalM.T, (9, 3))], str(tmp_path))
    assert outs[4].shape == (4, 1) and np.array_equal(outs[4][:, 0], g["optf_iter"][:4].astype(np.float64))
