import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def libtvf_path():
    """libtvf.so, (re)built with nvcc when sources are newer (cross-compiles without a GPU)."""
    from tft_vs_fund_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def hostcheck():
    """TEST-ONLY host build of the csrc headers (thread-level device math compiled with g++)."""
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostcheck.so")
    inc = os.path.join(ROOT, "tft_vs_fund_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in ("tvf_math.cuh", "tvf_pose.cuh", "tvf_scene.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-I", inc,
                               "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def tvf(libtvf_path):
    """The product package with a live handle (GPU tests only)."""
    import tft_vs_fund_b200 as pkg
    pkg.handle()          # raises loudly when there is no CUDA device / library
    return pkg


# ---- comparison helpers shared by the tests --------------------------------------------------
def rel_frob_up_to_sign(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    a = a / np.linalg.norm(a); b = b / np.linalg.norm(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))


def rot_angle(Ra, Rb):
    """angle of Ra'*Rb in radians, accurate near 0 (acos loses half the digits there)."""
    D = Ra.T @ Rb
    s = 0.5 * np.sqrt((D[2, 1] - D[1, 2]) ** 2 + (D[0, 2] - D[2, 0]) ** 2 + (D[1, 0] - D[0, 1]) ** 2)
    c = 0.5 * (np.trace(D) - 1.0)
    return float(np.arctan2(s, c))


def vec_angle(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.arctan2(np.linalg.norm(np.cross(a, b)), np.dot(a, b)))


# tolerances of BASELINE.json's north_star
TOL_MODEL = 1e-9      # T and F: relative Frobenius, up to sign and scale
TOL_ANGLE = 1e-6      # R / t angular difference, rad
TOL_REPR = 1e-8       # reprojection error, px


def assert_pose_close(ref, got, n_label=""):
    """ref/got: (R_t_2, R_t_3, Reconst, T, repr_err)."""
    R2, R3, Rec, T, rep = ref
    g2, g3, gRec, gT, grep = got
    assert rel_frob_up_to_sign(T, gT) < TOL_MODEL, n_label
    for a, b in ((R2, g2), (R3, g3)):
        assert rot_angle(a[:, :3], b[:, :3]) < TOL_ANGLE, n_label
        assert vec_angle(a[:, 3], b[:, 3]) < TOL_ANGLE, n_label
    assert abs(np.linalg.norm(R3[:, 3]) - np.linalg.norm(g3[:, 3])) <= 1e-9 * np.linalg.norm(R3[:, 3]), n_label
    assert abs(np.linalg.norm(g2[:, 3]) - 1.0) < 1e-12, n_label
    assert np.max(np.abs(Rec - gRec)) <= 1e-8 * max(1.0, np.max(np.abs(Rec))), n_label
    # 1e-8 px; the relative term only matters for failed (hundreds-of-px) solutions whose own
    # sensitivity to a 1-ulp input change already exceeds 1e-8 px ("tolerances")
    assert abs(rep - grep) <= TOL_REPR + 1e-10 * abs(rep), n_label
