"""The numerical design of the CUDA kernels (tests/kernel_model.py restates it in NumPy) against
the oracle: moments -> Gram, shifted-Cholesky inverse iteration, projected 15x15 problem,
closed-form pinv(E)*t, QR-based DLT.  Pins iteration caps and tolerances without a GPU."""
import numpy as np
import pytest

import kernel_model as km
import oracle as o
from oracle.reference_port import _tft_design_matrix
from conftest import rel_frob_up_to_sign


@pytest.mark.parametrize("n,noise", [(7, 1.0), (8, 0.0), (20, 0.0), (20, 1.0), (20, 3.0), (100, 1.0)])
def test_gram_and_constrained_solve(n, noise):
    for seed in (1, 2):
        CalM, _, C, _ = o.experiments_subsample(n, noise, seed)
        xs = [o.Normalize2Ddata(C[2 * v:2 * v + 2])[0] for v in range(3)]
        A = _tft_design_matrix(*xs)
        G = km.gram27_from_moments(km.moments96(*xs))
        assert np.abs(G - A.T @ A).max() <= 1e-14 * np.abs(G).max()
        T, P1, P2, P3, T1 = o.linearTFT(*xs, return_stage1=True)
        t, its = km.smallest_eigvec_spd(G)
        assert its < 40
        assert rel_frob_up_to_sign(t, T1.ravel(order="F")) < 1e-9
        e21, e31 = km.epipoles(t.reshape(3, 3, 3, order="F"))
        Tc, P2m, P3m, its2 = km.constrained_tft(G, e21, e31)
        assert rel_frob_up_to_sign(Tc, T) < 1e-9
        # closed-form minimum-norm a equals pinv(E)*t (linearTFT.m:86)
        E = np.hstack([np.kron(np.eye(3), np.kron(e31.reshape(3, 1), np.eye(3))), -np.kron(np.eye(9), e21.reshape(3, 1))])
        a = np.linalg.pinv(E) @ Tc.ravel(order="F")
        assert np.abs(a[:9].reshape(3, 3, order="F") - P2m[:, :3]).max() < 1e-12
        assert np.abs(a[9:].reshape(3, 3, order="F") - P3m[:, :3]).max() < 1e-12


def test_dlt_qr_inverse_iteration_matches_svd():
    CalM, R_t0, C, _ = o.generateSyntheticScene(40, 1.0, 3, 50, 0)
    K = CalM[:3]
    Ps = [K @ np.eye(3, 4), K @ R_t0[0], K @ R_t0[1]]
    X4 = o.triangulation3D(Ps, C)
    for i in range(40):
        x, its = km.dlt_null(km.dlt_rows(Ps, [C[0:2, i], C[2:4, i], C[4:6, i]]))
        assert its <= 6
        assert rel_frob_up_to_sign(x, X4[:, i]) < 1e-12
