"""Oracle restatement of the Gauss-Helmert refinement of the fundamental matrix (TEST INFRASTRUCTURE).

SURVEY.md section 8 row f4 (first step): ``Optimization/Gauss_Helmert.m``, ``F_methods/optimF.m`` and
``F_methods/OptimFPoseEstimation.m``, line by line in NumPy float64.  ``pinv`` follows MATLAB's definition
(SVD, tolerance ``max(size(A))*eps(norm(A))``).  Pinned by the reference's own unmodified ``.m`` files run
through ``oracle/mini_matlab.py`` (``ref_optf_*`` arrays of ``tests/golden/optimf_n20.npz``).
"""
import numpy as np

from .reference_port import (Normalize2Ddata, TFT_from_P, _scale_t3, crossM, linearF, matlab_svd, recover_R_t_F,
                             triangulation3D)


def matlab_pinv(A):
    """MATLAB ``pinv(A)``: SVD, singular values <= max(size(A))*eps(norm(A)) are treated as zero."""
    A = np.asarray(A, dtype=np.float64)
    if A.size == 0:
        return np.zeros((A.shape[1], A.shape[0]))
    U, s, Vh = np.linalg.svd(A, full_matrices=False)
    tol = max(A.shape) * np.spacing(s.max())
    r = int(np.sum(s > tol))
    return (Vh[:r].T / s[:r]) @ U[:, :r].T


def Gauss_Helmert(func, x0, t0, y0, x, P):
    """Optimization/Gauss_Helmert.m:38-83.  Vectors are 1-D arrays; returns (x_opt, t_opt, y_opt, iter)."""
    it_max = 400                                                    # :38
    tol = 1e-6                                                      # :39
    xi = x0.copy(); yi = y0.copy(); ti = t0.copy()                  # :41
    u = t0.shape[0]; s = y0.shape[0]                                # :42-43
    v0 = x0 - x                                                     # :45
    objFunc = float(v0 @ P @ v0)                                    # :46
    factor = 1.0                                                    # :47
    it = 0
    for it in range(1, it_max + 1):                                 # :49
        f, g, A, B, C, D = func(xi, ti, yi)                         # :50
        c2 = C.shape[0]                                             # :51
        W = B @ matlab_pinv(P) @ B.T                                # :52
        if np.any(np.isnan(W)) or np.any(np.isinf(W)):              # :53-55
            break
        W = matlab_pinv(W + 1e-12 * np.eye(W.shape[0])); W = W + 1e-12 * np.eye(W.shape[0])   # :57
        w = -f - B @ (x - xi)                                       # :58
        M = np.block([[A.T @ W @ A, np.zeros((u, s)), C.T],
                      [np.zeros((s, u + s)), D.T],
                      [C, D, np.zeros((c2, c2))]])                  # :59-61
        b = np.concatenate([A.T @ W @ w, np.zeros(s), -g])          # :62
        if np.any(np.isnan(M)) or np.any(np.isinf(M)):              # :63-65
            break
        aux = matlab_pinv(M + 1e-12 * np.eye(M.shape[0])) @ b       # :67
        dt = aux[0:u]; dy = aux[u:u + s]                            # :68
        v = -np.linalg.inv(P) @ B.T @ (W @ (A @ dt - w))            # :69
        if np.linalg.norm(dt) < tol and np.linalg.norm(dy) < tol and np.linalg.norm(xi - x - v) < tol:   # :71-73
            break
        if float(v @ P @ v) > objFunc * factor:                     # :75-76
            break
        else:
            objFunc = float(v @ P @ v)                              # :78
        xi = x + v; ti = ti + dt; yi = yi + dy                      # :80
    return xi, ti, yi, it                                           # :82-83


def constraintsGH_F(x, p, _y=None):
    """F_methods/optimF.m:81-109 (local function): f, g, A, B, C, D for the fundamental matrix."""
    N = x.shape[0] // 4                                             # :83
    x = x.reshape(4, N, order='F')                                  # :84
    F = p.reshape(3, 3, order='F')                                  # :86
    Fv = p                                                          # F(k) linear indexing, 1-based k -> Fv[k-1]
    g = np.array([np.linalg.det(F), np.sum(Fv ** 2) - 1.0])         # :88
    C = np.array([[Fv[4] * Fv[8] - Fv[5] * Fv[7], Fv[5] * Fv[6] - Fv[3] * Fv[8], Fv[3] * Fv[7] - Fv[4] * Fv[6],
                   Fv[2] * Fv[7] - Fv[1] * Fv[8], Fv[0] * Fv[8] - Fv[2] * Fv[6], Fv[1] * Fv[6] - Fv[0] * Fv[7],
                   Fv[1] * Fv[5] - Fv[2] * Fv[4], Fv[2] * Fv[3] - Fv[0] * Fv[5], Fv[0] * Fv[4] - Fv[1] * Fv[3]],
                  2.0 * Fv])                                        # :90-93
    f = np.zeros(N); A = np.zeros((N, 9)); B = np.zeros((N, 4 * N))  # :95-97
    for i in range(N):                                              # :99
        x1 = np.array([x[0, i], x[1, i], 1.0]); x2 = np.array([x[2, i], x[3, i], 1.0])   # :100
        f[i] = x2 @ F @ x1                                          # :101
        A[i, :] = [x1[0] * x2[0], x1[0] * x2[1], x1[0], x1[1] * x2[0], x1[1] * x2[1], x1[1], x2[0], x2[1], 1.0]  # :102
        B[i, 4 * i:4 * i + 4] = [Fv[2] + Fv[0] * x2[0] + Fv[1] * x2[1], Fv[5] + Fv[3] * x2[0] + Fv[4] * x2[1],
                                 Fv[6] + Fv[0] * x1[0] + Fv[3] * x1[1], Fv[7] + Fv[1] * x1[0] + Fv[4] * x1[1]]   # :103-104
    D = np.zeros((2, 0))                                            # :106
    return f, g, A, B, C, D


def optimF(p1, p2, return_detail=False):
    """F_methods/optimF.m:34-78.  Returns (F, iter)."""
    p1 = np.asarray(p1, dtype=np.float64); p2 = np.asarray(p2, dtype=np.float64)
    N = p1.shape[1]                                                 # :34
    if N != p2.shape[1] or N < 8:                                   # :36-38
        raise ValueError("At least 8 correspondences are necessary to compute the fundamental matrix linearly\\n")
    if p1.shape[0] == 3:                                            # :40-43
        p1 = p1[0:2, :] / p1[2:3, :]
        p2 = p2[0:2, :] / p2[2:3, :]
    x1, Normal1 = Normalize2Ddata(p1)                               # :46
    x2, Normal2 = Normalize2Ddata(p2)                               # :47
    F = linearF(x1, x2); F = F / np.sqrt(np.sum(F.ravel() ** 2))    # :50
    U, _, _ = matlab_svd(F); epi21 = U[:, 2]                        # :53
    P1 = np.eye(3, 4)                                               # :54
    P2 = np.column_stack([crossM(epi21) @ F, epi21])                # :55
    points3D = triangulation3D([P1, P2], np.vstack([x1, x2]))       # :56
    p1_est = P1 @ points3D; p1_est = p1_est[0:2, :] / p1_est[2:3, :]   # :59
    p2_est = P2 @ points3D; p2_est = p2_est[0:2, :] / p2_est[2:3, :]   # :60
    p = F.reshape(9, order='F')                                     # :61
    x = np.vstack([x1[0:2, :], x2[0:2, :]]).reshape(4 * N, order='F')   # :62
    x_est = np.vstack([p1_est, p2_est]).reshape(4 * N, order='F')   # :63
    y = np.zeros(0)                                                 # :64
    P = np.eye(4 * N)                                               # :65
    x_opt, p_opt, _, it = Gauss_Helmert(constraintsGH_F, x_est, p, y, x, P)   # :66
    F = p_opt.reshape(3, 3, order='F')                              # :69
    F_gh = F.copy()
    F = Normal2.T @ F @ Normal1                                     # :72
    U, s, V = matlab_svd(F); D = np.diag(s); D[2, 2] = 0.0          # :75
    F = U @ D @ V.T                                                 # :76
    if return_detail:
        return F, it, dict(F_gh=F_gh, x_opt=x_opt, x_est=x_est, F0=p.reshape(3, 3, order='F'))
    return F, it


def OptimFPoseEstimation(Corresp, CalM, return_F=False):
    """F_methods/OptimFPoseEstimation.m:43-72."""
    Corresp = np.asarray(Corresp, dtype=np.float64); CalM = np.asarray(CalM, dtype=np.float64)
    K1 = CalM[0:3, :]; K2 = CalM[3:6, :]; K3 = CalM[6:9, :]         # :44
    F21, it1 = optimF(Corresp[0:2, :], Corresp[2:4, :])             # :47
    F31, it2 = optimF(Corresp[0:2, :], Corresp[4:6, :])             # :48
    iter_ = it1 + it2                                               # :49
    R2, t2 = recover_R_t_F(K1, K2, F21, Corresp[0:2, :], Corresp[2:4, :])   # :52
    R3, t3 = recover_R_t_F(K1, K3, F31, Corresp[0:2, :], Corresp[4:6, :])   # :53
    if R2 is None or R3 is None:
        raise RuntimeError("recover_R_t left R_f undefined (all cheirality votes negative)")
    t3 = _scale_t3(K1, K2, K3, R2, t2, R3, t3, Corresp)             # :57-63
    R_t_2 = np.column_stack([R2, t2]); R_t_3 = np.column_stack([R3, t3])   # :65
    Reconst = triangulation3D([K1 @ np.eye(3, 4), K2 @ R_t_2, K3 @ R_t_3], Corresp)   # :68
    Reconst = Reconst[0:3, :] / Reconst[3:4, :]                     # :69
    T = TFT_from_P(K1 @ np.eye(3, 4), K2 @ R_t_2, K3 @ R_t_3)       # :70
    if return_F:
        return R_t_2, R_t_3, Reconst, T, iter_, F21, F31
    return R_t_2, R_t_3, Reconst, T, iter_


# --------------------------------------------------------------------------------------------------
# FaugPapaTFTPoseEstimation (TFT_methods/FaugPapaTFTPoseEstimation.m): Gauss-Helmert on the 27 tensor entries with
# the 12 Faugeras-Papadopoulo constraints
# --------------------------------------------------------------------------------------------------
def _minor(A, i, j):
    """FaugPapaTFTPoseEstimation.m:150-153 (1-based i, j)."""
    h, w = A.shape
    rows = [r for r in range(h) if r != i - 1]; cols = [c for c in range(w) if c != j - 1]
    return np.linalg.det(A[np.ix_(rows, cols)]) * (-1.0) ** (i + j)


def constrGH_FaugPapa(obs, x, _y=None):
    """FaugPapaTFTPoseEstimation.m:84-147 (local function constrGH): f, g, A, B, C, D."""
    T = x.reshape(3, 3, 3, order='F')                               # :86
    obs = obs.reshape(6, -1, order='F')                             # :87
    N = obs.shape[1]                                                # :88
    f = np.zeros(4 * N); A = np.zeros((4 * N, 27)); B = np.zeros((4 * N, 6 * N))   # :90-92
    J = np.array([[0.0, 1.0], [1.0, 0.0]])
    for i in range(N):                                              # :93
        x1 = obs[0:2, i]; x2 = obs[2:4, i]; x3 = obs[4:6, i]        # :95
        ind2 = 4 * i                                                # :98
        S2 = np.array([[0.0, -1.0], [-1.0, 0.0], [x2[1], x2[0]]])   # :99
        S3 = np.array([[0.0, -1.0], [-1.0, 0.0], [x3[1], x3[0]]])   # :100
        f[ind2:ind2 + 4] = (S2.T @ (x1[0] * T[:, :, 0] + x1[1] * T[:, :, 1] + T[:, :, 2]) @ S3).reshape(4, order='F')   # :101
        x1h = np.array([x1[0], x1[1], 1.0])
        A[ind2:ind2 + 4, :] = np.kron(S3, S2).T @ np.kron(x1h.reshape(1, 3), np.eye(9))   # :104
        B[ind2:ind2 + 4, 6 * i + 0] = (S2.T @ T[:, :, 0] @ S3).reshape(4, order='F')       # :105
        B[ind2:ind2 + 4, 6 * i + 1] = (S2.T @ T[:, :, 1] @ S3).reshape(4, order='F')       # :106
        B[ind2:ind2 + 4, 6 * i + 2:6 * i + 4] = np.kron((S3.T @ T[2, :, :].reshape(3, 3, order='F') @ x1h).reshape(2, 1), J)   # :107
        B[ind2:ind2 + 4, 6 * i + 4:6 * i + 6] = np.kron(J, (S2.T @ T[:, 2, :].reshape(3, 3, order='F') @ x1h).reshape(2, 1))   # :108
    g = np.zeros(12); C = np.zeros((12, 27)); D = np.zeros((12, 0))   # :111-113
    for i in range(1, 4):                                           # :114
        g[i - 1] = np.linalg.det(T[:, :, i - 1])                    # :115
        for j in range(1, 4):
            for k in range(1, 4):
                C[i - 1, (j + 3 * (k - 1) + 9 * (i - 1)) - 1] = _minor(T[:, :, i - 1], j, k)   # :118
    i = 0                                                           # :123
    for k2 in range(1, 3):
        for k3 in range(1, 3):
            for l2 in range(k2 + 1, 4):
                for l3 in range(k3 + 1, 4):
                    i += 1                                          # :128
                    t = lambda a, b: T[a - 1, b - 1, :]
                    A1 = np.vstack([t(k2, k3), t(k2, l3), t(l2, l3)])   # :129 reshape of a 1x3x3 array: ROW r is the r-th vector
                    A2 = np.vstack([t(k2, k3), t(l2, k3), t(l2, l3)])   # :130
                    A3 = np.vstack([t(l2, k3), t(k2, l3), t(l2, l3)])   # :131
                    A4 = np.vstack([t(k2, k3), t(l2, k3), t(k2, l3)])   # :132
                    d1, d2, d3, d4 = (np.linalg.det(M_) for M_ in (A1, A2, A3, A4))
                    g[3 + i - 1] = d1 * d2 - d3 * d4                # :133
                    for i1 in range(1, 4):                          # :134
                        col = lambda a, b: (a + 3 * (b - 1) + 9 * (i1 - 1)) - 1
                        C[3 + i - 1, col(k2, k3)] = _minor(A1, i1, 1) * d2 + d1 * _minor(A2, i1, 1) - d3 * _minor(A4, i1, 1)   # :135-136
                        C[3 + i - 1, col(k2, l3)] = _minor(A1, i1, 2) * d2 - _minor(A3, i1, 2) * d4 - d3 * _minor(A4, i1, 3)   # :137-138
                        C[3 + i - 1, col(l2, l3)] = _minor(A1, i1, 3) * d2 + d1 * _minor(A2, i1, 3) - _minor(A3, i1, 3) * d4   # :139-140
                        C[3 + i - 1, col(l2, k3)] = d1 * _minor(A2, i1, 2) - _minor(A3, i1, 1) * d4 - d3 * _minor(A4, i1, 2)   # :141-142
    return f, g, A, B, C, D


def FaugPapaTFTPoseEstimation(Corresp, CalM):
    """TFT_methods/FaugPapaTFTPoseEstimation.m:48-80."""
    from .reference_port import R_t_from_TFT, linearTFT, transform_TFT
    Corresp = np.asarray(Corresp, dtype=np.float64); CalM = np.asarray(CalM, dtype=np.float64)
    x1, Normal1 = Normalize2Ddata(Corresp[0:2, :])                  # :48-50
    x2, Normal2 = Normalize2Ddata(Corresp[2:4, :])
    x3, Normal3 = Normalize2Ddata(Corresp[4:6, :])
    T, P1, P2, P3 = linearTFT(x1, x2, x3)                           # :53
    points3D = triangulation3D([P1, P2, P3], np.vstack([x1, x2, x3]))   # :57
    est = []
    for P in (P1, P2, P3):                                          # :58-60
        p = P @ points3D
        est.append(p[0:2, :] / p[2:3, :])
    N = x1.shape[1]                                                 # :63
    param0 = T.reshape(27, order='F')                               # :64
    obs = np.vstack([x1[0:2, :], x2[0:2, :], x3[0:2, :]]).reshape(6 * N, order='F')   # :65
    obs_est = np.vstack(est).reshape(6 * N, order='F')              # :66
    y = np.zeros(0)                                                 # :67
    _, param, _, it = Gauss_Helmert(constrGH_FaugPapa, obs_est, param0, y, obs, np.eye(6 * N))   # :68
    T = param.reshape(3, 3, 3, order='F')                           # :69
    T = transform_TFT(T, Normal1, Normal2, Normal3, 1)              # :72
    R_t_2, R_t_3 = R_t_from_TFT(T, CalM, Corresp)                   # :75
    K1 = CalM[0:3, :]; K2 = CalM[3:6, :]; K3 = CalM[6:9, :]
    Reconst = triangulation3D([K1 @ np.eye(3, 4), K2 @ R_t_2, K3 @ R_t_3], Corresp)   # :78
    Reconst = Reconst[0:3, :] / Reconst[3:4, :]                     # :79
    return R_t_2, R_t_3, Reconst, T, it
