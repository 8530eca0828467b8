"""NumPy restatement of the reference's hot-path functions (TEST INFRASTRUCTURE).

Conventions: every array is float64; shapes are the MATLAB shapes
(``Corresp`` 6xN, ``CalM`` 9x3, ``T`` 3x3x3 with ``T[:, :, i]`` the i-th
slice, poses 3x4).  ``reshape`` calls use ``order='F'`` where MATLAB's
column-major reshape matters.  Citations are file:line into /root/reference.
"""
import numpy as np


class LinearFError(ValueError):
    """Raised where F_methods/linearF.m:35-37 calls error()."""


LINEARF_ERRMSG = ("At least 8 correspondences are necessary to compute the "
                  "fundamental matrix linearly\\n")


# --------------------------------------------------------------------------
# MATLAB built-ins the reference relies on
# --------------------------------------------------------------------------
def matlab_svd(A):
    """[U,S,V]=svd(A) (full).  Returns U, s (vector), V (not V')."""
    U, s, Vh = np.linalg.svd(np.asarray(A, dtype=np.float64), full_matrices=True)
    return U, s, Vh.T


def matlab_rank(A):
    """rank(A): number of singular values > max(size(A))*eps(max(s))."""
    s = np.linalg.svd(np.asarray(A, dtype=np.float64), compute_uv=False)
    if s.size == 0:
        return 0
    tol = max(A.shape) * np.spacing(s.max())
    return int(np.sum(s > tol))


def crossM(v):
    """auxiliar_functions/crossM.m:22."""
    v = np.asarray(v, dtype=np.float64).ravel()
    return np.array([[0.0, -v[2], v[1]],
                     [v[2], 0.0, -v[0]],
                     [-v[1], v[0], 0.0]])


# --------------------------------------------------------------------------
# a1  Normalize2Ddata
# --------------------------------------------------------------------------
def Normalize2Ddata(points):
    """auxiliar_functions/Normalize2Ddata.m:33-39.  points 2xn -> (2xn, 3x3)."""
    points = np.asarray(points, dtype=np.float64)
    n = points.shape[1]                                             # :33
    points0 = np.mean(points, axis=1)                               # :34
    norm0 = np.mean(np.sqrt(np.sum((points - points0[:, None]) ** 2, axis=0)))  # :35
    N_matrix = np.diag([np.sqrt(2) / norm0, np.sqrt(2) / norm0, 1.0])  # :36
    N_matrix[0:2, 2] = -np.sqrt(2) * points0 / norm0                # :37
    new_points = N_matrix[0:2, :] @ np.vstack([points, np.ones((1, n))])  # :39
    return new_points, N_matrix


# --------------------------------------------------------------------------
# a2/a3  linearTFT
# --------------------------------------------------------------------------
def _tft_design_matrix(p1, p2, p3):
    """TFT_methods/linearTFT.m:45-62 (4 trilinearity rows per point)."""
    N = p1.shape[1]
    A = np.zeros((4 * N, 27))
    for i in range(N):
        x1, y1 = p1[0, i], p1[1, i]
        x2, y2 = p2[0, i], p2[1, i]
        x3, y3 = p3[0, i], p3[1, i]
        A[4 * i + 0, :] = [x1, 0, -x1 * x2, 0, 0, 0, -x1 * x3, 0, x1 * x2 * x3,
                           y1, 0, -x2 * y1, 0, 0, 0, -x3 * y1, 0, x2 * x3 * y1,
                           1, 0, -x2, 0, 0, 0, -x3, 0, x2 * x3]
        A[4 * i + 1, :] = [0, x1, -x1 * y2, 0, 0, 0, 0, -x1 * x3, x1 * x3 * y2,
                           0, y1, -y1 * y2, 0, 0, 0, 0, -x3 * y1, x3 * y1 * y2,
                           0, 1, -y2, 0, 0, 0, 0, -x3, x3 * y2]
        A[4 * i + 2, :] = [0, 0, 0, x1, 0, -x1 * x2, -x1 * y3, 0, x1 * x2 * y3,
                           0, 0, 0, y1, 0, -x2 * y1, -y1 * y3, 0, x2 * y1 * y3,
                           0, 0, 0, 1, 0, -x2, -y3, 0, x2 * y3]
        A[4 * i + 3, :] = [0, 0, 0, 0, x1, -x1 * y2, 0, -x1 * y3, x1 * y2 * y3,
                           0, 0, 0, 0, y1, -y1 * y2, 0, -y1 * y3, y1 * y2 * y3,
                           0, 0, 0, 0, 1, -y2, 0, -y3, y2 * y3]
    return A


def _dehomogenise(p):
    p = np.asarray(p, dtype=np.float64)
    if p.shape[0] == 3:                                             # linearTFT.m:39-43
        p = p[0:2, :] / p[2:3, :]
    return p


def linearTFT(p1, p2, p3, return_stage1=False):
    """TFT_methods/linearTFT.m:36-91.  Returns T (3x3x3), P1, P2, P3 (3x4)."""
    p1 = _dehomogenise(p1); p2 = _dehomogenise(p2); p3 = _dehomogenise(p3)
    A = _tft_design_matrix(p1, p2, p3)                              # :45-62
    _, _, V = matlab_svd(A)                                         # :64
    t = V[:, -1]                                                    # :66
    T = t.reshape(3, 3, 3, order='F')                               # :67
    T_stage1 = T.copy()

    # epipoles                                                      # :71-79
    vs = []
    for i in range(3):
        _, _, V = matlab_svd(T[:, :, i]); vs.append(V[:, -1])
    _, _, V = matlab_svd(np.column_stack(vs).T); epi31 = V[:, -1]
    vs = []
    for i in range(3):
        _, _, V = matlab_svd(T[:, :, i].T); vs.append(V[:, -1])
    _, _, V = matlab_svd(np.column_stack(vs).T); epi21 = V[:, -1]

    # constrained re-solve                                          # :82-86
    E = np.hstack([np.kron(np.eye(3), np.kron(epi31.reshape(3, 1), np.eye(3))),
                   -np.kron(np.eye(9), epi21.reshape(3, 1))])       # :82
    U, s, V = matlab_svd(E)                                         # :83
    r = matlab_rank(E)
    Up = U[:, :r]; Vp = V[:, :r]; Sp = np.diag(s[:r])
    _, _, V2 = matlab_svd(A @ Up); tp = V2[:, -1]                   # :84
    t = Up @ tp                                                     # :85
    a = Vp @ np.linalg.inv(Sp) @ tp                                 # :86

    P1 = np.eye(3, 4)                                               # :88
    P2 = np.column_stack([a[0:9].reshape(3, 3, order='F'), epi21])  # :89
    P3 = np.column_stack([a[9:18].reshape(3, 3, order='F'), epi31])  # :90
    T = t.reshape(3, 3, 3, order='F')                               # :91
    if return_stage1:
        return T, P1, P2, P3, T_stage1
    return T, P1, P2, P3


# --------------------------------------------------------------------------
# a4  transform_TFT
# --------------------------------------------------------------------------
def transform_TFT(T_old, M1, M2, M3, inverse=0):
    """TFT_methods/transform_TFT.m:32-49."""
    T_old = np.asarray(T_old, dtype=np.float64)
    T_new = np.zeros((3, 3, 3))
    if inverse == 0:                                                # :36-40
        M1i = np.linalg.inv(M1)
        for i in range(3):
            T_new[:, :, i] = M2 @ (M1i[0, i] * T_old[:, :, 0] + M1i[1, i] * T_old[:, :, 1]
                                   + M1i[2, i] * T_old[:, :, 2]) @ M3.T
    elif inverse == 1:                                              # :42-46
        M2i = np.linalg.inv(M2); M3i = np.linalg.inv(M3)
        for i in range(3):
            T_new[:, :, i] = M2i @ (M1[0, i] * T_old[:, :, 0] + M1[1, i] * T_old[:, :, 1]
                                    + M1[2, i] * T_old[:, :, 2]) @ M3i.T
    return T_new / np.linalg.norm(T_new.ravel())                    # :49


# --------------------------------------------------------------------------
# a7  triangulation3D
# --------------------------------------------------------------------------
def triangulation3D(Pcam, image_points):
    """auxiliar_functions/triangulation3D.m:32-64.  Pcam: list of M 3x4; points 2MxN or 3MxN.
    Returns 4xN (unit null vectors, arbitrary sign) or None where MATLAB returns undefined."""
    M = len(Pcam)                                                   # :32
    if M < 2:                                                       # :33-35
        return None
    image_points = np.asarray(image_points, dtype=np.float64)
    N = image_points.shape[1]                                       # :37
    if image_points.shape[0] == 2 * M:                              # :39
        pass
    elif image_points.shape[0] == 3 * M:                            # :41-45
        aux = image_points.reshape(3, N * M, order='F')
        aux = aux[0:2, :] / aux[2:3, :]
        image_points = aux.reshape(2 * M, N, order='F')
    else:                                                           # :46-47
        return None
    space_points = np.zeros((4, N))                                 # :50
    for n in range(N):                                              # :51-64
        ls_matrix = np.zeros((2 * M, 4))
        for i in range(M):
            x, y = image_points[2 * i, n], image_points[2 * i + 1, n]
            ls_matrix[2 * i:2 * i + 2, :] = np.array([[0.0, -1.0, y], [1.0, 0.0, -x]]) @ Pcam[i]  # :58-59
        _, _, V = matlab_svd(ls_matrix)                             # :61
        space_points[:, n] = V[:, 3]                                # :62-63
    return space_points


# --------------------------------------------------------------------------
# a6  recover_R_t (two local variants)
# --------------------------------------------------------------------------
def _cheirality_select(R, Rp, t, P1, K2, x1, x2, return_votes=False):
    """Shared body of R_t_from_TFT.m:91-104 and LinearFPoseEstimation.m:94-107."""
    num_points_seen = 0
    R_f = None; t_f = None
    votes = []
    for k in range(1, 5):
        if k == 2 or k == 4:
            t = -t
        elif k == 3:
            R = Rp
        X1 = triangulation3D([P1, K2 @ np.column_stack([R, t])], np.vstack([x1, x2]))
        with np.errstate(divide='ignore', invalid='ignore'):
            X1 = X1 / X1[3:4, :]
            X2 = np.column_stack([R, t]) @ X1
            vote = np.sum(np.sign(X1[2, :]) + np.sign(X2[2, :]))
        votes.append(vote)
        if vote >= num_points_seen:        # NaN compares false, as in MATLAB
            R_f = R.copy(); t_f = t.copy()
            num_points_seen = vote
    if return_votes:
        return R_f, t_f, votes
    return R_f, t_f


def _decompose_essential(E21):
    W = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    U, _, V = matlab_svd(E21)
    R = U @ W @ V.T; Rp = U @ W.T @ V.T
    R = R * np.sign(np.linalg.det(R)); Rp = Rp * np.sign(np.linalg.det(Rp))
    t = U[:, 2].copy()
    return R, Rp, t


def recover_R_t_TFT(E21, K1, K2, x1, x2, return_votes=False):
    """TFT_methods/R_t_from_TFT.m:82-106."""
    R, Rp, t = _decompose_essential(E21)                            # :84-88
    P1 = K1 @ np.eye(3, 4)                                          # :98
    return _cheirality_select(R, Rp, t, P1, K2, x1, x2, return_votes)


def recover_R_t_F(K1, K2, F21, x1, x2, return_votes=False):
    """F_methods/LinearFPoseEstimation.m:84-109."""
    E21 = K2.T @ F21 @ K1                                           # :86
    R, Rp, t = _decompose_essential(E21)                            # :87-91
    P1 = np.column_stack([K1, np.zeros(3)])                         # :101
    return _cheirality_select(R, Rp, t, P1, K2, x1, x2, return_votes)


def _scale_t3(K1, K2, K3, R2, t2, R3, t3, Corresp):
    """R_t_from_TFT.m:68-74 == LinearFPoseEstimation.m:64-70."""
    N = Corresp.shape[1]
    u3 = K3 @ t3
    X = triangulation3D([K1 @ np.eye(3, 4), K2 @ np.column_stack([R2, t2])], Corresp[0:4, :])
    X = X[0:3, :] / X[3:4, :]
    X3 = K3 @ R3 @ X
    p3 = np.vstack([Corresp[4:6, :], np.ones((1, N))])
    U3 = np.tile(u3.reshape(3, 1), (1, N))
    c1 = np.cross(p3, X3, axis=0)
    c2 = np.cross(p3, U3, axis=0)
    lam = -np.sum(np.sum(c1 * c2, axis=0)) / np.sum(np.sum(c2 ** 2))
    return lam * t3


# --------------------------------------------------------------------------
# a5  R_t_from_TFT
# --------------------------------------------------------------------------
def _epipoles_from_TFT(T, sign_fix):
    vs = []
    for i in range(3):
        _, _, V = matlab_svd(T[:, :, i]); vs.append(V[:, -1])
    _, _, V = matlab_svd(np.column_stack(vs).T)
    epi31 = V[:, -1] * (np.sign(V[-1, -1]) if sign_fix else 1.0)
    vs = []
    for i in range(3):
        _, _, V = matlab_svd(T[:, :, i].T); vs.append(V[:, -1])
    _, _, V = matlab_svd(np.column_stack(vs).T)
    epi21 = V[:, -1] * (np.sign(V[-1, -1]) if sign_fix else 1.0)
    return epi21, epi31


def R_t_from_TFT(T, CalM, Corresp, return_votes=False):
    """TFT_methods/R_t_from_TFT.m:40-76."""
    CalM = np.asarray(CalM, dtype=np.float64); Corresp = np.asarray(Corresp, dtype=np.float64)
    K1 = CalM[0:3, :]; K2 = CalM[3:6, :]; K3 = CalM[6:9, :]         # :41
    T = transform_TFT(T, K1, K2, K3, 1)                             # :44
    epi21, epi31 = _epipoles_from_TFT(T, sign_fix=True)             # :47-55
    E21 = crossM(epi21) @ np.column_stack([T[:, :, i] @ epi31 for i in range(3)])      # :57
    E31 = -crossM(epi31) @ np.column_stack([T[:, :, i].T @ epi21 for i in range(3)])   # :58
    out2 = recover_R_t_TFT(E21, K1, K2, Corresp[0:2, :], Corresp[2:4, :], return_votes)  # :61
    out3 = recover_R_t_TFT(E31, K1, K3, Corresp[0:2, :], Corresp[4:6, :], return_votes)  # :64
    R2, t2 = out2[0], out2[1]; R3, t3 = out3[0], out3[1]
    if R2 is None or R3 is None:
        raise RuntimeError("recover_R_t left R_f undefined (all cheirality votes negative)")
    t3 = _scale_t3(K1, K2, K3, R2, t2, R3, t3, Corresp)             # :68-74
    R_t_2 = np.column_stack([R2, t2]); R_t_3 = np.column_stack([R3, t3])  # :76
    if return_votes:
        return R_t_2, R_t_3, out2[2], out3[2]
    return R_t_2, R_t_3


# --------------------------------------------------------------------------
# a8  LinearTFTPoseEstimation
# --------------------------------------------------------------------------
def LinearTFTPoseEstimation(Corresp, CalM):
    """TFT_methods/LinearTFTPoseEstimation.m:45-62."""
    Corresp = np.asarray(Corresp, dtype=np.float64); CalM = np.asarray(CalM, dtype=np.float64)
    x1, Normal1 = Normalize2Ddata(Corresp[0:2, :])                  # :45
    x2, Normal2 = Normalize2Ddata(Corresp[2:4, :])                  # :46
    x3, Normal3 = Normalize2Ddata(Corresp[4:6, :])                  # :47
    T = linearTFT(x1, x2, x3)[0]                                    # :50
    T = transform_TFT(T, Normal1, Normal2, Normal3, 1)              # :53
    R_t_2, R_t_3 = R_t_from_TFT(T, CalM, Corresp)                   # :56
    Reconst = triangulation3D([CalM[0:3, :] @ np.eye(3, 4), CalM[3:6, :] @ R_t_2,
                               CalM[6:9, :] @ R_t_3], Corresp)      # :59
    Reconst = Reconst[0:3, :] / Reconst[3:4, :]                     # :60
    iter_ = 0                                                       # :62
    return R_t_2, R_t_3, Reconst, T, iter_


# --------------------------------------------------------------------------
# a9  linearF
# --------------------------------------------------------------------------
def linearF(p1, p2):
    """F_methods/linearF.m:32-62."""
    p1 = np.asarray(p1, dtype=np.float64); p2 = np.asarray(p2, dtype=np.float64)
    N = p1.shape[1]                                                 # :32
    if N != p2.shape[1] or N < 8:                                   # :35-37
        raise LinearFError(LINEARF_ERRMSG)
    p1 = _dehomogenise(p1); p2 = _dehomogenise(p2)                  # :39-42
    p1, Normal1 = Normalize2Ddata(p1[0:2, :])                       # :45
    p2, Normal2 = Normalize2Ddata(p2[0:2, :])                       # :46
    A = np.zeros((N, 9))                                            # :48
    for i in range(N):                                              # :49-53
        x1 = p1[0:2, i]; x2 = p2[0:2, i]
        A[i, :] = [x1[0] * x2[0], x1[0] * x2[1], x1[0], x1[1] * x2[0],
                   x1[1] * x2[1], x1[1], x2[0], x2[1], 1.0]
    _, _, V = matlab_svd(A)                                         # :54
    F = V[:, V.shape[1] - 1].reshape(3, 3, order='F')               # :55
    F = Normal2.T @ F @ Normal1                                     # :58
    U, D, V = matlab_svd(F); D = D.copy(); D[2] = 0.0               # :61
    F = U @ np.diag(D) @ V.T                                        # :62
    return F


# --------------------------------------------------------------------------
# a11  TFT_from_P
# --------------------------------------------------------------------------
def TFT_from_P(P1, P2, P3):
    """TFT_methods/TFT_from_P.m:25-33."""
    T = np.zeros((3, 3, 3))
    for i in range(3):
        rows = [r for r in range(3) if r != i]
        for j in range(3):
            for k in range(3):
                Mx = np.vstack([P1[rows, :], P2[j:j + 1, :], P3[k:k + 1, :]])
                T[j, k, i] = (-1.0) ** (i + 2) * np.linalg.det(Mx)   # (-1)^(i+1), i 1-based
    return T / np.linalg.norm(T.ravel())


# --------------------------------------------------------------------------
# a10  LinearFPoseEstimation
# --------------------------------------------------------------------------
def LinearFPoseEstimation(Corresp, CalM, return_F=False):
    """F_methods/LinearFPoseEstimation.m:42-78."""
    Corresp = np.asarray(Corresp, dtype=np.float64); CalM = np.asarray(CalM, dtype=np.float64)
    K1 = CalM[0:3, :]; K2 = CalM[3:6, :]; K3 = CalM[6:9, :]         # :43
    x1, Normal1 = Normalize2Ddata(Corresp[0:2, :])                  # :46
    x2, Normal2 = Normalize2Ddata(Corresp[2:4, :])                  # :47
    x3, Normal3 = Normalize2Ddata(Corresp[4:6, :])                  # :48
    F21 = linearF(x1, x2)                                           # :51
    F31 = linearF(x1, x3)                                           # :52
    F21 = Normal2.T @ F21 @ Normal1                                 # :55
    F31 = Normal3.T @ F31 @ Normal1                                 # :56
    R2, t2 = recover_R_t_F(K1, K2, F21, Corresp[0:2, :], Corresp[2:4, :])  # :59
    R3, t3 = recover_R_t_F(K1, K3, F31, Corresp[0:2, :], Corresp[4:6, :])  # :60
    if R2 is None or R3 is None:
        raise RuntimeError("recover_R_t left R_f undefined (all cheirality votes negative)")
    t3 = _scale_t3(K1, K2, K3, R2, t2, R3, t3, Corresp)             # :64-70
    R_t_2 = np.column_stack([R2, t2]); R_t_3 = np.column_stack([R3, t3])  # :72
    Reconst = triangulation3D([K1 @ np.eye(3, 4), K2 @ R_t_2, K3 @ R_t_3], Corresp)  # :75
    Reconst = Reconst[0:3, :] / Reconst[3:4, :]                     # :76
    iter_ = 0                                                       # :77
    T = TFT_from_P(K1 @ np.eye(3, 4), K2 @ R_t_2, K3 @ R_t_3)       # :78
    if return_F:
        return R_t_2, R_t_3, Reconst, T, iter_, F21, F31
    return R_t_2, R_t_3, Reconst, T, iter_


# --------------------------------------------------------------------------
# a12  ReprError, project3Dpoints
# --------------------------------------------------------------------------
def ReprError(ProjM, Corresp, Points3D=None):
    """auxiliar_functions/ReprError.m:39-65."""
    Corresp = np.asarray(Corresp, dtype=np.float64)
    N = Corresp.shape[1]; M = len(ProjM)                            # :39-40
    if Points3D is None:                                            # :43-44
        Points3D_est = triangulation3D(ProjM, Corresp)
    elif Points3D.shape[0] == 3:                                    # :45-46
        Points3D_est = np.vstack([Points3D, np.ones((1, N))])
    else:                                                           # :47-48
        Points3D_est = np.asarray(Points3D, dtype=np.float64)
    if Corresp.shape[0] == 3 * M:                                   # :52-57
        C = Corresp.reshape(3, N * M, order='F')
        C = C[0:2, :] / C[2:3, :]
    else:
        C = Corresp.reshape(2, N * M, order='F')
    P = np.vstack(ProjM)                                            # :60
    est = (P @ Points3D_est).reshape(3, M * N, order='F')           # :61
    est = est[0:2, :] / est[2:3, :]                                 # :62
    return float(np.sqrt(np.mean(np.sum((est - C) ** 2, axis=0))))  # :65


def project3Dpoints(Points3D, Pcam):
    """auxiliar_functions/project3Dpoints.m:28-35."""
    M = len(Pcam); N = Points3D.shape[1]
    Corresp = np.zeros((2 * M, N))
    for m in range(M):
        x = Pcam[m] @ np.vstack([Points3D, np.ones((1, N))])
        Corresp[2 * m:2 * m + 2, :] = x[0:2, :] / x[2:3, :]
    return Corresp


# --------------------------------------------------------------------------
# a13  AngError
# --------------------------------------------------------------------------
def _matlab_abs_acos(x):
    """abs(acos(x)) with MATLAB's complex result for |x|>1 (AngError.m:25,28)."""
    if np.isnan(x):
        return np.nan
    if x > 1.0:
        return float(np.arccosh(x))
    if x < -1.0:
        return float(np.hypot(np.pi, np.arccosh(-x)))
    return float(np.arccos(x))


def AngError(R_t_true, R_t_est):
    """auxiliar_functions/AngError.m:21-28 (degrees)."""
    R_true = R_t_true[:, 0:3]; t_true = R_t_true[:, 3]
    R_est = R_t_est[:, 0:3]; t_est = R_t_est[:, 3]
    rot_err = abs(180.0 * _matlab_abs_acos((np.trace(R_true.T @ R_est) - 1.0) / 2.0) / np.pi)
    t_err = abs(180.0 * _matlab_abs_acos(np.dot(t_est / np.linalg.norm(t_est),
                                                t_true / np.linalg.norm(t_true))) / np.pi)
    return rot_err, t_err
