"""Oracle restatement of the synthetic scene generator (TEST INFRASTRUCTURE).

Follows auxiliar_functions/generateSyntheticScene.m:45-135 and the trial
sub-sampling of experiments.m:93-95 line by line.

RNG ("TVF scene RNG v2", documented substitute for MATLAB's generators):
  * ``rng(seed)`` + ``rand(r,c)``  -> ``numpy.random.RandomState(seed)``
    ``.random_sample``, filled column-major.  This is MT19937 ``genrand_res53``,
    the generator behind MATLAB's default ``rng(seed)``/``rand`` (widely
    documented equivalence; unverifiable here without MATLAB).
  * ``randn(r,c)`` -> the polar (Marsaglia) method on the same stream, filled
    column-major in NumPy-legacy order (second value of a pair first, the other
    cached), with the logarithm taken by ``tvf_log``: a fixed sequence of IEEE
    operations, so host and CUDA generators agree bit for bit (libm ``log``
    differs between glibc and CUDA in the last ulp).  MATLAB's ziggurat
    ``randn`` is proprietary and NOT reproduced.
  * ``randsample(n,k)`` -> first k entries of ``RandomState.permutation(n)``
    (MATLAB's Statistics-Toolbox ``randsample`` is NOT reproduced).
Parity between oracle, product generator and GPU is therefore defined on
bit-identical *arrays* produced by this documented generator.

Projection arithmetic is fixed (no BLAS, no FMA):
``x_r = ((P[r,0]*X + P[r,1]*Y) + P[r,2]*Z) + P[r,3]`` then ``x_r / x_3``.
"""
import numpy as np

from .reference_port import crossM


# ---- randn of "TVF scene RNG v2": polar method on the MT19937 stream with a reproducible logarithm ----------
_LN2_HI = 6.93147180369123816490e-01
_LN2_LO = 1.90821492927058770002e-10
_SQRT2_BITS = 0x3ff6a09e667f3bcd
_LOG_Q = [1.0 / k for k in (3.0, 5.0, 7.0, 9.0, 11.0, 13.0, 15.0, 17.0, 19.0, 21.0, 23.0)]


def tvf_log(x):
    """log(x) for positive normal doubles as a fixed sequence of IEEE operations -- the NumPy twin of
    tvf_log in tft_vs_fund_b200/csrc/tvf_scene.cuh (same operations in the same order, no FMA)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    bits = x.view(np.int64)
    e = ((bits >> 52) & 0x7ff) - 1023
    mb = (bits & 0x000fffffffffffff) | (1023 << 52)
    big = mb > _SQRT2_BITS
    mb = np.where(big, mb - (1 << 52), mb)
    e = e + big
    m = mb.view(np.float64)
    f = m + (-1.0)
    s = f / (2.0 + f)
    z = s * s
    q = np.full_like(z, _LOG_Q[10])
    for k in range(9, -1, -1):
        q = q * z + _LOG_Q[k]
    t = 2.0 * s
    lg = t + (t * z) * q
    ed = e.astype(np.float64)
    return ed * _LN2_HI + (lg + ed * _LN2_LO)


def polar_pairs(rs, m):
    """The next m accepted pairs of the polar method on RandomState `rs`, shape (m, 2) = [f*x2, f*x1] per pair
    (the order NumPy's legacy gauss returns them); consumes exactly the uniforms a sequential loop would."""
    if m <= 0:
        return np.empty((0, 2))
    st = rs.get_state()
    xs, oks = [], []
    have = 0
    while have < m:
        draw = int((m - have) * 1.4) + 16
        u = rs.random_sample(2 * draw).reshape(draw, 2)
        x = 2.0 * u + (-1.0)
        r2 = x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1]
        ok = (r2 < 1.0) & (r2 != 0.0)
        xs.append((x, r2)); oks.append(ok)
        have += int(ok.sum())
    x = np.concatenate([a for a, _ in xs]); r2 = np.concatenate([b for _, b in xs]); ok = np.concatenate(oks)
    idx = np.flatnonzero(ok)[:m]
    rs.set_state(st)
    rs.random_sample(2 * (int(idx[-1]) + 1))                      # leave the stream where the m-th acceptance left it
    x = x[idx]; r2 = r2[idx]
    f = np.sqrt((-2.0 * tvf_log(r2)) / r2)
    return np.stack([f * x[:, 1], f * x[:, 0]], axis=1)



class SceneRNG:
    """rng(seed) / rand / randn / randsample stand-ins (see module docstring)."""

    def __init__(self, seed):
        self.rs = np.random.RandomState(int(seed))
        self._cached = None                      # second value of the last pair when an odd count was drawn

    def rand(self, r, c):
        return self.rs.random_sample((c, r)).T.copy()

    def randn(self, r, c):
        count = r * c
        out = np.empty(count)
        k = 0
        if self._cached is not None and count > 0:
            out[0] = self._cached; self._cached = None; k = 1
        pairs = polar_pairs(self.rs, (count - k + 1) // 2).ravel()
        out[k:] = pairs[:count - k]
        if (count - k) % 2:
            self._cached = pairs[-1]
        return np.ascontiguousarray(out.reshape(c, r).T)

    def randsample(self, n, k):
        """0-based indices, k distinct values out of range(n)."""
        return self.rs.permutation(n)[:k]


def _rotation(u, v):
    """generateSyntheticScene.m:119-135 (local function `rotation`)."""
    u = np.asarray(u, dtype=np.float64).ravel(); v = np.asarray(v, dtype=np.float64).ravel()
    u = u / np.linalg.norm(u); v = v / np.linalg.norm(v)            # :126
    w = np.cross(u, v)                                              # :127
    s = np.linalg.norm(w)                                           # :128
    c = np.dot(u, v)                                                # :129
    w = w / s                                                       # :132
    return c * np.eye(3) + s * crossM(w) + (1 - c) * np.outer(w, w)  # :133


def scene_cameras(focalL, angle):
    """generateSyntheticScene.m:45-72: K, P1..P3 (scaled), and ground-truth R_t."""
    if angle is None or angle < 70 or angle > 180:                  # :45-50
        p_coll = 0.0
    else:
        a = angle * np.pi / 180.0
        p_coll = 1 - np.sin(a) / (np.sqrt(2) * (np.cos(a) - 1))
    k = focalL / 50.0                                               # :53
    pix = 50.0                                                      # :54
    K = np.array([[50 * k * pix, 0, 18 * pix],
                  [0, 50 * k * pix, 12 * pix],
                  [0, 0, 1.0]])                                     # :55-57
    C1 = k * np.array([0.0, -1400, 400]) + k * p_coll * np.array([0.0, 300, -300])    # :60
    C2 = k * np.array([-400.0, -1000, 0]) + k * p_coll * np.array([0.0, -100, 100])   # :61
    C3 = k * np.array([600.0, -800, -200]) + k * p_coll * np.array([0.0, -300, 300])  # :62
    R1 = _rotation(C1, [0, 0, -1]); R2 = _rotation(C2, [0, 0, -1]); R3 = _rotation(C3, [0, 0, -1])  # :65-67
    Ps = []
    for R, C in ((R1, C1), (R2, C2), (R3, C3)):                     # :70-72
        P = K @ R @ np.column_stack([np.eye(3), -C])
        P = P * np.sqrt(24) / np.linalg.norm(P, 2)
        Ps.append(P)
    R_t = [R2 @ np.column_stack([R1.T, C1 - C2]), R3 @ np.column_stack([R1.T, C1 - C3])]  # :113
    return K, Ps, R_t, pix


def _project(P, X):
    """P*[X;1] then ./ third row, in the fixed arithmetic order of the module docstring."""
    rows = []
    for r in range(3):
        rows.append(((P[r, 0] * X[0, :] + P[r, 1] * X[1, :]) + P[r, 2] * X[2, :]) + P[r, 3])
    x = np.vstack(rows)
    return x / x[2:3, :]


def generateSyntheticScene(N, noise, seed, focalL, angle, rng=None):
    """generateSyntheticScene.m:45-115.  Returns CalM (9x3), R_t (list of two 3x4),
    Corresp (6xN), points3D (3xN)."""
    K, (P1, P2, P3), R_t, pix = scene_cameras(focalL, angle)
    rng = SceneRNG(seed) if rng is None else rng                    # :75
    M = N                                                           # :76
    points3D = np.zeros((3, N)); Corresp = np.zeros((6, N)); ind1 = 0  # :77-79
    while M > 0:                                                    # :80
        X = 400 * rng.rand(3, M) - 200                              # :82
        x1 = _project(P1, X); x2 = _project(P2, X); x3 = _project(P3, X)   # :85-87
        x1n = x1[0:2, :] + rng.randn(2, M) * noise                  # :90
        x2n = x2[0:2, :] + rng.randn(2, M) * noise                  # :91
        x3n = x3[0:2, :] + rng.randn(2, M) * noise                  # :92
        inside = np.flatnonzero(                                     # :95-100
            (x1n[0] <= 36 * pix) & (x1n[1] <= 24 * pix) &
            (x2n[0] <= 36 * pix) & (x2n[1] <= 24 * pix) &
            (x3n[0] <= 36 * pix) & (x3n[1] <= 24 * pix) &
            (x1n[0] >= 0) & (x1n[1] >= 0) & (x2n[0] >= 0) & (x2n[1] >= 0) &
            (x3n[0] >= 0) & (x3n[1] >= 0))
        L = inside.size
        Corresp[:, ind1:ind1 + L] = np.vstack([x1n[:, inside], x2n[:, inside], x3n[:, inside]])  # :102-103
        points3D[:, ind1:ind1 + L] = X[:, inside]                   # :105
        ind1 += L                                                   # :107
        M = N - ind1                                                # :110
    CalM = np.tile(K, (3, 1))                                       # :115
    return CalM, R_t, Corresp, points3D


def experiments_subsample(N, noise, it, focalL=50, angle=0):
    """experiments.m:93-95: scene of N+100 points, rng(it), keep randsample(N+100,N) columns."""
    CalM, R_t0, Corresp, _ = generateSyntheticScene(N + 100, noise, it, focalL, angle)
    idx = SceneRNG(it).randsample(N + 100, N)
    return CalM, R_t0, Corresp[:, idx], idx
