"""Synthetic three-view scenes for the measured configurations (host side).

Product-side generator for the inputs of BASELINE.json's configs: the scene of
auxiliar_functions/generateSyntheticScene.m:45-115 and the trial sub-sampling
of experiments.m:93-95, vectorised over points and batched over trials.  It is
independent of ``oracle/`` (which restates the same script line by line); the
tests require the two to agree bit for bit.

RNG ("TVF scene RNG v1"): ``rng(seed)``/``rand`` -> MT19937 ``genrand_res53``
via ``numpy.random.RandomState(seed).random_sample`` (column-major fill, the
stream MATLAB's default generator produces); ``randn`` -> RandomState's frozen
``standard_normal``; ``randsample(n,k)`` -> ``RandomState.permutation(n)[:k]``.
MATLAB's own ``randn``/``randsample`` streams are proprietary and are not
reproduced.  Projection arithmetic is fixed and unfused:
``x_r = ((P[r,0]*X + P[r,1]*Y) + P[r,2]*Z) + P[r,3]`` followed by ``x_r/x_3``.
"""
import numpy as np

PIX = 50.0   # pixels per mm (generateSyntheticScene.m:54)


class SceneRNG:
    def __init__(self, seed):
        self.rs = np.random.RandomState(int(seed))

    def rand(self, r, c):
        return np.ascontiguousarray(self.rs.random_sample((c, r)).T)

    def randn(self, r, c):
        return np.ascontiguousarray(self.rs.standard_normal((c, r)).T)

    def randsample(self, n, k):
        return self.rs.permutation(n)[:k]


def _skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def _look_rotation(u, v):
    """Rotation taking direction u to v (generateSyntheticScene.m:119-135)."""
    u = np.asarray(u, dtype=np.float64) / np.linalg.norm(u)
    v = np.asarray(v, dtype=np.float64) / np.linalg.norm(v)
    w = np.cross(u, v)
    s = np.linalg.norm(w)
    c = np.dot(u, v)
    w = w / s
    return c * np.eye(3) + s * _skew(w) + (1 - c) * np.outer(w, w)


_camera_cache = {}


def scene_cameras(focalL=50, angle=0):
    """K, [P1,P2,P3] (spectral norm sqrt(24)), ground-truth [R_t_2, R_t_3]
    (generateSyntheticScene.m:45-72,113)."""
    key = (float(focalL), None if angle is None else float(angle))
    if key in _camera_cache:
        return _camera_cache[key]
    if angle is None or angle < 70 or angle > 180:
        p_coll = 0.0
    else:
        a = angle * np.pi / 180.0
        p_coll = 1 - np.sin(a) / (np.sqrt(2) * (np.cos(a) - 1))
    k = focalL / 50.0
    K = np.array([[50 * k * PIX, 0, 18 * PIX], [0, 50 * k * PIX, 12 * PIX], [0, 0, 1.0]])
    C = [k * np.array([0.0, -1400, 400]) + k * p_coll * np.array([0.0, 300, -300]),
         k * np.array([-400.0, -1000, 0]) + k * p_coll * np.array([0.0, -100, 100]),
         k * np.array([600.0, -800, -200]) + k * p_coll * np.array([0.0, -300, 300])]
    R = [_look_rotation(c, [0, 0, -1]) for c in C]
    Ps = []
    for Ri, Ci in zip(R, C):
        P = K @ Ri @ np.column_stack([np.eye(3), -Ci])
        Ps.append(P * np.sqrt(24) / np.linalg.norm(P, 2))
    R_t = [R[1] @ np.column_stack([R[0].T, C[0] - C[1]]), R[2] @ np.column_stack([R[0].T, C[0] - C[2]])]
    out = (K, Ps, R_t)
    _camera_cache[key] = out
    return out


def _project(P, X):
    x = [((P[r, 0] * X[0] + P[r, 1] * X[1]) + P[r, 2] * X[2]) + P[r, 3] for r in range(3)]
    return np.stack([x[0] / x[2], x[1] / x[2]])


def _inside(c6):
    hi = np.array([36 * PIX, 24 * PIX] * 3)[:, None]
    return np.all((c6 <= hi) & (c6 >= 0), axis=0)


def generateSyntheticScene(N, noise, seed, focalL=50, angle=0):
    """[CalM,R_t,Corresp,points3D]=generateSyntheticScene(N,noise,seed,focalL,angle)
    (auxiliar_functions/generateSyntheticScene.m:1,45-115)."""
    K, Ps, R_t = scene_cameras(focalL, angle)
    rng = SceneRNG(seed)
    Corresp = np.zeros((6, N)); points3D = np.zeros((3, N)); filled = 0
    M = N
    while M > 0:
        X = 400 * rng.rand(3, M) - 200
        clean = [_project(P, X) for P in Ps]
        noisy = [c + rng.randn(2, M) * noise for c in clean]
        c6 = np.vstack(noisy)
        keep = np.flatnonzero(_inside(c6))
        Corresp[:, filled:filled + keep.size] = c6[:, keep]
        points3D[:, filled:filled + keep.size] = X[:, keep]
        filled += keep.size
        M = N - filled
    return np.tile(K, (3, 1)), R_t, Corresp, points3D


def experiments_trial(n, noise, it, focalL=50, angle=0):
    """One trial of experiments.m:93-95 -> (CalM, R_t0, Corresp 6xn)."""
    CalM, R_t0, Corresp, _ = generateSyntheticScene(n + 100, noise, it, focalL, angle)
    idx = SceneRNG(it).randsample(n + 100, n)
    return CalM, R_t0, Corresp[:, idx]


def _sweep_range(args):
    """Trials [j0, j1) of the sweep (worker body; see sweep_batch)."""
    j0, j1, n, noise_levels, focalL, angle = args
    L = noise_levels.size
    K, Ps, R_t0 = scene_cameras(focalL, angle)
    M = n + 100
    out = np.empty((j1 - j0, 6, n))
    rs = np.random.RandomState(1)
    rs2 = np.random.RandomState(1)
    j = j0
    while j < j1:
        seed = j // L + 1
        lv0 = j % L
        lv1 = min(L, lv0 + (j1 - j))
        # first pass of the while-loop (generateSyntheticScene.m:80-92) is common to all noise levels
        rs.seed(seed)
        X = 400 * np.ascontiguousarray(rs.random_sample((M, 3)).T) - 200
        clean = np.vstack([_project(P, X) for P in Ps])
        Z = np.vstack([np.ascontiguousarray(rs.standard_normal((M, 2)).T) for _ in range(3)])
        rs2.seed(seed)
        idx = rs2.permutation(M)[:n]                                 # experiments.m:94-95
        for lv in range(lv0, lv1):
            noise = noise_levels[lv]
            c6 = clean + Z * noise
            ins = _inside(c6)
            o = j - j0 + (lv - lv0)
            if ins.all():
                out[o] = c6[:, idx]
                continue
            # rejections: replay the generator's stream past the first pass and keep filling
            keep = np.flatnonzero(ins)
            Corresp = np.empty((6, M))
            Corresp[:, :keep.size] = c6[:, keep]
            filled = keep.size
            rs.seed(seed)
            rs.random_sample((M, 3)); rs.standard_normal((3 * M, 2))
            while filled < M:
                m = M - filled
                Xm = 400 * np.ascontiguousarray(rs.random_sample((m, 3)).T) - 200
                cm = np.vstack([_project(P, Xm) for P in Ps])
                zm = np.vstack([np.ascontiguousarray(rs.standard_normal((m, 2)).T) for _ in range(3)])
                cm = cm + zm * noise
                kp = np.flatnonzero(_inside(cm))
                Corresp[:, filled:filled + kp.size] = cm[:, kp]
                filled += kp.size
            out[o] = Corresp[:, idx]
        j += lv1 - lv0
    return out


def sweep_batch(B, n=20, first_trial=0, noise_levels=None, focalL=50, angle=0, workers=None):
    """Config 3/4 of BASELINE.json: trial j (0-based, global index) uses noise_levels[j % L] and seed
    j // L + 1 (experiments.m:40,91-95).  Returns dict(Corresp (B,6,n), CalM (9,3), R_t0, noise (B,),
    seed (B,)).  The trial index alone determines a trial, so any rank can generate its own shard.
    `workers` > 1 forks that many generator processes (call before CUDA is initialised)."""
    if noise_levels is None:
        noise_levels = np.arange(0.0, 3.0 + 1e-9, 0.25)           # experiments.m:40
    noise_levels = np.asarray(noise_levels, dtype=np.float64)
    L = noise_levels.size
    K, Ps, R_t0 = scene_cameras(focalL, angle)
    j0, j1 = first_trial, first_trial + B
    if workers is None:
        workers = 1
    if workers <= 1 or B < 4096:
        out = _sweep_range((j0, j1, n, noise_levels, focalL, angle))
    else:
        import multiprocessing as mp
        step = max(L * 64, (B // (workers * 8) // L + 1) * L)
        jobs = [(a, min(a + step, j1), n, noise_levels, focalL, angle) for a in range(j0, j1, step)]
        with mp.get_context("fork").Pool(workers) as pool:
            out = np.concatenate(pool.map(_sweep_range, jobs), axis=0)
    jj = np.arange(j0, j1)
    return dict(Corresp=out, CalM=np.tile(K, (3, 1)), R_t0=R_t0, noise=noise_levels[jj % L], seed=jj // L + 1)


def sweep_batch_device(B, n=20, first_trial=0, noise_levels=None, focalL=50, angle=0, device=None, out_ptr=None):
    """Same trials as sweep_batch, generated by the CUDA kernel behind tvf_generate_sweep (one thread per trial).
    Returns the same dict (Corresp as a NumPy array) or, with `out_ptr` (a device pointer to 6*n*B doubles),
    fills that buffer in place and returns the dict without "Corresp".  Integer work and the projections are
    bit-exact with sweep_batch; noisy coordinates may differ in the last ulp (device log())."""
    import ctypes as C
    from . import _lib
    if noise_levels is None:
        noise_levels = np.arange(0.0, 3.0 + 1e-9, 0.25)
    noise_levels = np.ascontiguousarray(noise_levels, dtype=np.float64)
    L = noise_levels.size
    K, Ps, R_t0 = scene_cameras(focalL, angle)
    P = np.ascontiguousarray(np.stack(Ps), dtype=np.float64)            # (3,3,4) row-major
    h = _lib.handle(device)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
    jj = np.arange(first_trial, first_trial + B)
    d = dict(CalM=np.tile(K, (3, 1)), R_t0=R_t0, noise=noise_levels[jj % L], seed=jj // L + 1)
    if out_ptr is not None:
        h.call("tvf_generate_sweep_dev", first_trial, B, n, dp(noise_levels), L, dp(P), 36 * PIX, 24 * PIX, C.c_void_p(out_ptr))
        return d
    out = np.empty((B, n, 6))
    h.call("tvf_generate_sweep", first_trial, B, n, dp(noise_levels), L, dp(P), 36 * PIX, 24 * PIX, dp(out))
    d["Corresp"] = np.ascontiguousarray(out.transpose(0, 2, 1))
    return d
