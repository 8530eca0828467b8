"""Synthetic three-view scenes for the measured configurations (host side).

Product-side generator for the inputs of BASELINE.json's configs: the scene of
auxiliar_functions/generateSyntheticScene.m:45-115 and the trial sub-sampling
of experiments.m:93-95, vectorised over points and batched over trials.  It is
independent of ``oracle/`` (which restates the same script line by line); the
tests require the two to agree bit for bit.

RNG ("TVF scene RNG v2"): ``rng(seed)``/``rand`` -> MT19937 ``genrand_res53``
via ``numpy.random.RandomState(seed).random_sample`` (column-major fill, the
stream MATLAB's default generator produces); ``randn`` -> the polar method on
the same stream in NumPy-legacy order, with the logarithm taken by ``tvf_log``
(a fixed sequence of IEEE operations, so the CUDA generator reproduces every
noisy coordinate bit for bit; libm's ``log`` differs between glibc and CUDA in
the last ulp); ``randsample(n,k)`` -> ``RandomState.permutation(n)[:k]``.
MATLAB's own ``randn``/``randsample`` streams are proprietary and are not
reproduced.  Projection arithmetic is fixed and unfused:
``x_r = ((P[r,0]*X + P[r,1]*Y) + P[r,2]*Z) + P[r,3]`` followed by ``x_r/x_3``.
"""
import numpy as np

PIX = 50.0   # pixels per mm (generateSyntheticScene.m:54)


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
    def __init__(self, seed):
        self.rs = np.random.RandomState(int(seed))
        self._cached = None                      # second value of the last pair when an odd count was drawn

    def rand(self, r, c):
        return np.ascontiguousarray(self.rs.random_sample((c, r)).T)

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
        Z = np.vstack([np.ascontiguousarray(polar_pairs(rs, M).T) for _ in range(3)])
        after_first_pass = rs.get_state()                            # where every level's refill passes continue
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
            rs.set_state(after_first_pass)
            while filled < M:
                m = M - filled
                Xm = 400 * np.ascontiguousarray(rs.random_sample((m, 3)).T) - 200
                cm = np.vstack([_project(P, Xm) for P in Ps])
                zm = np.vstack([np.ascontiguousarray(polar_pairs(rs, m).T) for _ in range(3)])
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


def sweep_batch_device(B, n=20, first_trial=0, noise_levels=None, focalL=50, angle=0, device=None, out_ptr=None,
                       image=None, meta=True):
    """Same trials as sweep_batch, generated by the CUDA kernel behind tvf_generate_sweep (one warp per seed).
    Returns the same dict (Corresp as a NumPy array) or, with `out_ptr` (a device pointer to 6*n*B doubles),
    fills that buffer in place and returns the dict without "Corresp".  Bit-exact with sweep_batch.
    `image` = (width, height) of the inside-image test (default: the reference's 36 x 24 mm sensor in pixels);
    `meta=False` skips the per-trial noise / seed arrays (B-sized host work that can exceed the kernel time)."""
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
    d = dict(CalM=np.tile(K, (3, 1)), R_t0=R_t0)
    if meta:
        jj = np.arange(first_trial, first_trial + B)
        d.update(noise=noise_levels[jj % L], seed=jj // L + 1)
    hi_x, hi_y = (36 * PIX, 24 * PIX) if image is None else (float(image[0]), float(image[1]))
    if out_ptr is not None:
        h.call("tvf_generate_sweep_dev", first_trial, B, n, dp(noise_levels), L, dp(P), hi_x, hi_y, C.c_void_p(out_ptr))
        return d
    out = np.empty((B, n, 6))
    h.call("tvf_generate_sweep", first_trial, B, n, dp(noise_levels), L, dp(P), hi_x, hi_y, dp(out))
    d["Corresp"] = np.ascontiguousarray(out.transpose(0, 2, 1))
    return d
