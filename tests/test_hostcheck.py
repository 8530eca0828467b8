"""The thread-level device math of csrc/tvf_math.cuh + tvf_pose.cuh, compiled for the host
(TEST-ONLY build, tests/hostcheck) and compared with the oracle.  Same source the kernels inline."""
import ctypes as C

import numpy as np
import pytest

import oracle as o
from conftest import rel_frob_up_to_sign, assert_pose_close

dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
cm = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).ravel()   # 2-D column-major


def test_null3_and_svd3(hostcheck):
    rs = np.random.RandomState(0)
    for trial in range(50):
        M = rs.standard_normal((3, 3))
        if trial % 3 == 0:                                   # exactly rank 2
            M = np.outer(rs.standard_normal(3), rs.standard_normal(3)) + np.outer(rs.standard_normal(3), rs.standard_normal(3))
        v = np.zeros(3); hostcheck.hc_null3(dp(cm(M)), dp(v))
        vr = o.matlab_svd(M)[2][:, -1]
        assert rel_frob_up_to_sign(v, vr) < 1e-11
        U = np.zeros(9); s = np.zeros(3); V = np.zeros(9)
        if trial % 3 == 0:
            hostcheck.hc_svd3(dp(cm(M)), dp(U), dp(s), dp(V))
            U = U.reshape(3, 3).T; V = V.reshape(3, 3).T
            assert np.allclose(U @ np.diag(s) @ V.T, M, atol=1e-13)
            assert np.allclose(U.T @ U, np.eye(3), atol=1e-12) and abs(np.linalg.det(U) * np.linalg.det(V) - 1) < 2.1


def test_null3_qr_route_and_its_fallback(hostcheck):
    """null3's fast route (Householder QR + inverse iteration, tvf_math.cuh::null3_qr) against numpy's SVD on the matrices
    it meets -- nearly rank-2 slices at every gap up to s3/s2 = 0.3, exactly rank-2 matrices, both orientations -- and its
    refusals: two (nearly) equal smallest singular values, rank-1, zero, NaN / Inf input must be DECLINED (return 0) so that
    null3 takes the Jacobi route, and null3 itself must then still equal the SVD wherever that is defined."""
    rng = np.random.RandomState(5)
    answered = 0
    for trial in range(400):
        U, _ = np.linalg.qr(rng.randn(3, 3)); V, _ = np.linalg.qr(rng.randn(3, 3))
        gap = [0.0, 1e-12, 1e-6, 1e-3, 0.05, 0.3][trial % 6]
        s = np.array([1.0 + rng.rand(), 0.2 + rng.rand(), 0.0]); s[2] = gap * s[1]
        M = (U * s) @ V.T * 10.0 ** rng.randint(-3, 4)
        for tr in (0, 1):
            v = np.zeros(3)
            ok = hostcheck.hc_null3_qr(dp(cm(M)), tr, dp(v))
            ref = np.linalg.svd(M.T if tr else M)[2][2]
            if gap >= 0.3 and not ok:          # rate (s3/s2)^2 = 0.09: may need more than NULL3_MAX_IT steps -> declined, fine
                continue
            assert ok == 1, (trial, gap)
            answered += 1
            assert min(np.linalg.norm(v - ref), np.linalg.norm(v + ref)) < 1e-12 / max(1e-3, 1.0 - gap), (trial, gap, v, ref)
    # refusals
    v = np.zeros(3)
    U, _ = np.linalg.qr(rng.randn(3, 3)); V, _ = np.linalg.qr(rng.randn(3, 3))
    for s in ([1.0, 0.5, 0.5], [1.0, 0.5, 0.4999]):
        M = (U * np.array(s)) @ V.T
        ok = hostcheck.hc_null3_qr(dp(cm(M)), 0, dp(v))
        if ok:      # an answer is allowed only if it is right (a spectral gap the iteration could resolve in 12 steps)
            ref = np.linalg.svd(M)[2][2]
            assert s[1] != s[2] and min(np.linalg.norm(v - ref), np.linalg.norm(v + ref)) < 1e-9, s
    # rank 1: the null space is two-dimensional and V(:,3) is whatever the SVD picks in it (the reference's LAPACK too); an answer
    # must be a unit vector of that null space
    M = (U * np.array([1.0, 1e-20, 0.0])) @ V.T
    if hostcheck.hc_null3_qr(dp(cm(M)), 0, dp(v)):
        assert abs(np.linalg.norm(v) - 1.0) < 1e-12 and np.linalg.norm(M @ v) < 1e-15
    for bad in (np.zeros((3, 3)), np.full((3, 3), np.nan), np.array([[1.0, 2, 3], [4, np.inf, 6], [7, 8, 9]])):
        assert hostcheck.hc_null3_qr(dp(cm(bad)), 0, dp(v)) == 0
    # null3 (fast route + fallback) on a matrix the fast route declines: well separated s3 but s2 == s3 is not; use equal pair
    M = (U * np.array([1.0, 0.5, 0.49])) @ V.T
    hostcheck.hc_null3(dp(cm(M)), dp(v))
    ref = np.linalg.svd(M)[2][2]
    assert min(np.linalg.norm(v - ref), np.linalg.norm(v + ref)) < 1e-10
    assert answered >= 660            # every matrix with s3/s2 <= 0.05 was answered


def test_transform_tft_and_tft_from_p(hostcheck):
    rs = np.random.RandomState(1)
    for inverse in (0, 1):
        T = rs.standard_normal((3, 3, 3))
        Ms = [rs.standard_normal((3, 3)) + 2 * np.eye(3) for _ in range(3)]
        out = np.zeros(27)
        hostcheck.hc_transform_tft(dp(T.ravel(order="F").copy()), dp(cm(Ms[0])), dp(cm(Ms[1])), dp(cm(Ms[2])), inverse, dp(out))
        ref = o.transform_TFT(T, *Ms, inverse)
        assert np.abs(out.reshape(3, 3, 3, order="F") - ref).max() < 1e-13
    Ps = [rs.standard_normal((3, 4)) for _ in range(3)]
    out = np.zeros(27)
    hostcheck.hc_tft_from_p(dp(cm(Ps[0])), dp(cm(Ps[1])), dp(cm(Ps[2])), dp(out))
    assert np.abs(out.reshape(3, 3, 3, order="F") - o.TFT_from_P(*Ps)).max() < 1e-13


def test_onb_and_angerror(hostcheck):
    rs = np.random.RandomState(2)
    for _ in range(20):
        e = rs.standard_normal(3); e /= np.linalg.norm(e)
        u1 = np.zeros(3); u2 = np.zeros(3)
        hostcheck.hc_onb3(dp(e), dp(u1), dp(u2))
        Q = np.column_stack([e, u1, u2])
        assert np.abs(Q.T @ Q - np.eye(3)).max() < 1e-14
    A = np.column_stack([np.eye(3), [1.0, 2, 3]])
    c, s = np.cos(0.7), np.sin(0.7)
    B = np.column_stack([np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), [0.5, -1, 2.0]])
    r = np.zeros(1); t = np.zeros(1)
    hostcheck.hc_ang_error(dp(cm(A)), dp(cm(B)), dp(r), dp(t))
    rr, tt = o.AngError(A, B)
    assert abs(r[0] - rr) < 1e-11 and abs(t[0] - tt) < 1e-11


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,noise", [(8, 1.0), (20, 0.0), (20, 1.0), (20, 3.0), (60, 2.0)])
def test_pose_tail_matches_oracle(hostcheck, mode, n, noise):
    for seed in (1, 2, 3):
        CalM, R_t0, Cr, _ = o.experiments_subsample(n, noise, seed)
        K = CalM[:3]
        if mode == 0:
            R2, R3, Rec, T, _ = o.LinearTFTPoseEstimation(Cr, CalM)
            model = T.ravel(order="F").copy()
        else:
            R2, R3, Rec, T, _, F21, F31 = o.LinearFPoseEstimation(Cr, CalM, return_F=True)
            model = np.concatenate([cm(F21), cm(F31)])
        rep = o.ReprError([K @ np.eye(3, 4), K @ R2, K @ R3], Cr, Rec)
        Rt2 = np.zeros(12); Rt3 = np.zeros(12); rec = np.zeros(3 * n); rp = np.zeros(1); v = np.zeros(8, dtype=np.int32)
        st = hostcheck.hc_pose_tail(mode, dp(model), dp(cm(CalM)), dp(cm(Cr)), n, dp(Rt2), dp(Rt3), dp(rec), dp(rp),
                                    v.ctypes.data_as(C.POINTER(C.c_int)))
        assert st == 0
        got = (Rt2.reshape(3, 4, order="F"), Rt3.reshape(3, 4, order="F"), rec.reshape(n, 3).T, T, rp[0])
        assert_pose_close((R2, R3, Rec, T, rep), got, "n=%d noise=%g seed=%d" % (n, noise, seed))
        # votes: one candidate sees every point in front of both cameras
        assert sorted(v[:4])[-1] == 2 * n and sorted(v[4:])[-1] == 2 * n


def test_scene_generator_code_is_bit_exact_with_host_generator(hostcheck):
    """csrc/tvf_scene.cuh (MT19937 genrand_res53, polar Gaussian with the reproducible tvf_log, legacy shuffle,
    fixed-order projection, rejection loop, compaction, sub-sampling) compiled for the host ==
    tft_vs_fund_b200.scene, bit for bit -- noisy coordinates included."""
    from tft_vs_fund_b200 import scene
    n = 2000
    u = np.zeros(n); z = np.zeros(n); k = np.zeros(n, dtype=np.uint32)
    hostcheck.hc_mt_streams(C.c_uint(987654321), n, dp(u), dp(z), k.ctypes.data_as(C.POINTER(C.c_uint)), C.c_uint(119))
    rng = scene.SceneRNG(987654321)
    assert np.array_equal(u, rng.rs.random_sample(n)) and np.array_equal(z, rng.randn(1, n).ravel())
    rs = np.random.RandomState(987654321); rs.random_sample(n)
    assert np.abs(z - rs.standard_normal(n)).max() < 1e-14          # same stream as NumPy's legacy gauss up to log()'s last ulp
    assert k.max() <= 119
    K, Ps, _ = scene.scene_cameras(50, 0)
    P = np.ascontiguousarray(np.stack(Ps))
    d = scene.sweep_batch(13 * 12, 20)
    for j in range(13 * 12):
        out = np.zeros((20, 6))
        hostcheck.hc_scene_trial(dp(P), 20, C.c_double(d["noise"][j]), C.c_uint(int(d["seed"][j])), C.c_double(1800.0),
                                 C.c_double(1200.0), dp(out))
        assert np.array_equal(out.T, d["Corresp"][j]), j
    # the per-seed generator the CUDA kernel uses (first pass shared by the 13 levels) gives the same bits
    lv = np.ascontiguousarray(np.arange(0.0, 3.0 + 1e-9, 0.25))
    for seed in range(1, 13):
        out = np.zeros((13, 20, 6))
        hostcheck.hc_scene_seed(dp(P), 20, dp(lv), 0, 13, C.c_uint(seed), C.c_double(1800.0), C.c_double(1200.0), dp(out))
        assert np.array_equal(out.transpose(0, 2, 1), d["Corresp"][13 * (seed - 1):13 * seed]), seed
    out = np.zeros((4, 20, 6))                                  # a partial range of levels (shard boundaries)
    hostcheck.hc_scene_seed(dp(P), 20, dp(lv), 5, 9, C.c_uint(3), C.c_double(1800.0), C.c_double(1200.0), dp(out))
    assert np.array_equal(out.transpose(0, 2, 1), d["Corresp"][26 + 5:26 + 9])
    # many refill passes and state regenerations inside them: huge noise
    lvh = np.ascontiguousarray([3.0, 20.0, 60.0, 200.0, 500.0])
    dh = scene.sweep_batch(5 * 6, 12, noise_levels=lvh)
    for seed in range(1, 7):
        out = np.zeros((5, 12, 6))
        hostcheck.hc_scene_seed(dp(P), 12, dp(lvh), 0, 5, C.c_uint(seed), C.c_double(1800.0), C.c_double(1200.0), dp(out))
        assert np.array_equal(out.transpose(0, 2, 1), dh["Corresp"][5 * (seed - 1):5 * seed]), seed
    # another shape: n = 12 (experiments.m's default N), high noise -> more rejections
    d = scene.sweep_batch(26, 12, noise_levels=[3.0, 20.0])
    for j in range(26):
        out = np.zeros((12, 6))
        hostcheck.hc_scene_trial(dp(P), 12, C.c_double(d["noise"][j]), C.c_uint(int(d["seed"][j])), C.c_double(1800.0),
                                 C.c_double(1200.0), dp(out))
        assert np.array_equal(out.T, d["Corresp"][j]), j


@pytest.mark.parametrize("route", [2, 1, 0])
def test_certified_vote_signs_equal_the_accurate_route(hostcheck, route):
    """The votes' certified shortcuts -- route 2: dlt4_depth_signs_ray<true> (ray / plane intersection + sin-theta certificate
    in the form for a view-1 camera K1*[I | 0]: what the kernels run first), route 1: its general form, route 0: dlt4_depth_signs (normal equations + certificate) -- against the accurate
    Householder route, point by point and candidate by candidate: wherever a shortcut answers, its two signs are the
    accurate route's -- on the benchmark scenes, on the long-focal / collinear / minimal scenes of experiments.m:38-47,
    and under heavy noise; and it does answer for the overwhelming majority of regular points."""
    tot = np.zeros(3, dtype=np.int64)
    cases = [(20, nz, 50, 0) for nz in (0.0, 0.25, 1.0, 3.0)] + [(12, 1.0, f, 0) for f in (20, 100, 300)] + \
            [(12, 1.0, 50, a) for a in (166, 175, 179.5, 180)] + [(8, 3.0, 50, 0), (60, 10.0, 50, 0), (25, 30.0, 50, 0)] + \
            [(400, 1.0, 50, 0)]
    for n, noise, f, ang in cases:
        c = np.zeros(3, dtype=np.int64)
        for seed in range(1, 9):
            CalM, R_t0, Cr, _ = o.experiments_subsample(n, noise, seed, f, ang)
            try:
                T = o.LinearTFTPoseEstimation(Cr, CalM)[3]
            except RuntimeError:
                continue
            hostcheck.hc_vote_signs_check(route, 0, dp(T.ravel(order="F").copy()), dp(cm(CalM)), dp(cm(Cr)), n,
                                          c.ctypes.data_as(C.POINTER(C.c_longlong)))
            if n >= 8:
                F21, F31 = o.LinearFPoseEstimation(Cr, CalM, return_F=True)[5:7]
                hostcheck.hc_vote_signs_check(route, 1, dp(np.concatenate([cm(F21), cm(F31)])), dp(cm(CalM)), dp(cm(Cr)), n,
                                              c.ctypes.data_as(C.POINTER(C.c_longlong)))
        assert c[2] == 0, (n, noise, f, ang, c)
        if noise <= 3.0 and f == 50 and ang == 0 and n >= 12:
            assert c[1] >= 0.95 * c[0], (n, noise, f, ang, c)          # the benchmark scene: the shortcut almost always answers
        print("  route %d n=%d noise=%g f=%g angle=%g: %d DLTs, %.1f %% answered by the shortcut" % (route, n, noise, f, ang, c[0], 100.0 * c[1] / max(1, c[0])))
        tot += c
    print("certified sign shortcut (route %d): %d DLTs, %d answered (%.1f %%), %d disagreements" % (route, tot[0], tot[1], 100.0 * tot[1] / tot[0], tot[2]))
