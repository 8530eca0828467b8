"""The warp-per-seed scene generator's ALGORITHM (tests/generator_model.py: lane-level NumPy restatement of
sweep_seeds_warp_kernel) against the serial generators, bit for bit, on the CPU."""
import ctypes as C

import numpy as np
import pytest

from tft_vs_fund_b200 import scene
import generator_model as gm


def _cams():
    K, Ps, _ = scene.scene_cameras(50, 0)
    return Ps


def test_sweep_levels_equal_host_generator():
    """experiments.m's 13 noise levels for a few seeds: same bits as the NumPy host generator (scene.sweep_batch)."""
    levels = np.arange(0.0, 3.0 + 1e-9, 0.25)
    Ps, stats = _cams(), []
    for seed in (1, 2, 17, 4242):
        host = scene.sweep_batch(13, 20, first_trial=13 * (seed - 1), workers=1)["Corresp"]       # (13, 6, 20)
        got = gm.seed_levels(seed, 20, levels, Ps, 36 * scene.PIX, 24 * scene.PIX, stats)
        assert np.array_equal(got.transpose(0, 2, 1), host), seed
    assert all(s["rewinds"] == 0 for s in stats)                 # the two-block ring suffices at the reference's parameters
    assert sum(s["memo_reuses"] for s in stats) > 40             # most levels share their refill passes
    assert max(s["shuffle_rounds"] for s in stats) <= 16


@pytest.mark.parametrize("n,image,levels", [(20, (1100.0, 800.0), [0.0, 1.0, 2.5]), (45, (1000.0, 900.0), [0.5, 30.0]),
                                            (12, (1800.0, 1200.0), [3.0, 20.0, 60.0])])
def test_many_rejections_equal_serial_c_generator(hostcheck, n, image, levels):
    """Small images / large noise: refill passes with more than 32 points (chunked path), several passes per level,
    levels whose refill passes differ in size, streams over dozens of state blocks -- against the host build of the
    serial scene_trial (tests/hostcheck), which tests/test_scene.py pins to the NumPy generator."""
    Ps = _cams()
    P = np.ascontiguousarray(np.stack(Ps), dtype=np.float64)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    out = np.zeros(6 * n)
    stats = []
    for seed in (1, 2, 3):
        got = gm.seed_levels(seed, n, levels, Ps, image[0], image[1], stats)
        for lv, noise in enumerate(levels):
            hostcheck.hc_scene_trial(dp(P), n, C.c_double(noise), C.c_uint(seed), C.c_double(image[0]), C.c_double(image[1]), dp(out))
            assert np.array_equal(out.reshape(n, 6), got[lv]), (seed, noise)
    if image[0] < 1800.0:
        assert max(s["twists"] for s in stats) > 12              # long streams: the ring moved far beyond the snapshot block


def test_state_ring_matches_mt19937():
    """The lane-parallel, out-of-place state regeneration and the ring's look-ahead / rewind bookkeeping reproduce
    NumPy's MT19937 output stream at arbitrary (also backward) positions."""
    ref = np.random.RandomState(99).randint(0, 2 ** 32, size=6000, dtype=np.uint64).astype(np.uint32)
    mt = gm.WarpMT(99)
    for lo in (0, 600, 1200, 623, 2000, 100, 5000, 4400, 5375):
        mt.prepare(lo, lo + 623)
        pos = lo + np.arange(624)
        assert np.array_equal(mt.words(pos), ref[pos]), lo
    assert mt.rewinds == 3
