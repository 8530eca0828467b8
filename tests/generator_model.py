"""TEST INFRASTRUCTURE: a lane-level NumPy restatement of sweep_seeds_warp_kernel (tft_vs_fund_b200/csrc/
tvf_scene_kernels.cu) -- the same phases, the same stream bookkeeping (two-block state ring, look-ahead windows,
rewinds), the same ballot ranks, the same fixed-point acceptance of the shuffle, the same memoised / chunked refill
passes -- with a warp modelled as NumPy vectors of 32 lanes.  tests/test_generator_model.py compares it bit for bit
with the serial generators (NumPy host generator, host build of scene_trial), so the ALGORITHM of the warp kernel is
pinned on the CPU; the GPU tests then only have to show that the CUDA code implements it."""
import numpy as np

from tft_vs_fund_b200 import scene

LANES = np.arange(32)
CAP = 32
CAP2 = 8
U32 = np.uint32


def _temper(y):
    y = y.astype(np.uint32)
    y = y ^ (y >> U32(11))
    y = y ^ ((y << U32(7)) & U32(0x9d2c5680))
    y = y ^ ((y << U32(15)) & U32(0xefc60000))
    y = y ^ (y >> U32(18))
    return y


class WarpMT:
    """The warp's generator: raw[b & 1] holds block b of the MT19937 state sequence; the ring covers [cur, cur+1]."""

    def __init__(self, seed):
        s = int(seed) & 0xffffffff
        init = np.empty(624, dtype=np.uint32)
        for i in range(624):                                   # init_genrand (serial)
            init[i] = s
            s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xffffffff
        self.init = init
        self.raw = [np.zeros(624, dtype=np.uint32), np.zeros(624, dtype=np.uint32)]
        self.twists = 0
        self.rewinds = 0
        self.reset()

    def reset(self):
        self.raw[1] = self.init.copy()
        self.cur, self.have_next = -1, False

    def make_next(self):
        src, dst = self.raw[self.cur & 1], self.raw[(self.cur + 1) & 1]
        for s in range(20):                                    # 20 lane-parallel steps, out of place
            i = 32 * s + LANES
            i = i[i < 624]
            a = src[i]
            b = np.where(i < 623, src[np.minimum(i + 1, 623)], dst[0])
            m = np.where(i < 227, src[np.minimum(i + 397, 623)], dst[np.maximum(i - 227, 0)])
            y = (a & U32(0x80000000)) | (b & U32(0x7fffffff))
            dst[i] = m ^ (y >> U32(1)) ^ np.where(y & U32(1), U32(0x9908b0df), U32(0))
        self.have_next = True
        self.twists += 1

    def prepare(self, lo, hi):
        assert 0 <= lo <= hi and hi - lo < 624
        b_lo, b_hi = lo // 624, hi // 624
        if b_lo < self.cur:
            self.reset(); self.rewinds += 1
        while b_hi > self.cur + 1:
            if not self.have_next:
                self.make_next()
            self.cur += 1; self.have_next = False
        if b_hi == self.cur + 1 and not self.have_next:
            self.make_next()
        assert self.cur <= b_lo and b_hi <= self.cur + 1 and (b_lo > self.cur or self.cur >= 0)

    def words(self, pos):
        pos = np.asarray(pos)
        b = pos // 624
        assert np.all((b >= max(self.cur, 0)) & (b <= self.cur + 1))
        out = np.empty(pos.shape, dtype=np.uint32)
        for idx in np.ndindex(pos.shape):
            out[idx] = self.raw[int(b[idx]) & 1][int(pos[idx]) - 624 * int(b[idx])]
        return _temper(out)


def _res53(w0, w1):
    return ((w0 >> U32(5)).astype(np.float64) * 67108864.0 + (w1 >> U32(6)).astype(np.float64)) / 9007199254740992.0


def _points(mt, pos, Ps):
    """lane = point: six words each from `pos` (array), projected by the three cameras -> (len, 6)."""
    w = mt.words(pos[:, None] + np.arange(6)[None, :])
    X = np.stack([400.0 * _res53(w[:, 2 * k], w[:, 2 * k + 1]) + (-200.0) for k in range(3)])
    return np.vstack([scene._project(P, X) for P in Ps]).T


def _normals(mt, pos, M, i_lo, i_hi, buf):
    """sw_normals: 3*M accepted polar pairs from `pos`; pairs of points [i_lo, i_hi) stored in buf[i - i_lo, 2v:2v+2]."""
    need, got = 3 * M, 0
    while got < need:
        mt.prepare(pos, pos + 127)
        w = mt.words(pos + 4 * LANES[:, None] + np.arange(4)[None, :])
        x1 = 2.0 * _res53(w[:, 0], w[:, 1]) + (-1.0)
        x2 = 2.0 * _res53(w[:, 2], w[:, 3]) + (-1.0)
        r2 = x1 * x1 + x2 * x2
        acc = ~((r2 >= 1.0) | (r2 == 0.0))
        idx = got + np.cumsum(acc) - acc                       # got + popc(ballot & lanemask_lt)
        for lane in np.flatnonzero(acc & (idx < need)):
            v, i = divmod(int(idx[lane]), M)
            if i_lo <= i < i_hi:
                f = np.sqrt((-2.0 * scene.tvf_log(r2[lane:lane + 1])) / r2[lane:lane + 1])[0]
                buf[i - i_lo, 2 * v] = f * x2[lane]
                buf[i - i_lo, 2 * v + 1] = f * x1[lane]
        last = np.flatnonzero(acc & (idx == need - 1))
        if last.size:
            pos += 4 * (int(last[0]) + 1); got = need
        else:
            got += int(acc.sum()); pos += 128
    return pos


def _inside(p, hi_x, hi_y):
    x, y = p[:, 0::2], p[:, 1::2]
    return np.all((x <= hi_x) & (y <= hi_y) & (x >= 0.0) & (y >= 0.0), axis=1)


def seed_levels(seed, n, noise_levels, Ps, hi_x, hi_y, stats=None):
    """All noise levels of one seed, as the warp kernel computes them -> (L, n, 6)."""
    N = n + 100
    mt = WarpMT(seed)
    # ---- shuffle: fixed-point acceptance per batch of 32 words, serial swaps
    arr = np.arange(N)
    top, pos = N - 1, 0
    rounds = []
    while top >= 1:
        mt.prepare(pos, pos + 31)
        w = mt.words(pos + LANES)
        acc = np.zeros(32, dtype=bool)
        r = 0
        while True:
            my_top = top - (np.cumsum(acc) - acc)
            mask = np.array([(0xffffffff >> (32 - int(t).bit_length())) if t >= 1 else 0 for t in my_top], dtype=np.uint32)
            v = w & mask
            nacc = (my_top >= 1) & (v <= np.maximum(my_top, 0).astype(np.uint32))
            r += 1
            if np.array_equal(nacc, acc):
                break
            acc = nacc
        rounds.append(r)
        past = np.flatnonzero(my_top < 1)
        used = int(past[0]) if past.size else 32
        for lane in np.flatnonzero(acc):                       # lane 0 applies the swaps in order
            t, j = int(my_top[lane]), int(v[lane])
            arr[t], arr[j] = arr[j], arr[t]
        top -= int(acc.sum()); pos += used
    outpos = np.full(N, -1)
    outpos[arr[:n]] = np.arange(n)
    # ---- first pass
    clean = np.zeros((N, 6))
    for i0 in range(0, N, 32):
        i1 = min(i0 + 32, N)
        mt.prepare(6 * i0, 6 * i1 - 1)
        clean[i0:i1] = _points(mt, 6 * np.arange(i0, i1), Ps)
    z = np.zeros((N, 6))
    snap_pos = _normals(mt, 6 * N, N, 0, N, z)
    # ---- levels
    out = np.zeros((len(noise_levels), n, 6))
    # memo slot 0: first refill pass of a level (starts at snap_pos), up to CAP points; slot 1: second pass, up to CAP2
    m0_M, m0_end, m1_M, m1_pos, m1_end = -1, 0, -1, 0, 0
    cc, cz = np.zeros((CAP, 6)), np.zeros((CAP, 6))
    cc2, cz2 = np.zeros((CAP2, 6)), np.zeros((CAP2, 6))
    reused = 0

    def emit(p, filled, o):
        ins = _inside(p, hi_x, hi_y)
        rank = filled + np.cumsum(ins) - ins
        for t in np.flatnonzero(ins):
            k = outpos[rank[t]]
            if k >= 0:
                o[k] = p[t]
        return int(ins.sum())

    for lv, noise in enumerate(noise_levels):
        o = out[lv]
        filled = 0
        for i0 in range(0, N, 32):
            i1 = min(i0 + 32, N)
            filled += emit(clean[i0:i1] + z[i0:i1] * noise, filled, o)
        M, pos, npass = N - filled, snap_pos, 0
        while M > 0:
            s0 = npass == 0 and M <= CAP
            s1 = npass == 1 and M <= CAP2
            if s0 and m0_M == M:
                end_pos = m0_end; reused += 1
                filled += emit(cc[:M] + cz[:M] * noise, filled, o)
            elif s1 and m1_M == M and m1_pos == pos:
                end_pos = m1_end; reused += 1
                filled += emit(cc2[:M] + cz2[:M] * noise, filled, o)
            else:
                bc, bz = (cc2, cz2) if s1 else (cc, cz)
                if not s0 and not s1:
                    m0_M = -1                                  # the big buffers are about to be overwritten
                end_pos = pos
                for c0 in range(0, M, CAP):
                    c1 = min(c0 + CAP, M)
                    mt.prepare(pos + 6 * c0, pos + 6 * c1 - 1)
                    bc[:c1 - c0] = _points(mt, pos + 6 * np.arange(c0, c1), Ps)
                    end_pos = _normals(mt, pos + 6 * M, M, c0, c1, bz)
                    filled += emit(bc[:c1 - c0] + bz[:c1 - c0] * noise, filled, o)
                if s0:
                    m0_M, m0_end = M, end_pos
                if s1:
                    m1_M, m1_pos, m1_end = M, pos, end_pos
            pos, M, npass = end_pos, N - filled, npass + 1
    if stats is not None:
        stats.append(dict(twists=mt.twists, rewinds=mt.rewinds, shuffle_rounds=max(rounds), memo_reuses=reused))
    return out
