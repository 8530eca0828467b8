#!/usr/bin/env python
"""End-to-end rate of the host-pointer call (pinned buffers, all outputs) against the chunk size of its H2D / kernels / D2H
pipeline.  Usage: e2e_chunk_sweep.py [chunk ...]   (B = 1 M trials, n = 20)"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tft_vs_fund_b200 import _lib, scene

B, n = 1000000, 20
chunks = [int(a) for a in sys.argv[1:]] or [8192, 16384, 32768, 65536, 131072]
h = _lib.Handle(0); lib = h.lib
lib.tvf_host_alloc.restype = C.c_void_p
def pinned(shape, dtype=np.float64):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.tvf_host_alloc(nbytes)
    return np.frombuffer((C.c_char * nbytes).from_address(p), dtype=dtype).reshape(shape)
d = torch.empty((B, n, 6), dtype=torch.float64, device="cuda:0")
scene.sweep_batch_device(B, n, first_trial=0, device=0, out_ptr=d.data_ptr(), meta=False)
CalM = scene.generateSyntheticScene(n, 0.0, 1, 50, 0)[0]
h_in = pinned((B, n, 6)); h_in[...] = d.cpu().numpy()
h_calm = np.ascontiguousarray(CalM.T)
outs = [pinned((B, 12)), pinned((B, 12)), pinned((B, 3 * n)), pinned((B, 27)), pinned((B,))]
st = pinned((B,), np.int32)
dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
def step():
    return h.call("tvf_linear_tft_pose", dp(h_in), dp(h_calm), 0, n, B, dp(outs[0]), dp(outs[1]), dp(outs[2]), dp(outs[3]), dp(outs[4]),
                  st.ctypes.data_as(_lib.c_int32_p))
for c in chunks:
    h.call("tvf_set_chunk", c)
    for _ in range(2): step()
    t0 = time.perf_counter()
    for _ in range(8): step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("chunk %7d: %.4g solves/s end to end (%.2f ms per 1 M)" % (c, B * 8 / dt, dt / 8 * 1e3), flush=True)
