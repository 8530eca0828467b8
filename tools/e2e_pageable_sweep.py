#!/usr/bin/env python
"""End-to-end rate of the host-pointer call with PAGEABLE buffers (a MEX caller's mxArrays) against the number of host
staging threads and the chunk size.  Usage: e2e_pageable_sweep.py [chunk,chunk,... [threads,threads,...]]   (0 = library default; B = 1 M trials, n = 20)"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tft_vs_fund_b200 import _lib, scene

B, n = 1000000, 20
h = _lib.Handle(0); lib = h.lib
d = torch.empty((B, n, 6), dtype=torch.float64, device="cuda:0")
scene.sweep_batch_device(B, n, first_trial=0, device=0, out_ptr=d.data_ptr(), meta=False)
CalM = scene.generateSyntheticScene(n, 0.0, 1, 50, 0)[0]
h_in = d.cpu().numpy().copy()
h_calm = np.ascontiguousarray(CalM.T)
outs = [np.empty((B, 12)), np.empty((B, 12)), np.empty((B, 3 * n)), np.empty((B, 27)), np.empty((B,))]
st = np.zeros((B,), np.int32)
dp = lambda a: a.ctypes.data_as(_lib.c_double_p)
def step():
    return h.call("tvf_linear_tft_pose", dp(h_in), dp(h_calm), 0, n, B, dp(outs[0]), dp(outs[1]), dp(outs[2]), dp(outs[3]), dp(outs[4]),
                  st.ctypes.data_as(_lib.c_int32_p))
print("host cores", os.cpu_count(), flush=True)
chunks = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32768, 65536, 131072]
threads = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8, 12, 16, 24]
for chunk in chunks:                 # 0 = the library's default schedule / thread count
    for thr in threads:
        h.call("tvf_set_chunk", chunk); h.call("tvf_set_host_threads", thr)
        for _ in range(2): step()
        t0 = time.perf_counter()
        for _ in range(5): step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("chunk %7d threads %2d: %.4g solves/s (%.2f ms per 1 M)" % (chunk, thr, B * 5 / dt, dt / 5 * 1e3), flush=True)
