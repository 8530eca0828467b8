"""PCIe ceiling for the end-to-end path: pinned host <-> device copies of the bench's per-step volumes
(960 MB in, 900 MB out), each direction alone and both at once, whole buffers and 63 / 59 MB chunks."""
import json
import torch

dev = torch.device("cuda", 0)
nin, nout = 960_000_000 // 8, 900_000_000 // 8
h_in = torch.empty(nin, dtype=torch.float64).pin_memory(); h_in.fill_(1.0)
h_out = torch.empty(nout, dtype=torch.float64).pin_memory()
d_in = torch.empty(nin, dtype=torch.float64, device=dev)
d_out = torch.ones(nout, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def h2d(chunks=1):
    with torch.cuda.stream(s1):
        for c in range(chunks):
            lo, hi = c * nin // chunks, (c + 1) * nin // chunks
            d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)


def d2h(chunks=1):
    with torch.cuda.stream(s2):
        for c in range(chunks):
            lo, hi = c * nout // chunks, (c + 1) * nout // chunks
            h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)


res = {}
for chunks in (1, 16):
    t = timed(lambda: (h2d(chunks), s1.synchronize()))
    res["h2d_alone_gbs_%dchunks" % chunks] = 0.96 / t
    t = timed(lambda: (d2h(chunks), s2.synchronize()))
    res["d2h_alone_gbs_%dchunks" % chunks] = 0.90 / t
    t = timed(lambda: (h2d(chunks), d2h(chunks), s1.synchronize(), s2.synchronize()))
    res["both_seconds_%dchunks" % chunks] = t
    res["both_h2d_gbs_%dchunks" % chunks] = 0.96 / t
    res["both_solves_per_s_ceiling_%dchunks" % chunks] = 1e6 / t
print(json.dumps(res))
