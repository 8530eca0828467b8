"""Host-side multi-GPU logic on CPU: shard ranges, shard-local input generation, rank-ordered gather and
reduction over a world_size-2 gloo group (the N>1 path of bench.py / experiments.run_sweep)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_ranges_partition():
    from tft_vs_fund_b200.sharding import shard_range
    for total in (0, 1, 7, 13, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tft_vs_fund_b200 import sharding, scene, experiments
    lo, hi = sharding.shard_range(101, rank, world)
    d = scene.sweep_batch(hi - lo, 20, first_trial=lo)
    g = sharding.gather_arrays({"Corresp": d["Corresp"], "noise": d["noise"], "idx": np.arange(lo, hi)})
    t = sharding.max_over_ranks(1.0 + rank)

    def fake_solver(m, Corresp, CalM, R_t0):      # stands in for the GPU call: any per-trial numbers will do
        s = Corresp.reshape(Corresp.shape[0], -1)
        return s[:, 0] * m, s[:, 1], s[:, 2]

    table = experiments.run_sweep(101, 20, methods=(1, 7), solver=fake_solver)
    if rank == 0:
        np.savez(os.path.join(tmp, "out.npz"), Corresp=g["Corresp"], noise=g["noise"], idx=g["idx"], tmax=t,
                 t1=table[1], t7=table[7])
    else:
        assert g is None and table is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_and_reduce(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(os.path.join(str(tmp_path), "out.npz"))
    from tft_vs_fund_b200 import scene, experiments
    whole = scene.sweep_batch(101, 20)
    assert np.array_equal(out["idx"], np.arange(101))
    assert np.array_equal(out["Corresp"], whole["Corresp"]) and np.array_equal(out["noise"], whole["noise"])
    assert float(out["tmax"]) == 2.0
    # the sharded table equals the single-process one
    def fake_solver(m, Corresp, CalM, R_t0):
        s = Corresp.reshape(Corresp.shape[0], -1)
        return s[:, 0] * m, s[:, 1], s[:, 2]
    single = experiments.run_sweep(101, 20, methods=(1, 7), solver=fake_solver)
    assert np.allclose(out["t1"], single[1], rtol=1e-14) and np.allclose(out["t7"], single[7], rtol=1e-14)
    assert out["t1"].shape == (13, 3)
