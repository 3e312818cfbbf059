"""N>1 path on CPU: two gloo ranks each solve their contiguous shard (kernel logic through the CPU
emulator) and gather exit flags; the result must equal the single-process batch."""
import os
import sys

import numpy as np
import pytest

from conftest import EMU_LIB, ROOT


def _worker(rank, world, port, batch, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from eicos_b200.binding import BatchSolver, Library
    from eicos_b200.sharding import gather_exit_flags, max_over_ranks, shard_range
    from eicos_b200.workloads import perturbed
    P = oracle.load_fixture("update_data_1")
    W = perturbed(P, batch, rel=0.05, seed=11)  # same seed on every rank -> same global batch
    lo, hi = shard_range(batch, rank, world)
    out = BatchSolver(P, lib=Library(EMU_LIB), capacity=8).solve(hi - lo, hs=W["hs"][lo:hi], bs=W["bs"][lo:hi])
    flags = gather_exit_flags(out["exit"], batch)
    tmax = max_over_ranks(1.0 + rank)
    if rank == 0:
        ret["flags"] = flags
        ret["x0"] = out["x"]
        ret["tmax"] = tmax
    dist.destroy_process_group()


def test_shard_ranges():
    from eicos_b200.sharding import shard_range
    for batch in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            r = [shard_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_matches_single_process(emu_lib, oracle_mod):
    import torch.multiprocessing as mp
    from eicos_b200.binding import BatchSolver
    from eicos_b200.workloads import perturbed
    batch = 21
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, batch, ret), nprocs=2, join=True)
    P = oracle_mod.load_fixture("update_data_1")
    W = perturbed(P, batch, rel=0.05, seed=11)
    single = BatchSolver(P, lib=emu_lib, capacity=32).solve(batch, hs=W["hs"], bs=W["bs"])
    assert np.array_equal(ret["flags"], single["exit"])
    assert np.array_equal(ret["x0"], single["x"][:11])  # rank 0 owns the first 11 of 21; bit-identical
    assert ret["tmax"] == 2.0
