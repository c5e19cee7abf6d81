"""CPU, world_size 2 over gloo: the p-table exchange of the multi-GPU path
(one all-gather; every rank ends up with the same joined table) and the joined
q mapping equals the reference's file-based join (oracle)."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from graph_peak_caller_b200 import multigraph
from oracle import pipeline as opipeline, stats as ostats


def chromosome_tracks(rank):
    rng = np.random.default_rng(100 + rank)
    n = 400 + 50 * rank
    idx = np.r_[0, np.sort(rng.choice(np.arange(1, 50000), size=n - 1, replace=False))]
    val = np.round(rng.exponential(1.0, size=n), 1)        # many repeated p-values
    val[rng.random(n) < 0.4] = 0.0
    return idx, val, 50000 + 1000 * rank


def local_table(idx, val, size):
    lens = ostats.run_lengths(idx, size)
    nz = val != 0
    p, inv = np.unique(val[nz], return_inverse=True)
    return p, np.bincount(inv, weights=lens[nz]).astype(np.int64)


def worker(rank, world, port, out_dir, capacity=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    if capacity is not None:
        multigraph._capacity = capacity        # too small: the exchange has to grow it
    idx, val, size = chromosome_tracks(rank)
    p, cnt = local_table(idx, val, size)
    gp, gc, total = multigraph.gather_p_tables(torch.from_numpy(p), torch.from_numpy(cnt), size)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), p=gp.numpy(), c=gc.numpy(), total=total)
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("capacity", [None, 4])
def test_gather_p_tables_world_size_2(tmp_path, capacity):
    port = 29500 + os.getpid() % 500 + (0 if capacity is None else 500)
    mp.spawn(worker, args=(2, port, str(tmp_path), capacity), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(r0["p"], r1["p"]) and np.array_equal(r0["c"], r1["c"])
    assert int(r0["total"]) == int(r1["total"]) == 50000 + 51000
    # the joined table gives the reference's joined mapping (sparsepvalues.py:76-93)
    chroms = [chromosome_tracks(r) for r in range(2)]
    sp, cum = opipeline.joined_p_table(chroms)
    want = ostats.q_table(sp, cum)
    # rebuild from the exchanged (p, bp) pairs + the p == 0 remainder
    p, c = r0["p"], r0["c"]
    rest = int(r0["total"]) - int(c.sum())
    sp2, cum2 = ostats.p_table(np.r_[p, 0.0], np.r_[c, rest])
    got = ostats.q_table(sp2, cum2)
    assert got == want


def test_single_process_is_identity():
    p = torch.tensor([3.0, 1.5], dtype=torch.float64)
    c = torch.tensor([10, 20], dtype=torch.int64)
    gp, gc, total = multigraph.gather_p_tables(p, c, 100)
    assert gp is p and gc is c and total == 100
