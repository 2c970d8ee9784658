"""N>1 host logic on CPU: world_size-2 gloo run of the column sharding + packed all-reduce.
Each rank forms the partial sums of ITS columns (numpy stands in for the CUDA kernels, which need a
GPU); after the all-reduce both ranks must hold the full-batch aggregates and produce identical W."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onmf_ontf_ndl_b200.parallel import allreduce_packed, pack_partial, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 1000, 262144):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import onmf_oracle as O
    rng = np.random.default_rng(0)              # same data on every rank
    d, k, n = 30, 6, 101
    X = rng.random((d, n)); H = rng.random((k, n)) * (rng.random((k, n)) < 0.3)
    W = rng.random((d, k)); A = np.zeros((k, k)); B = np.zeros((k, d))
    lo, hi = shard_range(n, world, rank)
    P = pack_partial(torch.from_numpy(H[:, lo:hi] @ H[:, lo:hi].T), torch.from_numpy(H[:, lo:hi] @ X[:, lo:hi].T))
    allreduce_packed(P)
    Pn = P.numpy()
    w = 1.0 / 3.0
    A1 = (1 - w) * A + w * Pn[:, :k]
    B1 = (1 - w) * B + w * Pn[:, k:]
    W1 = O.update_dict(W, A1, B1)
    Aref, Bref = O.aggregate(A, B, H, X, 3.0)
    ok = np.allclose(A1, Aref, rtol=1e-13, atol=1e-13) and np.allclose(B1, Bref, rtol=1e-13, atol=1e-13)
    gathered = [torch.zeros(d, k, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(W1))
    same = all(torch.equal(gathered[0], g) for g in gathered)       # bit-identical dictionaries on all ranks
    out[rank] = bool(ok and same)
    dist.destroy_process_group()


def test_two_rank_packed_allreduce_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
