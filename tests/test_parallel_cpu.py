"""CPU, world_size 2 over gloo: the host-side sharding logic of stag_b200.parallel (sample
shards, flat-bucket gradient all-reduce, MC mean, row partition with halo all-gather and dX
reduce-scatter).  The local operator is injected (the oracle), since there is no GPU here."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_spmm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stag_b200 import parallel as P
    try:
        # --- MC sample shards cover [0,S) exactly once --------------------------------------
        base, n = P.shard_samples(7, rank, world)
        got = [None] * world
        dist.all_gather_object(got, list(range(base, base + n)))
        assert sorted(sum(got, [])) == list(range(7))
        assert P.shard_items(5, rank, world) == list(range(rank, 5, world))
        # --- flat-bucket gradient averaging == mean of per-rank grads -------------------------
        torch.manual_seed(0)
        lin = torch.nn.Linear(4, 3)
        x = torch.full((2, 4), float(rank + 1))
        lin(x).sum().backward()
        n = P.allreduce_gradients(lin.parameters())
        assert n == 15
        assert torch.allclose(lin.weight.grad, torch.full((3, 4), 2 * 1.5))
        # --- the same through GradBucket: gradients are views of the flat buffer, no copies -----
        lin2 = torch.nn.Linear(4, 3)
        bucket = P.GradBucket(lin2.parameters())
        for _ in range(2):   # second round: zero() really clears what backward accumulated
            bucket.zero()
            lin2(x).sum().backward()
            assert bucket.allreduce() == 15
            assert lin2.weight.grad.data_ptr() == bucket.flat.data_ptr()
            assert torch.allclose(lin2.weight.grad, torch.full((3, 4), 2 * 1.5)) and torch.allclose(lin2.bias.grad, torch.full((3,), 2.0))
        # --- MC mean --------------------------------------------------------------------------
        m = P.mc_mean(torch.full((3, 2), float(rank + 1)) * 2, 4)
        assert torch.allclose(m, torch.full((3, 2), 1.5))
        # --- row partition: halo all-gather fwd, reduce-scatter bwd == unpartitioned oracle ----
        rng = np.random.default_rng(3)
        N, E, D = 37, 400, 6
        src, dst = torch.from_numpy(rng.integers(0, N, E)), torch.from_numpy(rng.integers(0, N, E))
        X = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32))
        W = torch.from_numpy((1 + 0.3 * rng.standard_normal((E, D))).astype(np.float32))
        G = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32))
        part = P.RowPartition(src, dst, N, rank, world, halo=False)
        xfull = part.gather_features(X[part.lo:part.hi].clone()).requires_grad_(True)
        assert torch.equal(xfull.detach(), X)
        # local aggregation over the owned destinations, noise indexed by GLOBAL edge id
        out = ref_spmm.aggregate(part.src, part.dst, N, xfull, W[part.edge_ids])
        own = out[part.lo:part.hi]
        own.backward(G[part.lo:part.hi])
        dx_block = part.scatter_gradients(xfull.grad)
        Xo = X.clone().requires_grad_(True)
        full = ref_spmm.aggregate(src, dst, N, Xo, W)
        full.backward(G)
        assert torch.allclose(own.detach(), full.detach()[part.lo:part.hi], atol=1e-6)
        assert torch.allclose(dx_block, Xo.grad[part.lo:part.hi], atol=1e-5)
        assert out.detach()[: part.lo].abs().sum() == 0 and out.detach()[part.hi:].abs().sum() == 0
        # --- halo form: only the referenced source rows travel, bipartite local graph ------------------------
        hp = P.RowPartition(src, dst, N, rank, world).setup_halo()
        assert hp.n_halo == len(set(hp.src.tolist()) - set(range(hp.lo, hp.hi))) and hp.n_ext == hp.n_own + hp.n_halo
        xb = X[hp.lo:hp.hi].clone().requires_grad_(True)
        x_ext = hp.exchange(xb.detach())
        assert torch.equal(x_ext[: hp.n_own], X[hp.lo:hp.hi]) and torch.equal(x_ext[hp.n_own:], X[hp.need])
        x_ext.requires_grad_(True)
        out_h = ref_spmm.aggregate(hp.src_ext, hp.dst_loc, hp.n_ext, x_ext, W[hp.edge_ids])[: hp.n_own]
        assert torch.equal(out_h.detach(), own.detach())        # same edges, same order: bitwise
        out_h.backward(G[hp.lo:hp.hi])
        dx_h = hp.exchange_back(x_ext.grad)
        assert torch.allclose(dx_h, Xo.grad[hp.lo:hp.hi], atol=1e-5)
        # --- the same with edge-balanced cuts: rows split so that the ranks hold about E / world in-edges each --------
        he = P.RowPartition(src, dst, N, rank, world, balance="edges").setup_halo()
        assert he.bounds[0] == 0 and he.bounds[-1] == N and abs(int(he.src.numel()) - src.numel() // world) <= int(torch.bincount(dst).max())
        xe = he.exchange(X[he.lo:he.hi].contiguous()).requires_grad_(True)
        out_e = ref_spmm.aggregate(he.src_ext, he.dst_loc, he.n_ext, xe, W[he.edge_ids])[: he.n_own]
        assert torch.equal(out_e.detach(), full.detach()[he.lo:he.hi])
        out_e.backward(G[he.lo:he.hi])
        assert torch.allclose(he.exchange_back(xe.grad), Xo.grad[he.lo:he.hi], atol=1e-5)
        # sample-batched exchange [S, rows, D]
        X3 = torch.stack([X, 2 * X])
        x3 = hp.exchange(X3[:, hp.lo:hp.hi].contiguous())
        assert torch.equal(x3[1, hp.n_own:], 2 * X[hp.need])
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_edge_balanced_blocks_single_process():
    import pytest
    from stag_b200 import parallel as P
    g = torch.Generator().manual_seed(3)
    N, E = 1000, 20000
    dst = (torch.rand(E, generator=g) ** 3 * N).long().clamp_(max=N - 1)      # skewed in-degrees
    for world in (1, 2, 3, 8):
        bounds, per = P.edge_balanced_blocks(dst, N, world)
        assert bounds[0] == 0 and bounds[-1] == N and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:])) and len(bounds) == world + 1
        cnt = [int(((dst >= bounds[r]) & (dst < bounds[r + 1])).sum()) for r in range(world)]
        assert sum(cnt) == E and max(cnt) - E // world <= int(torch.bincount(dst).max())
        assert per == max(b1 - b0 for b0, b1 in zip(bounds, bounds[1:]))
    part = P.RowPartition(torch.zeros(E, dtype=torch.long), dst, N, 1, 3, balance="edges")
    rows = torch.arange(N)
    own = part._owner(rows)
    assert all(int(own[b]) == r for r, b in enumerate(part.bounds[:-1]) if part.bounds[r + 1] > b)
    with pytest.raises(ValueError):
        P.RowPartition(torch.zeros(E, dtype=torch.long), dst, N, 0, 2, halo=False, balance="edges")


def test_row_blocks_and_sample_shards_single_process():
    from stag_b200 import parallel as P
    bounds, per = P.row_blocks(10, 4)
    assert bounds == [0, 3, 6, 9, 10] and per == 3
    assert [P.shard_samples(16, r, 8) for r in range(8)] == [(2 * r, 2) for r in range(8)]
    assert [P.shard_samples(3, r, 4)[1] for r in range(4)] == [1, 1, 1, 0]
