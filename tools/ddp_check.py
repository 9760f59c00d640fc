"""Multi-GPU checks over NCCL (run under torchrun, one rank per GPU):

1. MC-sample sharding: every rank computes StagModel.loss on its shard of the S samples (Philox sample
   indices [sample_base, sample_base + S/G)), gradients are averaged with ONE flat-bucket all-reduce
   (stag_b200.parallel.allreduce_gradients) and must equal the single-process gradients over all S samples.
2. MC predictive mean: all-reduce of the per-rank sums of the [N,C] outputs == single-process mean.
3. Row partition: halo all-gather of X blocks, local fused aggregation keyed by global edge ids,
   reduce-scatter of dX == the unpartitioned result.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import stag_b200 as stag
from stag_b200 import parallel as P


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max().clamp(min=1e-30))


def build(seed=0):
    torch.manual_seed(seed)
    mk = lambda d: torch.distributions.Normal(torch.ones(d), 0.3 * torch.ones(d))  # noqa: E731
    layers = torch.nn.ModuleList([
        stag.layers.StagLayer(stag.zoo.GCN(32, 64, activation=torch.relu), q_a=mk(32), p_a=mk(32), vi=True),
        stag.layers.StagLayer(stag.zoo.GCN(64, 10, activation=lambda x: torch.softmax(x, dim=-1)), q_a=mk(64), p_a=mk(64), vi=True),
    ]).cuda()
    return stag.models.StagModel(layers, kl_scaling=0.1), layers


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(0)
    N, E, S = 20000, 150000, 8
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    g = stag.Graph(torch.from_numpy(src), torch.from_numpy(dst), N).to("cuda")
    x = torch.from_numpy(rng.standard_normal((N, 32)).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.integers(0, 10, N)).cuda()

    # ---- 1 + 2: MC samples sharded over the ranks ---------------------------------------------------------
    model, layers = build()
    base, n_local = P.shard_samples(S, rank, world)
    stag.manual_seed(5)
    outs = model._forward_samples(g, x, n_local, sample_base=base)                       # [S/G, N, C]
    nll = sum(-model.likelihood.log_prob(outs[s], y).mean() for s in range(n_local)) / S  # this rank's share of the mean
    reg = sum(layer.kl_divergence() for layer in layers) * model.kl_scaling / world
    (nll + reg).backward()
    n_bucket = P.allreduce_gradients(layers.parameters(), average=False)
    mean_probs = P.mc_mean(outs.detach().sum(0), S)
    # single-process reference over all S samples (same seed -> same Philox offsets per layer)
    model1, layers1 = build()
    stag.manual_seed(5)
    outs1 = model1._forward_samples(g, x, S, sample_base=0)
    nll1 = sum(-model1.likelihood.log_prob(outs1[s], y).mean() for s in range(S)) / S
    reg1 = sum(layer.kl_divergence() for layer in layers1) * model1.kl_scaling
    (nll1 + reg1).backward()
    errs = {k: rel(p.grad, q.grad) for (k, p), (_, q) in zip(layers.named_parameters(), layers1.named_parameters())}
    err_mean = rel(mean_probs, outs1.detach().mean(0))

    # ---- 3: row partition ------------------------------------------------------------------------------------
    ts, td = torch.from_numpy(src), torch.from_numpy(dst)
    part = P.RowPartition(ts, td, N, rank, world, halo=False)
    lg = part.local_graph(stag.Graph).to("cuda")
    D = 64
    X = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32)).cuda()
    G = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32)).cuda()
    one, sg = torch.ones((), device="cuda"), torch.full((), 0.4, device="cuda")
    spec = lambda e: stag.ops.NoiseSpec("normal", one, sg, D, e, seed=9, offset=1)  # noqa: E731
    xfull = part.gather_features(X[part.lo:part.hi].contiguous()).requires_grad_(True)       # halo all-gather
    out = stag.ops.stochastic_aggregate(lg, xfull, spec(lg.number_of_edges()))
    out[part.lo:part.hi].backward(G[part.lo:part.hi])
    dx_block = part.scatter_gradients(xfull.grad)                                            # reduce-scatter
    Xr = X.clone().requires_grad_(True)
    full = stag.ops.stochastic_aggregate(g, Xr, spec(E))
    full.backward(G)
    err_out = rel(out[part.lo:part.hi], full[part.lo:part.hi])
    err_dx = rel(dx_block, Xr.grad[part.lo:part.hi])

    # ---- 4: row partition with a proper halo (only the referenced source rows travel; bipartite local graph) ----
    hp = P.RowPartition(ts.cuda(), td.cuda(), N, rank, world, balance="edges").setup_halo()
    hg = hp.local_graph(stag.Graph)
    x_ext = hp.exchange(X[hp.lo:hp.hi].contiguous()).requires_grad_(True)
    S = 3
    spec_s = lambda e: stag.ops.NoiseSpec("normal", one, sg, D, e, seed=9, offset=1, n_samples=S, batched=True)  # noqa: E731
    out_h = stag.ops.stochastic_aggregate(hg, x_ext, spec_s(hg.number_of_edges()), n_samples=S)     # [S, n_own, D]
    G3 = torch.stack([G, 0.5 * G, -G])
    out_h.backward(G3[:, hp.lo:hp.hi])
    dx_h = hp.exchange_back(x_ext.grad)
    Xs = X.clone().requires_grad_(True)
    full_s = stag.ops.stochastic_aggregate(g, Xs, spec_s(E), n_samples=S)
    full_s.backward(G3)
    halo_bitwise = float(torch.equal(out_h, full_s[:, hp.lo:hp.hi]))
    err_dx_h = rel(dx_h, Xs.grad[hp.lo:hp.hi])
    recv_b, _ = hp.halo_bytes(D)
    halo_frac = recv_b / float((world - 1) * hp.per * D * 4)

    res = torch.tensor([max(errs.values()), err_mean, err_out, err_dx, err_dx_h, 1.0 - halo_bitwise, halo_frac],
                       device="cuda", dtype=torch.float64)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("world=%d bucket=%d floats  max rel err: grads %.2e  mc-mean %.2e  row-partition out %.2e dx %.2e"
              % (world, n_bucket, *res.tolist()[:4]))
        print("halo form: forward bitwise == unpartitioned: %s   dx rel err %.2e   halo rows received = %.0f%% of a full all-gather"
              % ("yes" if res[5] == 0 else "NO", res[4], 100 * res[6]))
        ok = res[0] < 1e-4 and res[1] < 1e-5 and res[2] < 1e-6 and res[3] < 1e-5 and res[4] < 1e-5 and res[5] == 0
        print("DDP CHECK", "OK" if ok else "FAILED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
