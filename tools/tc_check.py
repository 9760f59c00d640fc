"""GPU check of the tensor-core noise path (STAG_NOISE_NORMAL_HADAMARD): fused kernel vs a float64 torch aggregation
fed the emitted noise, bitwise exactness of the tensor-core sums against the emitted stream, and timings at the arxiv
shape against the Box-Muller kernel.  (The comparison of the emitted stream with its numpy restatement lives in
tests/test_gpu_hadamard.py: only the tests use oracle/.)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stag_b200 as sb                      # noqa: E402
from stag_b200.ops import NoiseSpec, stochastic_aggregate   # noqa: E402

dev = torch.device("cuda", 0)


def spec(p0, p1, K, E, **kw):
    t = lambda v: torch.as_tensor(v, dtype=torch.float32).to(dev)   # noqa: E731
    return NoiseSpec("normal", t(p0), t(p1), K, E, generator="hadamard", **kw)


def dense_ref(src, dst, N, x, w, ss, ds):
    # out[s,v,c] = ds[v] * sum_e w[s,e,c] * ss[u] * x[s,u,c]
    S = w.shape[0]
    out = torch.zeros((S, N, x.shape[-1]), dtype=torch.float64, device=dev)
    for s in range(S):
        xs = (x if x.dim() == 2 else x[s]).double()
        m = w[s].double() * (xs[src] * ss.double()[src, None])
        out[s].index_add_(0, dst, m)
        out[s] *= ds.double()[:, None]
    return out


def check_fused(N, E, D, S, shared, hub=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, N, (E,), generator=g)
    dst = torch.randint(0, N, (E,), generator=g)
    if hub:
        dst[: E // 3] = 7
        src[E // 3: E // 2] = 11
    gr = sb.Graph(src, dst, N).to(dev)
    x = torch.randn((N, D) if shared else (S, N, D), generator=g).to(dev)
    ss = torch.rand(N, generator=g).to(dev) + 0.5
    ds = torch.rand(N, generator=g).to(dev) + 0.5
    sp = spec(1.0, 0.4, D, E, seed=11, offset=5, n_samples=S, batched=True)
    out = stochastic_aggregate(gr, x, sp, src_scale=ss, dst_scale=ds, n_samples=S)
    w = sp.materialize(n_samples=S)
    ref = dense_ref(src.to(dev), dst.to(dev), N, x, w, ss, ds)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    print("fused N=%d E=%d D=%d S=%d shared=%s hub=%s: rel err %.3g" % (N, E, D, S, shared, hub, err))
    # transposed pass through autograd
    xr = x.clone().requires_grad_(True)
    o2 = stochastic_aggregate(gr, xr, sp, src_scale=ss, dst_scale=ds, n_samples=S)
    go = torch.randn(o2.shape, generator=g).to(dev)
    o2.backward(go)
    xd = x.double().clone().requires_grad_(True)
    dense_ref(src.to(dev), dst.to(dev), N, xd, w, ss, ds).backward(go.double())
    errg = (xr.grad.double() - xd.grad).abs().max().item() / xd.grad.abs().max().item()
    print("   dX rel err %.3g" % errg)
    return max(err, errg)


def check_exact():
    # one in-edge per node, x = 1, loc = 0, scale = 1, no degree scales: out[v, c] == w[e(v), c] bit for bit
    N, D = 1000, 128
    src = torch.randperm(N)
    dst = torch.arange(N)
    gr = sb.Graph(src, dst, N).to(dev)
    sp = spec(0.0, 1.0, D, N, seed=3, offset=9, n_samples=2, batched=True)
    x = torch.ones(N, D, device=dev)
    out = stochastic_aggregate(gr, x, sp, n_samples=2)
    w = sp.materialize(n_samples=2)
    print("tensor-core sums bitwise equal to the emitted stream:", torch.equal(out, w),
          " max|diff| %.3g" % (out - w).abs().max().item())


def timing():
    import bench
    src, dst = bench.synth_graph()
    for normal in ("hadamard", "boxmuller"):
        path = bench.Path(dev, src, dst, 16, 0, False, normal=normal)
        for _ in range(3):
            path.step()
        torch.cuda.synchronize()
        ev = []
        for _ in range(5):
            path.step(ev)
        torch.cuda.synchronize()
        per = {}
        for tag, a, b in ev:
            per.setdefault(tag, []).append(a.elapsed_time(b))
        print(normal, {k: round(float(np.mean(v)), 4) for k, v in per.items()},
              "step ms %.3f" % sum(float(np.mean(v)) * (1 if k == "fwd_shared" else (2 if k == "fwd" else 3)) for k, v in per.items()))
        if normal == "hadamard":
            a_ = path.act[2].clone()
        else:
            b_ = path.act[2]
    print("act[2] stats hadamard mean %.4g std %.4g | boxmuller mean %.4g std %.4g" % (a_.mean().item(), a_.std().item(), b_.mean().item(), b_.std().item()))


if __name__ == "__main__":
    what = sys.argv[1:] or ["exact", "fused", "timing"]
    if "exact" in what:
        check_exact()
    if "fused" in what:
        check_fused(500, 3000, 128, 3, True)
        check_fused(500, 3000, 128, 3, False)
        check_fused(3000, 40000, 256, 2, False, hub=True)
        check_fused(20000, 300000, 128, 4, False, hub=True)
        check_fused(300, 100, 128, 1, True)
    if "timing" in what:
        timing()
    torch.cuda.synchronize()
    print("done")
