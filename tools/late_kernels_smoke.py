"""Small invocations of the kernels added late in round 1 (a target for compute-sanitizer where the pool allows it): both forms of the tensor-core
generator, the likelihood epilogue, the edge-parallel per-edge gradient kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stag_b200 as sb  # noqa: E402

dev = "cuda"
g = torch.Generator().manual_seed(0)
N, E, D, S = 600, 9000, 128, 2
src, dst = torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)
dst[:1500] = 7
gr = sb.Graph(src, dst, N).to(dev)
x = torch.randn(S, N, D, generator=g).to(dev).requires_grad_(True)
one, sg = torch.ones((), device=dev), torch.full((), 0.4, device=dev)
sp = sb.ops.NoiseSpec("normal", one, sg, D, E, generator="hadamard", n_samples=S, batched=True)
out = sb.ops.stochastic_aggregate(gr, x, sp, n_samples=S)
out.sum().backward()
w = sp.materialize(n_samples=S)
# per-edge parameters with gradients, all samples in one launch
loc = torch.ones(E, 1, device=dev, requires_grad=True)
scale = torch.full((E, 1), 0.3, device=dev, requires_grad=True)
sp2 = sb.ops.NoiseSpec("normal", loc, scale, D, E, n_samples=S, batched=True)
out2 = sb.ops.stochastic_aggregate(gr, x, sp2, n_samples=S)
out2.sum().backward()
# likelihood epilogue
probs = torch.rand(S, N, 7, device=dev, requires_grad=True)
y = torch.randint(0, 7, (N,), device=dev)
sb.ops.fused_nll(probs, y, torch.rand(N, device=dev) < 0.5, "categorical").sum().backward()
sb.ops.fused_nll(probs, (torch.rand(N, 7, device=dev) < 0.5).float(), None, "bernoulli").sum().backward()
torch.cuda.synchronize()
print("ok", float(out.sum()), float(out2.sum()), float(w.mean()))
