"""Whole training steps of the reference's arxiv models on the arxiv-shaped synthetic graph, through the drop-in module API:
scripts/arxiv_mle/gcn/run.py (Normal(1, std) edge noise) and scripts/arxiv_rec/gcn/run.py (AmortizedDistribution(in, 1)
posteriors), 3 stag GCN layers 128-128-128-40 with BatchNorm / ReLU / Dropout in between, n_samples_training = 1
(the scripts' default) and 16, model.loss + backward + Adam step.  Device time per step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import stag_b200 as stag  # noqa: E402

dev = torch.device("cuda", 0)
src, dst = bench.synth_graph()
g = stag.Graph(torch.from_numpy(src), torch.from_numpy(dst), bench.N_NODES).to(dev)
N, D, H, C = bench.N_NODES, 128, 128, 40
feat = torch.randn(N, D, device=dev)
y = torch.randint(0, C, (N,), device=dev)
mask = torch.rand(N, device=dev) < 0.54


def build(kind, std=0.4):
    p_a = torch.distributions.Normal(1.0, std, validate_args=False)

    def q(width):
        return p_a if kind == "mle" else stag.distributions.AmortizedDistribution(width, 1, init_like=p_a)

    def mid():
        return stag.layers.FeatOnlyLayer(torch.nn.Sequential(torch.nn.BatchNorm1d(H), torch.nn.ReLU(), torch.nn.Dropout(0.5)))
    layers = torch.nn.ModuleList([
        stag.layers.StagLayer(stag.zoo.GCN(D, H), q_a=q(D)), mid(),
        stag.layers.StagLayer(stag.zoo.GCN(H, H), q_a=q(H), p_a=p_a), mid(),
        stag.layers.StagLayer(stag.zoo.GCN(H, C, activation=lambda x: torch.nn.functional.softmax(x, dim=-1)),
                              q_a=q(H), p_a=p_a)])
    return stag.models.StagModel(layers=layers).to(dev)


for kind in ("mle", "rec"):
    for S in (1, 16):
        model = build(kind)
        opt = torch.optim.Adam(model.parameters(), 1e-3)

        def step():
            opt.zero_grad()
            loss = model.loss(g, feat, y=y, mask=mask, n_samples=S)
            loss.backward()
            opt.step()
            return loss
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 10 if S == 1 else 4
        for _ in range(n):
            loss = step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print("arxiv_%s  n_samples_training %2d  %8.2f ms / training step  (%.3f GEdge-samples/s over 3 layers fwd+bwd)  loss %.4f"
              % (kind, S, ms, bench.N_EDGES * S * 3 / ms / 1e6, float(loss)))
