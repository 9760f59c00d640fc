"""Top CUDA kernels of one arxiv_mle training step (16 samples) by device time: where a whole model step goes
outside the fused aggregation."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.argv = [sys.argv[0]]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import stag_b200 as stag  # noqa: E402

dev = torch.device("cuda", 0)
src, dst = bench.synth_graph()
g = stag.Graph(torch.from_numpy(src), torch.from_numpy(dst), bench.N_NODES).to(dev)
N, D, H, C, S = bench.N_NODES, 128, 128, 40, 16
feat = torch.randn(N, D, device=dev)
y = torch.randint(0, C, (N,), device=dev)
mask = torch.rand(N, device=dev) < 0.54
p_a = torch.distributions.Normal(1.0, 0.4, validate_args=False)
mid = lambda: stag.layers.FeatOnlyLayer(torch.nn.Sequential(torch.nn.BatchNorm1d(H), torch.nn.ReLU(), torch.nn.Dropout(0.5)))  # noqa: E731
layers = torch.nn.ModuleList([
    stag.layers.StagLayer(stag.zoo.GCN(D, H), q_a=p_a), mid(),
    stag.layers.StagLayer(stag.zoo.GCN(H, H), q_a=p_a, p_a=p_a), mid(),
    stag.layers.StagLayer(stag.zoo.GCN(H, C, activation=lambda x: torch.nn.functional.softmax(x, dim=-1)), q_a=p_a, p_a=p_a)])
model = stag.models.StagModel(layers=layers).to(dev)
opt = torch.optim.Adam(model.parameters(), 1e-3)


def step():
    opt.zero_grad()
    loss = model.loss(g, feat, y=y, mask=mask, n_samples=S)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
