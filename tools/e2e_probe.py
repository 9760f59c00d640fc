"""Probe of the end-to-end leg: per-step wall times with CUDA-event timestamps of the copy and compute phases."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import stag_b200 as sb

dev = torch.device("cuda", 0)
src, dst = bench.synth_graph()
S = 16
g = sb.Graph(torch.from_numpy(src), torch.from_numpy(dst), bench.N_NODES).to(dev)
st = g._s
st.csx(True), st.csx(False)
ss, ds = st.scale(False, "rsqrt"), st.scale(True, "rsqrt")
loc = torch.ones((), device=dev); scale = torch.full((), bench.SIGMA, device=dev)
x_host = torch.randn(bench.N_NODES, bench.WIDTH).pin_memory()
dx_host = torch.empty(bench.N_NODES, bench.WIDTH).pin_memory()

def ev():
    return torch.cuda.Event(enable_timing=True)

def compute(x):
    h = x
    for layer in range(bench.N_LAYERS):
        spec = sb.ops.NoiseSpec("normal", loc, scale, bench.WIDTH, bench.N_EDGES, n_samples=S, sample_base=0, batched=True)
        h = sb.ops.stochastic_aggregate(g, h, spec, src_scale=ss, dst_scale=ds, n_samples=S)
    obj = h.mean()
    obj.backward()
    return obj

for it in range(4):
    e = [ev() for _ in range(6)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e[0].record()
    x = x_host.to(dev, non_blocking=True).requires_grad_(True)
    e[1].record()
    obj = compute(x)
    e[2].record()
    dx_host.copy_(x.grad, non_blocking=True)
    e[3].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("serial step %d: wall %.2f ms | h2d %.2f compute %.2f d2h %.2f" % (it, (t1 - t0) * 1e3, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))

# copy on a side stream while computing
side = torch.cuda.Stream(dev)
xb = torch.empty(bench.N_NODES, bench.WIDTH, device=dev)
for it in range(3):
    torch.cuda.synchronize()
    a, b, c, d = ev(), ev(), ev(), ev()
    t0 = time.perf_counter()
    with torch.cuda.stream(side):
        c.record(side)
        xb.copy_(x_host, non_blocking=True)
        d.record(side)
    a.record()
    x = x_host.to(dev, non_blocking=True).requires_grad_(True)
    obj = compute(x)
    b.record()
    torch.cuda.synchronize()
    print("overlap step %d: wall %.2f ms | compute(+own h2d) %.2f side h2d %.2f" % (it, (time.perf_counter() - t0) * 1e3, a.elapsed_time(b), c.elapsed_time(d)))
