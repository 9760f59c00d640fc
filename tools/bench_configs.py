"""Fused stochastic aggregation (forward + transposed pass) on a synthetic graph of every shape BASELINE.json
names (C1 Cora, C2 arxiv, C3 PPI minibatch, C4 molhiv batch, C5 products), one GPU, through
the public operator (stag_b200.ops.stochastic_aggregate + autograd), features resident on the device.

One STEP = every aggregation of the configuration's layers, forward and backward, over S Monte-Carlo samples:
GEdge-samples/s = E * S * layers / time; the fraction of the HBM roof uses the algorithmic bytes of SURVEY 8(d)
(per edge-sample, mle / vi) against MEASURED_PEAKS.json.  Small configurations are launch-bound: their step time
in microseconds is the figure to read."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import stag_b200 as sb
from stag_b200.ops import NoiseSpec

dev = torch.device("cuda", 0)
T = torch.from_numpy
peak, _ = bench.peaks()


def powerlaw_graph(N, E, seed, max_deg):
    rng = np.random.default_rng(seed)

    def endpoints():
        p = np.arange(1, N + 1, dtype=np.float64) ** (-1.0 / 1.1)
        p /= p.sum()
        p = np.minimum(p, max_deg / E)
        p /= p.sum()
        cdf = np.cumsum(p)
        return rng.permutation(N)[np.minimum(np.searchsorted(cdf, rng.random(E)), N - 1)]
    return endpoints().astype(np.int64), endpoints().astype(np.int64)


def batched_graphs(sizes, edges_per_node, seed):
    rng = np.random.default_rng(seed)
    gs = []
    for n in sizes:
        e = max(1, int(edges_per_node * n / 2))
        s, d = rng.integers(0, n, e), rng.integers(0, n, e)
        gs.append(sb.Graph(T(np.concatenate([s, d])), T(np.concatenate([d, s])), int(n)))
    return sb.batch(gs)


def bytes_per_edge_sample(N, E, D, vi):
    nd = 4.0 * N * D
    fwd = 2 * nd + 4 * E + 4 * (N + 1) + 8 * N
    bwd = (3 if vi else 2) * nd + 8 * E + 4 * (N + 1) + 8 * N
    return (fwd + bwd) / E


def run(name, g, widths, S, vi, per_channel, iters):
    g = g.to(dev)
    N, E = g.number_of_nodes(), g.number_of_edges()
    st = g._s
    st.csx(True), st.csx(False)
    ss, ds = st.scale(False, "rsqrt"), st.scale(True, "rsqrt")
    gen = torch.Generator(device=dev).manual_seed(7)
    layers = []
    for li, D in enumerate(widths):
        shape = (D,) if per_channel else ()
        loc = torch.ones(shape, device=dev).requires_grad_(vi)
        scale = torch.full(shape, 0.4, device=dev).requires_grad_(vi)
        x = torch.randn((N, D) if li == 0 else (S, N, D), device=dev, generator=gen).requires_grad_(True)
        gout = torch.randn(S, N, D, device=dev, generator=gen)
        layers.append((D, loc, scale, x, gout))

    def step():
        for D, loc, scale, x, gout in layers:
            spec = NoiseSpec("normal", loc, scale, D, E, n_samples=S, batched=True)
            out = sb.ops.stochastic_aggregate(g, x, spec, src_scale=ss, dst_scale=ds, n_samples=S)
            out.backward(gout)
            x.grad = None

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    es = E * S * len(widths)
    nbytes = sum(bytes_per_edge_sample(N, E, D, vi) for D in widths) * E * S
    row = {"config": name, "N": N, "E": E, "widths": widths, "S": S, "mode": "vi" if vi else "mle", "ms_per_step": ms,
           "GEdge_samples_per_s": es / ms / 1e6, "frac_of_hbm_roof": nbytes / ms / 1e6 / peak}
    print("%-44s N %8d E %9d D %-16s S %2d %-3s %10.3f ms/step %8.3f GEdge-samples/s %5.1f%% of HBM roof" % (
        name, N, E, widths, S, row["mode"], ms, row["GEdge_samples_per_s"], 100 * row["frac_of_hbm_roof"]), flush=True)
    return row


rows = []
rng = np.random.default_rng(0)
# C1 Cora (scripts/citation_mle: widths 1433 -> 16 -> 16 -> 7, S 4)
s, d = powerlaw_graph(2708, 10556, 0x57A6 + 1, 168)
rows.append(run("C1 Cora-shaped, 3 layers", sb.Graph(T(s), T(d), 2708), [1433, 16, 16], 4, False, False, 50))
# C2 arxiv (the bench.py workload)
s, d = bench.synth_graph()
rows.append(run("C2 arxiv-shaped, 3 layers", sb.Graph(T(s), T(d), bench.N_NODES), [128, 128, 128], 16, False, False, 5))
rows.append(run("C2 arxiv-shaped, 3 layers, learned rc noise", sb.Graph(T(s), T(d), bench.N_NODES), [128, 128, 128], 16, True, True, 5))
# C3 PPI: minibatch of 2 graphs of the 24 (56 944 nodes, 818 716 edges in all), widths 50 -> 256 -> 256 -> 121
sizes = rng.integers(1000, 3500, 2)
rows.append(run("C3 PPI-shaped minibatch (2 graphs)", batched_graphs(sizes, 28.7, 3), [50, 256, 256], 4, False, False, 50))
sizes = rng.integers(1000, 3500, 24)
rows.append(run("C3 PPI-shaped, all 24 graphs in one batch", batched_graphs(sizes, 28.7, 3), [50, 256, 256], 4, False, False, 20))
# C4 molhiv: batches of 32 / 128 molecule-sized graphs, per-channel learned Normal (vi), widths 9 -> H -> H
for nb, H in ((32, 16), (128, 256), (4096, 256)):
    sizes = np.clip(rng.normal(25.5, 12, nb), 2, 80).astype(int)
    rows.append(run("C4 molhiv-shaped batch of %d, H %d" % (nb, H), batched_graphs(sizes, 2.15, 4), [9, H], 4, True, True, 50))
# C5 products: 4 of the 32 samples (one GPU's share at 8 GPUs), one layer
s, d = powerlaw_graph(2449029, 61859140, 0x57A6 + 5, 17000)
rows.append(run("C5 products-shaped, 1 layer, 4 of 32 samples", sb.Graph(T(s), T(d), 2449029), [100], 4, False, False, 3))
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "configs.json"), "w"), indent=1)
