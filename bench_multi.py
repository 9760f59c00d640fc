"""Multi-GPU legs of bench.py (``--gpus N`` with N > 1): every number below has a collective inside its timed region.

SURVEY.md 8(e) names three axes along which the stochastic aggregation shards; the reference has no distributed
code (single ``cuda:0``), the loops that are sharded are ``/root/reference/stag/models.py:67-78`` (MC samples) and
``/root/reference/scripts/ppi_mle/run.py:69-76`` (minibatches).

  strong_mc      STRONG scaling over the Monte-Carlo samples: the S samples of ONE step are split over the ranks
                 (``parallel.shard_samples``; rank r draws Philox sample indices [base_r, base_r + S/N)), every step
                 ends with the flat-bucket gradient all-reduce and the all-reduce of the [N,C] probability sums (the MC
                 predictive mean); C2 arxiv (16 samples, 3 layers) and C5 products (32 samples, 1 layer, graph replicated).
  minibatch_dp   C3 PPI: every rank trains on its own minibatch of 2 PPI-shaped graphs through the module API
                 (StagModel.loss -> backward -> ``parallel.GradBucket.allreduce`` -> Adam.step), weak scaling.
  row_partition  C5 products, ONE graph 1-D row-partitioned (``parallel.RowPartition``): halo exchange of the source
                 rows a rank's edges reference (one variable-size all-to-all), local fused aggregation of the owned rows
                 over all 32 samples, transposed pass, the partial gradients of the halo rows sent back to their owners.

Timing: CUDA events on the compute stream (torch makes it wait for each collective), max over ranks.  ``compute_ms`` is
the same step without the collectives, ``collective_ms`` the collectives alone; ``exposed_ms`` = step - compute.
"""
import ctypes
import time

import numpy as np
import torch

import stag_b200 as sb
from stag_b200 import _lib
from stag_b200 import parallel as P


# ---- synthetic graphs, generated ON DEVICE from a seed (every rank builds the same graph) -----------------------
def powerlaw_graph_device(N, E, seed, max_deg, dev):
    """Power-law in / out degrees (alpha ~ 2.1, clipped), random edge order, duplicates and self loops allowed
    (SURVEY.md 8(d)); same construction as bench.synth_graph, on the GPU."""
    gen = torch.Generator(device=dev).manual_seed(int(seed))

    def endpoints():
        p = torch.arange(1, N + 1, dtype=torch.float64, device=dev) ** (-1.0 / 1.1)
        p /= p.sum()
        p = torch.clamp(p, max=float(max_deg) / E)
        p /= p.sum()
        cdf = torch.cumsum(p, 0)
        u = torch.rand(E, dtype=torch.float64, device=dev, generator=gen)
        ids = torch.searchsorted(cdf, u).clamp_(max=N - 1)
        perm = torch.randperm(N, device=dev, generator=gen)
        return perm[ids]
    return endpoints(), endpoints()


def ppi_minibatch_device(seed, dev, n_graphs=2, in_features=50, n_labels=121):
    """A minibatch of PPI-shaped graphs (1 000-3 500 nodes each, mean degree 28.7 with both directions),
    block-diagonal, with features and multi-label targets."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1000, 3500, n_graphs)
    gs = []
    for n in sizes:
        e = int(28.7 * n / 2)
        s, d = rng.integers(0, n, e), rng.integers(0, n, e)
        gs.append(sb.Graph(torch.from_numpy(np.concatenate([s, d])), torch.from_numpy(np.concatenate([d, s])), int(n)))
    g = sb.batch(gs).to(dev)
    n = g.number_of_nodes()
    gen = torch.Generator(device=dev).manual_seed(int(seed))
    feat = torch.randn(n, in_features, device=dev, generator=gen)
    y = (torch.rand(n, n_labels, device=dev, generator=gen) < 0.3).float()
    return g, feat, y


# ---- the aggregation work of a layer stack at the C ABI (like bench.Path, any graph / widths) --------------------
class AggStack:
    def __init__(self, dev, graph, widths, S, sample_base, src_scale, dst_scale, normal="boxmuller", bwd_chunk=None):
        self.lib, self.dev, self.S, self.sample_base = _lib.load(), dev, S, sample_base
        st = graph._s
        self.csc, self._k1 = st.csx(True)
        self.csr, self._k2 = st.csx(False)
        self.n_dst, self.n_src, self.E = st.num_nodes, st.num_src, st.num_edges
        self.ss, self.ds = src_scale, dst_scale
        self.logical = list(widths)
        # what stag_b200.ops.stochastic_aggregate does with generator "auto": a width within a third of a multiple of 128
        # runs zero-padded on the tensor-core generator (products' 100 -> 128)
        pad = lambda D: (D + 127) // 128 * 128 if (normal == "hadamard" and D >= 96 and D % 128 and 3 * ((D + 127) // 128 * 128) <= 4 * D) else D  # noqa: E731
        widths = [pad(D) for D in widths]
        self.widths = list(widths)
        gen = torch.Generator(device=dev).manual_seed(4321)
        self.x = [torch.randn((self.n_src, widths[0]), device=dev, generator=gen)]                   # layer-1 input: shared
        self.x += [torch.randn((S, self.n_src, D), device=dev, generator=gen) for D in widths[1:]]   # per-sample activations
        self.out = [torch.empty((S, self.n_dst, D), device=dev) for D in widths]
        self.gout = [torch.randn((S, self.n_dst, D), device=dev, generator=gen) for D in widths]
        # transposed pass in chunks of bwd_chunk samples (the [S, n_src, D] gradient of a shared operand is summed
        # over the samples: chunks bound its size)
        self.chunk = S if bwd_chunk is None else max(1, min(S, bwd_chunk))
        self.dx = [torch.empty((self.chunk, self.n_src, D), device=dev) for D in widths]
        self._dx_sum = torch.zeros((self.n_src, widths[0]), device=dev)
        self.dx_sum = self._dx_sum[:, : self.logical[0]]            # the caller's view: logical width
        for k, (Dl, Dk) in enumerate(zip(self.logical, widths)):    # padded channels carry zeros
            if Dk != Dl:
                self.x[k][..., Dl:] = 0
                self.gout[k][..., Dl:] = 0
        self._xin = None
        self.loc = torch.ones(1, device=dev)
        self.scale = torch.full((1,), 0.4, device=dev)
        self.normal = normal
        wsb = 256
        for D in set(widths):
            wsb = max(wsb, self.lib.stag_spmm_workspace_bytes(ctypes.byref(self.csc), D, S),
                      self.lib.stag_spmm_workspace_bytes(ctypes.byref(self.csr), D, S))
        self.ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.current_stream(dev).cuda_stream
        self.offset = 0

    def _noise(self, layer, D, sample_base):
        n = _lib.StagNoise()
        had = self.normal == "hadamard" and D % 128 == 0
        n.kind, n.K, n.param_shape = (_lib.NOISE_NORMAL_HADAMARD if had else _lib.NOISE_NORMAL), D, _lib.PARAM_SCALAR
        n.relu = n.in_norm = 0
        n.sample_base = sample_base
        n.p0, n.p1, n.external = self.loc.data_ptr(), self.scale.data_ptr(), 0
        n.seed, n.offset = 42, self.offset + layer
        return n

    def fwd(self, layer, x=None):
        D = self.widths[layer]
        x = self.x[layer] if x is None else x
        if x.shape[-1] != D:             # operand handed in at its logical width: into a zero-padded buffer (the operator's pad)
            if self._xin is None or self._xin.shape[:-1] != x.shape[:-1]:
                self._xin = torch.zeros(tuple(x.shape[:-1]) + (D,), device=self.dev)
            self._xin[..., : x.shape[-1]].copy_(x)
            x = self._xin
        shared = x.dim() == 2
        nz = self._noise(layer, D, self.sample_base)
        _lib.check(self.lib.stag_spmm_fwd(
            ctypes.byref(self.csc), x.data_ptr(), D, 0 if shared else self.n_src * D, D, self.S, ctypes.byref(nz),
            self.ss.data_ptr(), self.ds.data_ptr(), self.out[layer].data_ptr(), D, self.n_dst * D, 0,
            self.ws.data_ptr(), self.ws.numel(), self.stream))

    def bwd(self, layer):
        """dX of the layer: the fused kernel on the CSR with the scales swapped, noise regenerated from the edge ids."""
        D = self.widths[layer]
        for s0 in range(0, self.S, self.chunk):
            ns = min(self.chunk, self.S - s0)
            nz = self._noise(layer, D, self.sample_base + s0)
            _lib.check(self.lib.stag_spmm_fwd(
                ctypes.byref(self.csr), self.gout[layer][s0:].data_ptr(), D, self.n_dst * D, D, ns, ctypes.byref(nz),
                self.ds.data_ptr(), self.ss.data_ptr(), self.dx[layer].data_ptr(), D, self.n_src * D, 0,
                self.ws.data_ptr(), self.ws.numel(), self.stream))
            if layer == 0:   # shared operand: its gradient is the sum over the samples
                if s0 == 0:
                    torch.sum(self.dx[0][:ns], 0, out=self._dx_sum)
                else:
                    self._dx_sum += self.dx[0][:ns].sum(0)

    def forward_all(self):
        for layer in range(len(self.widths)):
            self.fwd(layer)

    def backward_all(self):
        for layer in reversed(range(len(self.widths))):
            self.bwd(layer)
        self.offset += len(self.widths)


def _timed(dev, dist, steps, warmup, fn):
    """ms per call of fn: `warmup` untimed calls, barrier, `steps` timed calls between CUDA events, max over ranks."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- (i) strong scaling over the MC samples ---------------------------------------------------------------------
def strong_mc(dev, dist, rank, world, name, src, dst, N, widths, S_total, n_classes, steps, warmup, normal, bwd_chunk=None):
    base, S = P.shard_samples(S_total, rank, world)
    g = sb.Graph(src, dst, N).to(dev)
    st = g._s
    ss, ds = st.scale(False, "rsqrt"), st.scale(True, "rsqrt")
    stack = AggStack(dev, g, widths, S, base, ss, ds, normal=normal, bwd_chunk=bwd_chunk)
    E = st.num_edges
    dims = list(widths) + [n_classes]
    n_params = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(widths)))   # GCN weights + biases
    grads = torch.randn(n_params, device=dev)
    probs = torch.rand(N, n_classes, device=dev)   # this rank's sum over its samples of the [N,C] outputs

    def compute():
        stack.forward_all()
        stack.backward_all()

    def collectives():
        dist.all_reduce(probs)
        dist.all_reduce(grads)

    def step():
        stack.forward_all()
        w = dist.all_reduce(probs, async_op=True)     # MC-mean partial sums travel under the transposed passes
        stack.backward_all()
        w.wait()
        dist.all_reduce(grads)                        # flat bucket, one call (KB-MB: latency-bound)

    t_step = _timed(dev, dist, steps, warmup, step)
    t_comp = _timed(dev, dist, steps, 1, compute)
    t_coll = _timed(dev, dist, steps, 1, collectives)
    es = float(E) * S_total * len(widths)
    return {"workload": name, "scaling": "strong", "samples_total": S_total, "samples_per_gpu": S, "layers": len(widths),
            "ms_per_step": t_step, "compute_ms": t_comp, "collective_ms": t_coll, "exposed_ms": max(0.0, t_step - t_comp),
            "value": es / (t_step * 1e-3) / 1e9, "unit": "GEdge-samples/s",
            "collectives": {"gradient_bucket_floats": n_params, "mc_mean_bytes": int(probs.numel() * 4),
                            "what": "all_reduce([N,C] probability sums) overlapped with the transposed passes + "
                                    "all_reduce(flat gradient bucket), every step"}}


# ---- (ii) minibatch data parallel (C3 PPI) through the module API --------------------------------------------------
def minibatch_dp(dev, dist, rank, world, steps, warmup):
    torch.manual_seed(0)   # identical initial weights on every rank
    mk = lambda i, o, act: sb.layers.StagLayer(sb.zoo.GCN(i, o, activation=act),   # noqa: E731
                                               q_a=torch.distributions.Normal(1.0, 1.0, validate_args=False))
    layers = torch.nn.ModuleList([mk(50, 256, torch.relu), mk(256, 256, torch.relu), mk(256, 121, torch.sigmoid)]).to(dev)
    model = sb.models.StagModel(layers=layers, likelihood=sb.likelihoods.BernoulliLikelihood())
    opt = torch.optim.Adam(model.parameters(), 5e-3)
    sb.random.fold_rank(rank)      # ranks work on different minibatches: independent noise streams
    batches = [ppi_minibatch_device(1000 + 17 * rank + k, dev) for k in range(2)]
    for g, _, _ in batches:
        g._s.csx(True), g._s.csx(False)
    bucket = P.GradBucket(model.parameters())   # gradients are views of one flat buffer: the all-reduce needs no copies
    state = {"k": 0, "bucket": bucket.flat.numel()}

    def step(reduce=True):
        g, feat, y = batches[state["k"] % len(batches)]
        state["k"] += 1
        bucket.zero()
        loss = model.loss(g, feat, y)
        loss.backward()
        if reduce:
            bucket.allreduce()
        opt.step()

    t_step = _timed(dev, dist, steps, warmup, step)
    t_comp = _timed(dev, dist, steps, 1, lambda: step(False))
    flat = torch.zeros(max(state["bucket"], 1), device=dev)
    t_coll = _timed(dev, dist, steps, 1, lambda: dist.all_reduce(flat))
    e_mean = float(np.mean([g.number_of_edges() for g, _, _ in batches]))
    n_mean = float(np.mean([g.number_of_nodes() for g, _, _ in batches]))
    return {"workload": "C3 PPI-shaped minibatches (2 graphs per rank and step), 3-layer stag GCN 50-256-256-121, Bernoulli "
                        "likelihood, Adam: StagModel.loss -> backward -> flat-bucket all-reduce -> step",
            "scaling": "weak", "ms_per_step": t_step, "compute_ms": t_comp, "collective_ms": t_coll,
            "exposed_ms": max(0.0, t_step - t_comp), "minibatches_per_s": world / (t_step * 1e-3),
            "value": e_mean * 3 * world / (t_step * 1e-3) / 1e9, "unit": "GEdge-samples/s",
            "edges_per_minibatch": e_mean, "nodes_per_minibatch": n_mean,
            "collectives": {"gradient_bucket_floats": state["bucket"], "what": "all_reduce(flat gradient bucket), every step"}}


# ---- (iii) one large graph, 1-D row partition with halo exchange (C5 products) -------------------------------------
def row_partition(dev, dist, rank, world, src, dst, N, D, S, steps, warmup, bwd_chunk=8):
    part = P.RowPartition(src, dst, N, rank, world, balance="edges").setup_halo()
    lg = part.local_graph(sb.Graph)
    # GCN degree scalings of the UNPARTITIONED graph: sources in extended order (owned rows, then halo rows), owned destinations
    outdeg = torch.bincount(src, minlength=N).clamp_(min=1).to(torch.float32).pow_(-0.5)
    indeg = torch.bincount(dst, minlength=N).clamp_(min=1).to(torch.float32).pow_(-0.5)
    ext_ids = torch.cat([torch.arange(part.lo, part.hi, device=dev), part.need])
    ss_ext, ds_own = outdeg[ext_ids].contiguous(), indeg[part.lo:part.hi].contiguous()
    del outdeg, indeg
    stack = AggStack(dev, lg, [D], S, 0, ss_ext, ds_own, normal="hadamard", bwd_chunk=bwd_chunk)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    x_block = torch.randn(part.n_own, D, device=dev, generator=gen)
    x_ext = torch.empty(part.n_ext, D, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    acc = {"halo": 0.0, "fwd": 0.0, "bwd": 0.0, "back": 0.0, "n": 0}

    def step(record=False):
        if record:
            ev[0].record()
        part.exchange(x_block, out=x_ext)                 # halo: only the rows this rank's edges reference
        if record:
            ev[1].record()
        stack.fwd(0, x_ext)                               # owned rows, all S samples, noise keyed by GLOBAL edge ids
        if record:
            ev[2].record()
        stack.bwd(0)                                      # partial dX of the extended operand, summed over the samples
        if record:
            ev[3].record()
        dx_own = part.exchange_back(stack.dx_sum)         # halo rows' partials go home and are added there
        if record:
            ev[4].record()
            torch.cuda.synchronize()
            for k, (i, j) in zip(("halo", "fwd", "bwd", "back"), ((0, 1), (1, 2), (2, 3), (3, 4))):
                acc[k] += ev[i].elapsed_time(ev[j])
            acc["n"] += 1
        return dx_own

    t_step = _timed(dev, dist, steps, warmup, step)
    for _ in range(3):
        step(True)
    parts = torch.tensor([acc[k] / acc["n"] for k in ("halo", "fwd", "bwd", "back")], device=dev, dtype=torch.float64)
    dist.all_reduce(parts, op=dist.ReduceOp.MAX)
    halo_ms, fwd_ms, bwd_ms, back_ms = parts.tolist()
    recv_b, send_b = part.halo_bytes(D)
    sizes = torch.tensor([recv_b, send_b, part.n_halo, lg.number_of_edges()], device=dev, dtype=torch.float64)
    dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
    E = int(src.numel())
    return {"workload": "C5 products-shaped graph (N=%d, E=%d, D=%d), ONE graph row-partitioned over %d GPUs, 1 stag GCN "
                        "aggregation forward + transposed pass over %d MC samples (shared X); D = 100 runs zero-padded at 128 on the tensor-core "
                        "generator, as the public operator does" % (N, E, D, world, S),
            "scaling": "strong", "ms_per_step": t_step, "compute_ms": fwd_ms + bwd_ms, "collective_ms": halo_ms + back_ms,
            "exposed_ms": halo_ms + back_ms, "exposed_frac_of_step": (halo_ms + back_ms) / max(t_step, 1e-9),
            "halo_exchange_ms": halo_ms, "forward_ms": fwd_ms, "transposed_ms": bwd_ms, "gradient_return_ms": back_ms,
            "value": float(E) * S / (t_step * 1e-3) / 1e9, "unit": "GEdge-samples/s",
            "collectives": {"halo_rows_received_max": int(sizes[2].item()), "halo_bytes_received_max": int(sizes[0].item()),
                            "halo_bytes_sent_max": int(sizes[1].item()), "full_all_gather_bytes": int((world - 1) * part.per * D * 4),
                            "what": "all_to_all_single of the referenced source rows (forward) and of their partial "
                                    "gradients (backward), once per step: the operand is shared by the %d samples, so one "
                                    "exchange is amortised over all of them" % S}}


def run_all(dev, dist, rank, world, steps, warmup, normal):
    """The `multi_gpu` dict of the bench line.  Every rank calls this; rank 0 prints."""
    import bench
    out = {}
    t0 = time.perf_counter()
    steps = max(3, min(steps, 10))
    warmup = max(3, min(warmup, 5))
    # (i) C2 arxiv, 16 samples / N, 3 layers
    s2, d2 = bench.synth_graph()
    out["c2_arxiv_strong_mc"] = strong_mc(
        dev, dist, rank, world, bench.WORKLOAD + " -- the 16 samples of a step split over the GPUs", torch.from_numpy(s2),
        torch.from_numpy(d2), bench.N_NODES, [bench.WIDTH] * bench.N_LAYERS, bench.N_SAMPLES, 40, steps, warmup, normal)
    torch.cuda.empty_cache()
    # (ii) C3 PPI minibatches
    out["c3_ppi_minibatch_dp"] = minibatch_dp(dev, dist, rank, world, steps * 3, warmup)
    torch.cuda.empty_cache()
    # (iii) + (i) C5 products: one graph, same on every rank
    N5, E5, D5, S5 = 2449029, 61859140, 100, 32
    s5, d5 = powerlaw_graph_device(N5, E5, 0x57A6 + 5, 17000, dev)
    out["c5_products_row_partition"] = row_partition(dev, dist, rank, world, s5, d5, N5, D5, S5, max(3, steps // 2), 3)
    torch.cuda.empty_cache()
    out["c5_products_strong_mc"] = strong_mc(
        dev, dist, rank, world, "C5 products-shaped graph (N=%d, E=%d, D=%d), graph replicated, the 32 MC samples split over "
        "the GPUs, 1 aggregation forward + transposed pass; D = 100 runs zero-padded at 128 on the tensor-core generator" % (N5, E5, D5),
        s5, d5, N5, [D5], S5, 47, max(3, steps // 2), 3,
        "hadamard", bwd_chunk=8)
    out["seconds"] = time.perf_counter() - t0
    return out


# ---- the other BASELINE.json configurations on one GPU (the `configs` dict of the N = 1 bench line) -----------------
def _bytes_per_edge_sample(N, E, D, vi):
    """algorithmic bytes of one layer, forward + transposed pass, per edge-sample."""
    nd = 4.0 * N * D
    fwd = 2 * nd + 4 * E + 4 * (N + 1) + 8 * N
    bwd = (3 if vi else 2) * nd + 8 * E + 4 * (N + 1) + 8 * N
    return (fwd + bwd) / E


def config_row(dev, name, g, widths, S, vi, per_channel, iters, peak, warmup=3, graphed=False):
    """One STEP = every aggregation of the configuration's layers, forward and backward, over S Monte-Carlo samples,
    through the public operator (stag_b200.ops.stochastic_aggregate + autograd); CUDA events, features resident."""
    from stag_b200.ops import NoiseSpec
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

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    es = float(E) * S * len(widths)
    nbytes = sum(_bytes_per_edge_sample(N, E, D, vi) for D in widths) * E * S
    row = {"workload": name, "N": N, "E": E, "widths": list(widths), "mc_samples": S, "mode": "vi" if vi else "mle",
           "ms_per_step": ms, "value": es / ms / 1e6, "unit": "GEdge-samples/s",
           "frac_of_hbm_roof": nbytes / ms / 1e6 / peak, "launches_per_step": 2 * len(widths)}
    if graphed:
        # launch-bound shapes: the same step captured ONCE in a CUDA graph and replayed; the device-side call counter
        # (stag_b200.random.enable_device_counter) gives every replay fresh noise (tests/test_gpu_graphs.py)
        try:
            ctr = sb.random.enable_device_counter(dev)
            ctr.zero_()

            def gstep():
                step()
                sb.random.advance_device_counter()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    gstep()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                gstep()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                graph.replay()
            b.record()
            torch.cuda.synchronize()
            gms = a.elapsed_time(b) / iters
            row["cuda_graph"] = {"ms_per_step": gms, "value": es / gms / 1e6, "frac_of_hbm_roof": nbytes / gms / 1e6 / peak,
                                 "what": "the same step replayed from one captured CUDA graph, fresh noise per replay "
                                         "(device-side Philox call counter)"}
        except Exception as exc:   # capture problems must not take the bench line down
            row["cuda_graph"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        finally:
            sb.random.disable_device_counter()
    return row


def _batched_graphs(sizes, edges_per_node, seed):
    rng = np.random.default_rng(seed)
    gs = []
    for n in sizes:
        e = max(1, int(edges_per_node * n / 2))
        s, d = rng.integers(0, n, e), rng.integers(0, n, e)
        gs.append(sb.Graph(torch.from_numpy(np.concatenate([s, d])), torch.from_numpy(np.concatenate([d, s])), int(n)))
    return sb.batch(gs)


def configs_single_gpu(dev, peak):
    """C1, C3, C4, C5 and the vi mode of C2 (SURVEY section 8 shapes), same timing method as the headline."""
    import bench
    rows = {}
    rng = np.random.default_rng(0)
    s, d = powerlaw_graph_device(2708, 10556, 0x57A6 + 1, 168, dev)
    rows["c1_cora"] = config_row(dev, "C1 Cora-shaped (N 2 708, E 10 556), 3 layers 1433-16-16, S 4", sb.Graph(s, d, 2708),
                                 [1433, 16, 16], 4, False, False, 50, peak, graphed=True)
    s2, d2 = bench.synth_graph()
    g2 = sb.Graph(torch.from_numpy(s2), torch.from_numpy(d2), bench.N_NODES)
    rows["c2_arxiv_vi_rc"] = config_row(dev, "C2 arxiv-shaped, 3 layers, learned per-channel Normal noise (vi=True: d loc / d scale)",
                                        g2, [128, 128, 128], 16, True, True, 5, peak)
    del g2
    torch.cuda.empty_cache()
    sizes = rng.integers(1000, 3500, 2)
    rows["c3_ppi_minibatch"] = config_row(dev, "C3 PPI-shaped minibatch (2 graphs), 3 layers 50-256-256, S 1 (training step)",
                                          _batched_graphs(sizes, 28.7, 3), [50, 256, 256], 1, False, False, 50, peak, graphed=True)
    sizes = rng.integers(1000, 3500, 24)
    rows["c3_ppi_all_graphs"] = config_row(dev, "C3 PPI-shaped, all 24 graphs in one batch, S 4 (inference)",
                                           _batched_graphs(sizes, 28.7, 3), [50, 256, 256], 4, False, False, 20, peak)
    sizes = np.clip(rng.normal(25.5, 12, 32), 2, 80).astype(int)
    rows["c4_molhiv_batch32"] = config_row(dev, "C4 molhiv-shaped batch of 32 molecules, 2 layers 9-16, per-channel learned Normal (vi), S 4",
                                           _batched_graphs(sizes, 2.15, 4), [9, 16], 4, True, True, 50, peak, graphed=True)
    sizes = np.clip(rng.normal(25.5, 12, 4096), 2, 80).astype(int)
    rows["c4_molhiv_batch4096"] = config_row(dev, "C4 molhiv-shaped batch of 4 096 molecules, 2 layers 9-256, vi, S 4",
                                             _batched_graphs(sizes, 2.15, 4), [9, 256], 4, True, True, 20, peak)
    s5, d5 = powerlaw_graph_device(2449029, 61859140, 0x57A6 + 5, 17000, dev)
    rows["c5_products_4_of_32_samples"] = config_row(dev, "C5 products-shaped (N 2 449 029, E 61 859 140, D 100), 1 layer, 4 of the 32 samples "
                                                     "(one GPU's share at 8 GPUs)", sb.Graph(s5, d5, 2449029), [100], 4, False, False, 3, peak, warmup=2)
    torch.cuda.empty_cache()
    return rows


def unfused_gpu_baseline(dev, src, dst, N, D, sigma=0.4, iters=5):
    """the reference's algorithm on the SAME GPU, un-fused, in plain torch ops -- noise tensor [E,D]
    materialised by torch.normal, message = x[src] * w, index_add into the destinations, autograd backward (what
    stag/layers.py:115-129 + DGL's gspmm / gsddmm do, one layer x one MC sample per step).  This is NOT this library."""
    src_t, dst_t = torch.as_tensor(src, device=dev), torch.as_tensor(dst, device=dev)
    E = src_t.numel()
    outdeg = torch.bincount(src_t, minlength=N).clamp_(min=1).float().pow_(-0.5)
    indeg = torch.bincount(dst_t, minlength=N).clamp_(min=1).float().pow_(-0.5)
    x = torch.randn(N, D, device=dev, requires_grad=True)
    gout = torch.randn(N, D, device=dev)

    def step():
        w = 1.0 + sigma * torch.randn(E, D, device=dev)
        msg = (x * outdeg[:, None])[src_t] * w
        out = torch.zeros(N, D, device=dev).index_add_(0, dst_t, msg) * indeg[:, None]
        out.backward(gout)
        x.grad = None

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    return {"value": E / ms / 1e6, "unit": "GEdge-samples/s", "ms_per_layer_sample": ms,
            "what": "the reference's algorithm un-fused in plain torch CUDA ops on this GPU (torch.randn [E,D] noise, gather, "
                    "multiply, index_add, autograd backward), 1 layer x 1 MC sample per step; not DGL's kernels (DGL is not "
                    "installable here) and not this library"}
