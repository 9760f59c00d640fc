"""GPU: the other BASELINE.json configurations as parity cases (shapes from SURVEY.md section 8):
C1 Cora-shaped (N 2 708, E 10 556, widths 1433 -> 16 -> 7, S 4), C3 PPI-shaped minibatch of 2 graphs
(D 50 / 256, 121 labels, Bernoulli likelihood), C4 molhiv-shaped batch of 32 small graphs (per-channel
learned Normal, vi=True, sum readout), C5 products-shaped large graph (N 2.4 M, E 62 M, D 100) through
size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import ref_spmm

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def oracle_gcn_model(src, dst, N, feat, Ws, bs, noises, acts):
    h = feat
    for W, b, w, act in zip(Ws, bs, noises, acts):
        h = ref_spmm.gcn_forward(src, dst, N, h, w, W, b, "both", act)
    return h


def test_c1_cora_shaped_model_shared_noise():
    """3-layer GCN 1433 -> 16 -> 16 -> 7 (scripts/citation_mle/run.sh:7), external noise per layer."""
    import stag_b200 as stag
    rng = np.random.default_rng(1)
    N, E = 2708, 10556
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    widths = [1433, 16, 16, 7]
    x = (rng.random((N, 1433)) < 0.012).astype(np.float32)
    x /= np.maximum(x.sum(1, keepdims=True), 1)          # row-normalised bag of words
    acts = [torch.relu, torch.relu, lambda t: torch.softmax(t, dim=-1)]
    layers = torch.nn.ModuleList([stag.layers.StagLayer(stag.zoo.GCN(widths[i], widths[i + 1], activation=acts[i]),
                                                        q_a=torch.distributions.Normal(1.0, 0.4)) for i in range(3)]).cuda()
    noises = [T((1 + 0.4 * rng.standard_normal((E, widths[i]))).astype(np.float32)) for i in range(3)]
    for layer, w in zip(layers, noises):
        layer.rsample_noise = lambda graph, sample_dimension, w=w: w.cuda()
    g = stag.Graph(T(src), T(dst), N).to("cuda")
    model = stag.models.StagModel(layers)
    out = model._forward(g, T(x).cuda())
    ref = oracle_gcn_model(T(src), T(dst), N, T(x), [l.base_layer.weight.detach().cpu() for l in layers],
                           [l.base_layer.bias.detach().cpu() for l in layers], noises, acts)
    assert rel(out, ref) < 1e-5
    # and the fused path: 4 MC samples, probabilities stay normalised
    for layer in layers:
        del layer.rsample_noise
    p = model.forward(g, T(x).cuda(), n_samples=4, return_parameters=True)
    assert p.shape == (N, 7) and torch.allclose(p.sum(-1), torch.ones(N, device="cuda"), atol=1e-5)


@pytest.mark.parametrize("D", [50, 256])
def test_c3_ppi_shaped_minibatch(D):
    """Two PPI-sized graphs batched block-diagonally (scripts/ppi_mle/run.py:70-77), GCN D -> 121 with a
    sigmoid, Bernoulli likelihood: loss and gradients against the oracle under shared noise."""
    import stag_b200 as stag
    rng = np.random.default_rng(D)
    sizes, gs, srcs, dsts, off = [2245, 2481], [], [], [], 0
    for n in sizes:
        e = int(n * 14.4)
        s, d = rng.integers(0, n, e), rng.integers(0, n, e)
        gs.append(stag.Graph(T(np.concatenate([s, d])), T(np.concatenate([d, s])), n))   # both directions
        srcs.append(np.concatenate([s, d]) + off)
        dsts.append(np.concatenate([d, s]) + off)
        off += n
    bg = stag.batch(gs).to("cuda")
    src, dst, N = np.concatenate(srcs), np.concatenate(dsts), off
    assert np.array_equal(bg.edges()[0].cpu().numpy(), src) and bg.batch_num_nodes().tolist() == sizes
    E = len(src)
    x = rng.standard_normal((N, D)).astype(np.float32)
    y = (rng.random((N, 121)) < 0.3).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal((E, D))).astype(np.float32)
    layer = stag.layers.StagLayer(stag.zoo.GCN(D, 121, activation=torch.sigmoid),
                                  q_a=torch.distributions.Normal(1.0, 0.2)).cuda()
    layer.rsample_noise = lambda graph, sample_dimension: T(w).cuda()
    model = stag.models.StagModel(torch.nn.ModuleList([layer]), likelihood=stag.likelihoods.BernoulliLikelihood())
    xc = T(x).cuda().requires_grad_(True)
    loss = model.loss(bg, xc, T(y).cuda())
    loss.backward()
    xo = T(x).requires_grad_(True)
    Wo = layer.base_layer.weight.detach().cpu().requires_grad_(True)
    bo = layer.base_layer.bias.detach().cpu().requires_grad_(True)
    h = ref_spmm.gcn_forward(T(src), T(dst), N, xo, T(w), Wo, bo, "both", torch.sigmoid)
    lo = -torch.distributions.Bernoulli(probs=h).log_prob(T(y)).mean()
    lo.backward()
    assert rel(loss, lo) < 1e-5
    assert rel(xc.grad, xo.grad) < 2e-5
    assert rel(layer.base_layer.weight.grad, Wo.grad) < 2e-5 and rel(layer.base_layer.bias.grad, bo.grad) < 2e-5


def test_c4_molhiv_shaped_batch_vi_readout():
    """32 molecule-sized graphs, GCN 9 -> 16 -> 16 with per-channel learned Normal noise (vi=True), sum
    readout (scripts/molhiv_rec/run.py): fused forward/backward == the emitted-noise path of the same call."""
    import stag_b200 as stag
    rng = np.random.default_rng(4)
    gs = []
    for _ in range(32):
        n = int(np.clip(rng.normal(25.5, 12), 2, 80))
        e = max(1, int(1.1 * n))
        s, d = rng.integers(0, n, e), rng.integers(0, n, e)
        gs.append(stag.Graph(T(np.concatenate([s, d])), T(np.concatenate([d, s])), n))
    bg = stag.batch(gs).to("cuda")
    N = bg.number_of_nodes()
    mk = lambda d: torch.distributions.Normal(torch.ones(d), 0.3 * torch.ones(d))  # noqa: E731
    layers = torch.nn.ModuleList([
        stag.layers.StagLayer(stag.zoo.GCN(9, 16, activation=torch.relu), q_a=mk(9), p_a=mk(9), vi=True),
        stag.layers.StagLayer(stag.zoo.GCN(16, 16), q_a=mk(16), p_a=mk(16), vi=True),
        stag.layers.SumNodes(),
    ]).cuda()
    x = T(rng.standard_normal((N, 9)).astype(np.float32)).cuda()

    def run(fused):
        stag.manual_seed(77)
        for p in layers.parameters():
            p.grad = None
        h = x
        for i, layer in enumerate(layers):
            if isinstance(layer, stag.layers.StagLayer) and not fused:
                spec = layer.noise_spec(bg, h.shape[-1])
                spec.seed, spec.offset = 77, i
                h = layer.base_layer(bg, h, edge_weight=spec.materialize())
            else:
                h = layer(bg, h)
        kl = sum(layer.kl_divergence() for layer in layers if layer.vi)
        (h.pow(2).mean() + 0.1 * kl).backward()
        return h.detach().clone(), {k: p.grad.clone() for k, p in layers.named_parameters()}
    h1, g1 = run(True)
    h2, g2 = run(False)
    # ... and against the ORACLE at this shape: the first layer's aggregation with the noise of the numpy
    # restatement of the generator (not the emitted tensor), summed by the torch restatement of update_all
    from oracle import ref_philox, ref_spmm
    src_b, dst_b = bg.edges()
    loc = layers[0].q_a.loc.detach()
    scale = layers[0].q_a.log_scale.detach().exp()
    spec0 = stag.ops.NoiseSpec("normal", loc, scale, 9, bg.number_of_edges(), seed=5, offset=2)
    ss, ds = bg._s.scale(False, "rsqrt"), bg._s.scale(True, "rsqrt")
    got = stag.ops.stochastic_aggregate(bg, x, spec0, src_scale=ss, dst_scale=ds)
    w = ref_philox.noise("normal", bg.number_of_edges(), 9, 0, 5, 2, loc.cpu().numpy(), scale.cpu().numpy())
    exp = ref_spmm.aggregate(src_b.cpu(), dst_b.cpu(), N, x.cpu().double(), T(w).double(),
                             src_scale=ss.cpu().double(), dst_scale=ds.cpu().double())
    assert float((got.cpu().double() - exp).abs().max()) <= 1e-5 * float(exp.abs().max())
    assert bool(((got.cpu().double() - exp).abs() <= 1e-5 * exp.abs() + 1e-6 * exp.abs().max()).all())
    assert h1.shape == (32, 16)
    assert rel(h1, h2) < 1e-5
    for k in g1:
        assert rel(g1[k], g2[k]) < 1e-4, k


def test_c5_products_shaped_properties():
    """N 2 449 029, E 61 859 140, D 100: structure invariants, adjointness of the fused forward / transposed
    pass, and the 1-D row partition reproducing the unpartitioned rows with the same noise."""
    import stag_b200 as sb
    from stag_b200 import parallel as P
    N, E, D = 2449029, 61859140, 100
    rng = np.random.default_rng(5)
    p = np.arange(1, N + 1, dtype=np.float64) ** -0.6
    cdf = np.cumsum(p / p.sum())
    perm = rng.permutation(N)
    dst = perm[np.searchsorted(cdf, rng.random(E))].astype(np.int64)
    src = rng.integers(0, N, E).astype(np.int64)
    g = sb.Graph(T(src), T(dst), N).to("cuda")
    indptr, indices, eid = g.adj_tensors("csc")
    assert int(indptr[-1]) == E and bool((indptr[1:] >= indptr[:-1]).all())
    deg = torch.bincount(T(dst).cuda(), minlength=N)
    assert torch.equal((indptr[1:] - indptr[:-1]).long(), deg)
    assert torch.equal(torch.sort(eid.long()).values, torch.arange(E, device="cuda"))
    chk = torch.randint(0, E, (100000,), device="cuda")
    assert torch.equal(T(src).cuda()[eid.long()[chk]], indices.long()[chk])
    from stag_b200.ops import NoiseSpec
    one, sg = torch.ones((), device="cuda"), torch.full((), 0.4, device="cuda")
    spec = lambda e, **kw: NoiseSpec("normal", one, sg, D, e, seed=3, offset=9, **kw)  # noqa: E731
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, D, device="cuda", generator=gen).requires_grad_(True)
    y = torch.randn(1, N, D, device="cuda", generator=gen)
    out = sb.ops.stochastic_aggregate(g, x, spec(E), reduce="mean", n_samples=1)
    lhs = (out.double() * y.double()).sum()
    out.backward(y)
    rhs = (x.detach().double() * x.grad.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * float(out.detach().double().abs().mul(y.double().abs()).sum())
    # ---- the fused kernels against the ORACLE at this shape: 3 000 random rows, their in-edges from the
    # CSC, the noise of exactly those edges from the numpy restatement of the generator (oracle/ref_philox.py), the
    # reference's sum in edge-id order in float64 (stag/zoo/gcn.py:63,94-96 with fn.mean)
    from oracle import ref_philox
    from stag_b200 import _lib
    full = sb.ops.stochastic_aggregate(g, x.detach(), spec(E), reduce="mean", n_samples=1)[0]
    rows = torch.from_numpy(np.random.default_rng(1).choice(N, 3000, replace=False)).cuda()

    def oracle_rows(indptr, indices, eid, rows, operand, row_scale):
        a, b = indptr[rows].long(), indptr[rows + 1].long()
        cnt = (b - a).cpu().numpy()
        pos = torch.cat([torch.arange(int(lo_), int(hi_), device="cuda") for lo_, hi_ in zip(a.tolist(), b.tolist())])
        ids, nbr = eid.long()[pos].cpu().numpy(), indices.long()[pos]
        if spec(E).lib_kind == _lib.NOISE_NORMAL_HADAMARD:   # D = 100 runs padded to one 128-channel group on the tensor cores
            w = ref_philox.noise("normal_hadamard", ids, spec(E).hadamard_width, 0, 3, 9, 1.0, 0.4)[:, :D].astype(np.float64)
        else:
            w = ref_philox.noise("normal", ids, D, 0, 3, 9, 1.0, 0.4).astype(np.float64)
        msg = w * operand[nbr].double().cpu().numpy()
        seg = np.repeat(np.arange(len(cnt)), cnt)
        exp = np.zeros((len(cnt), D))
        np.add.at(exp, seg, msg)                       # stored-edge order = edge-id order within a row
        return exp * row_scale[:, None]
    inv_in = 1.0 / np.maximum((indptr[rows + 1] - indptr[rows]).cpu().numpy(), 1)
    exp = oracle_rows(indptr, indices, eid, rows, x.detach(), inv_in)
    got = full[rows].double().cpu().numpy()
    assert np.all(np.abs(got - exp) <= 1e-5 * np.abs(exp) + 1e-6 * np.abs(exp).max())
    # transposed pass: dX[u] = sum over out-edges of w * (dOut[v] / clamp(indeg(v), 1)), 3 000 random source rows
    ptr_r, idx_r, eid_r = g.adj_tensors("csr")
    dscale = 1.0 / torch.clamp((indptr[1:] - indptr[:-1]).float(), min=1)
    expd = oracle_rows(ptr_r, idx_r, eid_r, rows, y[0] * dscale[:, None], np.ones(len(rows)))
    gotd = x.grad[rows].double().cpu().numpy()
    assert np.all(np.abs(gotd - expd) <= 1e-5 * np.abs(expd) + 1e-6 * np.abs(expd).max())
    del full
    # ---- 1-D row partition, rank 3 of 8: whole-block form and halo form reproduce the unpartitioned rows bitwise ----
    ref = sb.ops.stochastic_aggregate(g, x.detach(), spec(E), n_samples=1)[0]
    part = P.RowPartition(T(src), T(dst), N, 3, 8, halo=False)
    lg = part.local_graph(sb.Graph).to("cuda")
    lo = sb.ops.stochastic_aggregate(lg, x.detach(), spec(lg.number_of_edges()), n_samples=1)[0]
    assert torch.equal(lo[part.lo:part.hi], ref[part.lo:part.hi])
    del lo, lg
    hp = P.RowPartition(T(src).cuda(), T(dst).cuda(), N, 3, 8)     # halo plan; the exchange itself is emulated by indexing
    hg = hp.local_graph(sb.Graph)
    assert hg.number_of_nodes() == hp.n_own and hg.number_of_src_nodes() == hp.n_ext
    x_ext = torch.cat([x.detach()[hp.lo:hp.hi], x.detach()[hp.need]])
    lo = sb.ops.stochastic_aggregate(hg, x_ext, spec(hg.number_of_edges()), n_samples=1)[0]
    assert lo.shape == (hp.n_own, D) and torch.equal(lo, ref[hp.lo:hp.hi])


def test_c2_arxiv_shaped_training_step():
    """The BASELINE configuration itself as a program: 3-layer stag GCN 128 -> 128 -> 128 -> 40 with
    BatchNorm / ReLU / Dropout feature layers (scripts/arxiv_mle/gcn/run.py:76-131), Normal(1, 0.4) edge noise,
    16 MC samples batched through every layer, loss.backward(), Adam step; then MC inference."""
    import bench
    import stag_b200 as stag
    src, dst = bench.synth_graph()
    N = bench.N_NODES
    g = stag.Graph(T(src), T(dst), N).to("cuda")
    q_a = torch.distributions.Normal(1.0, 0.4)
    feat_layer = lambda: stag.layers.FeatOnlyLayer(torch.nn.Sequential(  # noqa: E731
        torch.nn.BatchNorm1d(128), torch.nn.ReLU(), torch.nn.Dropout(0.5)))
    layers = torch.nn.ModuleList([
        stag.layers.StagLayer(stag.zoo.GCN(128, 128), q_a=q_a), feat_layer(),
        stag.layers.StagLayer(stag.zoo.GCN(128, 128), q_a=q_a), feat_layer(),
        stag.layers.StagLayer(stag.zoo.GCN(128, 40, activation=lambda x: torch.softmax(x, dim=-1)), q_a=q_a),
    ]).cuda()
    model = stag.models.StagModel(layers)
    opt = torch.optim.Adam(layers.parameters(), 1e-2)
    gen = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(N, 128, device="cuda", generator=gen)
    y = torch.randint(0, 40, (N,), device="cuda", generator=gen)
    mask = torch.rand(N, device="cuda", generator=gen) < 0.5
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = model.loss(g, x, y, mask=mask, n_samples=16)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    for p in layers.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()
    layers.eval()
    with torch.no_grad():
        probs = model.forward(g, x, n_samples=16, return_parameters=True)
    assert probs.shape == (N, 40) and torch.allclose(probs.sum(-1), torch.ones(N, device="cuda"), atol=1e-4)
