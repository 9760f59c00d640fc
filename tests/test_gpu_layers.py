"""GPU: module-level drop-in behaviour (StagLayer / StagModel / distributions) against the
reference-made golden vectors, plus the reference's own shape tests (stag/tests/test_layers.py)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().cpu().double() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


# --- the reference's shape tests, on the fused path ------------------------------------------------
def test_forward_r1():
    import stag_b200 as stag
    layer = stag.layers.StagLayer(stag.zoo.GCN(16, 32)).cuda()
    g = stag.rand_graph(3, 9).to("cuda")
    h = layer(g, torch.randn(3, 16).cuda())
    assert h.shape == torch.Size([3, 32]) and torch.isfinite(h).all()
    assert layer._edge_weight_sample.shape == (9, 16)


def test_forward_rc():
    import stag_b200 as stag
    q_a = torch.distributions.Normal(torch.ones(16), torch.ones(16))
    layer = stag.layers.StagLayer(stag.zoo.GCN(16, 32), q_a=q_a).cuda()
    g = stag.rand_graph(3, 9).to("cuda")
    assert layer(g, torch.randn(3, 16).cuda()).shape == torch.Size([3, 32])


@pytest.mark.parametrize("outf", [1, 16])
def test_forward_re_rec(outf):
    import stag_b200 as stag
    q_a = stag.distributions.AmortizedDistribution(16, outf)
    layer = stag.layers.StagLayer(stag.zoo.GCN(16, 32), q_a=q_a).cuda()
    g = stag.rand_graph(3, 9).to("cuda")
    assert layer(g, torch.randn(3, 16).cuda()).shape == torch.Size([3, 32])


def test_forward_gat_and_others():
    import stag_b200 as stag
    g = stag.add_self_loop(stag.rand_graph(30, 90)).to("cuda")
    h = torch.randn(30, 16).cuda()
    assert stag.layers.StagLayer(stag.zoo.GAT(16, 8, num_heads=4)).cuda()(g, h).shape == (30, 32)
    assert stag.layers.StagLayer(stag.zoo.GAT(16, 8, num_heads=4, last=True)).cuda()(g, h).shape == (30, 8)
    assert stag.layers.StagLayer(stag.zoo.GraphSAGE(16, 8)).cuda()(g, h).shape == (30, 8)
    assert stag.layers.StagLayer(stag.zoo.GIN(16, 8)).cuda()(g, h).shape == (30, 8)
    # the reference's GatedGCN adds A(h) [N,out] to the aggregate of raw h [N,in] (stag/zoo/gated_gcn.py:30-43):
    # with edge weights it only works for input_dim == output_dim
    assert stag.layers.StagLayer(stag.zoo.GatedGCN(16, 16)).cuda()(g, h).shape == (30, 16)


def test_fused_layer_equals_tensor_path_on_its_own_sample():
    """The fused forward (noise never stored) equals base_layer.forward on the materialised
    _edge_weight_sample of the same call (reference attribute, stag/layers.py:107)."""
    import stag_b200 as stag
    for kw in ({}, {"relu": True}, {"q_a": torch.distributions.Bernoulli(probs=0.7), "norm": True},
               {"q_a": torch.distributions.Uniform(0.5, 1.5)}):
        layer = stag.layers.StagLayer(stag.zoo.GCN(20, 8), **kw).cuda()
        g = stag.rand_graph(100, 900).to("cuda")
        h = torch.randn(100, 20).cuda()
        out = layer(g, h)
        w = layer._edge_weight_sample
        assert w.shape == (900, 20)
        ref = layer.base_layer(g, h, edge_weight=w)
        assert rel(out, ref) < 1e-5


# --- golden: model level -----------------------------------------------------------------------
def _build_model_rc_vi(stag, d, **kw):
    D0, H = d["p_0__base_layer__weight"].shape
    C = d["p_1__base_layer__weight"].shape[1]
    mk = lambda n: torch.distributions.Normal(torch.ones(n), torch.ones(n))  # noqa: E731
    layers = torch.nn.ModuleList([
        stag.layers.StagLayer(stag.zoo.GCN(D0, H, activation=torch.relu), q_a=mk(D0), p_a=mk(D0), vi=True),
        stag.layers.StagLayer(stag.zoo.GCN(H, C, activation=lambda x: torch.softmax(x, dim=-1)), q_a=mk(H),
                              p_a=mk(H), vi=True)])
    sd = {k[2:].replace("__", "."): torch.from_numpy(d[k]) for k in d.files if k.startswith("p_")}
    layers.load_state_dict(sd)   # state_dict keys identical to the reference's
    return stag.models.StagModel(layers.cuda(), kl_scaling=float(d["kl_scaling"]), **kw), layers


def test_model_loss_and_grads_match_reference_golden():
    import stag_b200 as stag
    d = golden("model_rc_vi")
    model, layers = _build_model_rc_vi(stag, d, batch_samples=False)
    eps = [torch.from_numpy(d["eps0"]).cuda(), torch.from_numpy(d["eps1"]).cuda()]
    counters = [0, 0]
    for i, layer in enumerate(layers):
        def rsample(graph, sample_dimension, i=i, layer=layer):
            base = layer.q_a.base_distribution
            w = base.loc + eps[i][counters[i]] * base.scale
            counters[i] += 1
            return w
        layer.rsample_noise = rsample
    g = stag.Graph(torch.from_numpy(d["src"]), torch.from_numpy(d["dst"]), int(d["num_nodes"])).to("cuda")
    nll, reg = model.loss_terms(g, torch.from_numpy(d["feat"]).cuda(), torch.from_numpy(d["y"]).cuda(),
                                mask=torch.from_numpy(d["mask"]).cuda(), n_samples=eps[0].shape[0])
    (nll + reg).backward()
    assert rel(nll, d["nll"]) < 1e-5 and rel(reg, d["reg"]) < 1e-5
    for k, p in layers.named_parameters():
        assert rel(p.grad, d["g_" + k.replace(".", "__")]) < 1e-4, k


def test_model_fused_batched_equals_sequential():
    """S samples in one launch per layer (fused Philox, sample index = s) == S sequential passes
    consuming the same Philox stream through the emitted-noise path."""
    import stag_b200 as stag
    d = golden("model_rc_vi")
    g = stag.Graph(torch.from_numpy(d["src"]), torch.from_numpy(d["dst"]), int(d["num_nodes"])).to("cuda")
    feat, y = torch.from_numpy(d["feat"]).cuda(), torch.from_numpy(d["y"]).cuda()
    S = 4
    model, layers = _build_model_rc_vi(stag, d, batch_samples=True)
    stag.manual_seed(123)
    loss = model.loss(g, feat, y, n_samples=S)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in layers.named_parameters()}
    # sequential replay: layer i, sample s uses (seed, offset=i, sample=s)
    for p in layers.parameters():
        p.grad = None
    total = 0.0
    for s in range(S):
        h = feat
        for i, layer in enumerate(layers):
            spec = layer.noise_spec(g, h.shape[-1])
            spec.seed, spec.offset, spec.sample_base = 123, i, s
            w = spec.materialize()
            h = layer.base_layer(g, h, edge_weight=w)
        total = total - model.likelihood.log_prob(h, y).mean()
    reg = sum(layer.kl_divergence() for layer in layers)
    loss2 = total / S + reg * model.kl_scaling
    loss2.backward()
    assert rel(loss, loss2) < 1e-5
    for k, p in layers.named_parameters():
        assert rel(grads[k], p.grad) < 1e-4, k


@pytest.mark.parametrize("tag", ["re", "rec"])
def test_amortized_matches_reference_golden(tag):
    import stag_b200 as stag
    d = golden("amortized_" + tag)
    outf = d["p_q_a__parameters_mlp__loc__weight"].shape[0]
    q_a = stag.distributions.AmortizedDistribution(16, outf)
    layer = stag.layers.StagLayer(stag.zoo.GCN(16, 32), q_a=q_a, p_a=torch.distributions.Normal(1.0, 1.0), vi=True)
    layer.load_state_dict({k[2:].replace("__", "."): torch.from_numpy(d[k]) for k in d.files if k.startswith("p_")})
    layer = layer.cuda()
    eps = torch.from_numpy(d["eps"]).cuda()

    def rsample(graph, sample_dimension):
        dist = layer.q_a.expand([graph.number_of_edges(), sample_dimension])
        return dist.loc + eps * dist.scale
    layer.rsample_noise = rsample
    g = stag.Graph(torch.from_numpy(d["src"]), torch.from_numpy(d["dst"]), int(d["num_nodes"])).to("cuda")
    feat = torch.from_numpy(d["feat"]).cuda().requires_grad_(True)
    out = layer(g, feat)
    kl = layer.kl_divergence()
    ((out * torch.from_numpy(d["gout"]).cuda()).sum() + kl).backward()
    assert rel(out, d["out"]) < 1e-5 and rel(kl, d["kl"]) < 1e-5
    assert rel(feat.grad, d["dfeat"]) < 1e-5
    for k, p in layer.named_parameters():
        key = "g_" + k.replace(".", "__")
        if key in d:
            assert rel(p.grad, d[key]) < 1e-4, k


def test_amortized_fused_param_grads_flow():
    """Fused path with per-edge parameters: gradients reach the edge MLP."""
    import stag_b200 as stag
    q_a = stag.distributions.AmortizedDistribution(16, 16)
    layer = stag.layers.StagLayer(stag.zoo.GCN(16, 8), q_a=q_a, vi=True).cuda()
    g = stag.rand_graph(50, 400).to("cuda")
    out = layer(g, torch.randn(50, 16).cuda())
    (out.sum() + layer.kl_divergence()).backward()
    for k, p in layer.named_parameters():
        if k.startswith("q_a"):
            assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0, k


def test_mc_forward_mean_and_determinism():
    import stag_b200 as stag
    d = golden("model_rc_vi")
    model, layers = _build_model_rc_vi(stag, d)
    g = stag.Graph(torch.from_numpy(d["src"]), torch.from_numpy(d["dst"]), int(d["num_nodes"])).to("cuda")
    feat = torch.from_numpy(d["feat"]).cuda()
    stag.manual_seed(7)
    a = model.forward(g, feat, n_samples=8, return_parameters=True)
    stag.manual_seed(7)
    b = model.forward(g, feat, n_samples=8, return_parameters=True)
    assert torch.equal(a, b)
    assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=1e-5)
    yhat = model.forward(g, feat, n_samples=2)
    assert yhat.shape == (int(d["num_nodes"]),)


def test_kl_fallback_with_mixture_prior():
    """No analytic KL(Normal || MixtureSameFamily) is registered: the reference falls back to
    log q(w) - log p(w) on the stored sample (stag/layers.py:141-143,
    scripts/citation_rec_contrastive/gcn/run.py:44-52).  Here the sample is materialised from the same
    Philox stream the fused forward consumed."""
    import stag_b200 as stag
    D = 12
    mix = torch.distributions.MixtureSameFamily(
        torch.distributions.Categorical(torch.tensor([0.5, 0.5]).cuda()),
        torch.distributions.Normal(torch.tensor([0.0, 1.0]).cuda(), torch.tensor([0.3, 0.3]).cuda()))
    q = torch.distributions.Normal(torch.ones(D), 0.2 * torch.ones(D))
    layer = stag.layers.StagLayer(stag.zoo.GCN(D, 4), q_a=q, p_a=mix, vi=True).cuda()
    g = stag.rand_graph(40, 300).to("cuda")
    x = torch.randn(40, D).cuda()
    out = layer(g, x)
    kl = layer.kl_divergence()
    w = layer._edge_weight_sample
    assert w.shape == (300, D)
    want = layer.q_a.log_prob(w).sum(-1).mean() - mix.log_prob(w).sum(-1).mean()
    assert rel(kl, want) < 1e-6
    assert rel(out, layer.base_layer(g, x, edge_weight=w)) < 1e-5      # the sample IS the one the kernel used
    (out.sum() + kl).backward()
    assert layer.q_a.loc.grad is not None and torch.isfinite(layer.q_a.log_scale.grad).all()


@pytest.mark.parametrize("H", [1, 3, 4, 8, 32])
def test_edge_softmax_matches_the_scatter_formulation(H):
    """stag_edge_softmax (forward + backward) == dgl.nn.edge_softmax restated with scatter ops (oracle/dgl_shim),
    on a graph with a hub destination, duplicate edges and nodes without in-edges."""
    import stag_b200 as stag
    g0 = torch.Generator().manual_seed(H)
    N, E = 300, 4000
    src, dst = torch.randint(0, N, (E,), generator=g0), torch.randint(5, N, (E,), generator=g0)
    dst[:900] = 17
    g = stag.Graph(src, dst, N).to("cuda")
    logits = (3 * torch.randn(E, H, generator=g0)).cuda()
    a = logits.clone().requires_grad_(True)
    b = logits.clone().double().requires_grad_(True)
    out = stag.ops.edge_softmax(g, a)
    d = dst.cuda()
    idx = d.unsqueeze(-1).expand(E, H)
    mx = torch.full((N, H), float("-inf"), dtype=torch.float64, device="cuda").scatter_reduce(0, idx, b, reduce="amax")
    ex = torch.exp(b - mx[d])
    ref = ex / torch.zeros((N, H), dtype=torch.float64, device="cuda").index_add(0, d, ex)[d]
    assert rel(out, ref) < 1e-6
    sums = torch.zeros((N, H), device="cuda").index_add(0, d, out)
    assert torch.allclose(sums[d.unique()], torch.ones_like(sums[d.unique()]), atol=1e-5)
    gout = torch.randn(E, H, generator=g0).cuda()
    out.backward(gout)
    ref.backward(gout.double())
    assert rel(a.grad, b.grad) < 1e-5
    assert stag.ops.edge_softmax(g, logits[:, 0]).shape == (E,)


def test_sample_shards_draw_their_own_sample_indices():
    """Monte-Carlo samples sharded over ranks (emulated on one GPU): with a common seed the shards [base, base + n)
    concatenate to the unsharded result bitwise -- also shards of ONE sample, which must not fall back to sequential
    passes (those always draw sample index 0)."""
    import stag_b200 as stag
    from stag_b200 import parallel as P
    rng = np.random.default_rng(0)
    N, E, S = 3000, 20000, 4
    g = stag.Graph(torch.from_numpy(rng.integers(0, N, E)), torch.from_numpy(rng.integers(0, N, E)), N).to("cuda")
    x = torch.from_numpy(rng.standard_normal((N, 32)).astype(np.float32)).cuda()
    torch.manual_seed(0)
    mk = lambda d: torch.distributions.Normal(torch.ones(d), 0.3 * torch.ones(d))  # noqa: E731
    layers = torch.nn.ModuleList([
        stag.layers.StagLayer(stag.zoo.GCN(32, 64, activation=torch.relu), q_a=mk(32), p_a=mk(32), vi=True),
        stag.layers.StagLayer(stag.zoo.GCN(64, 10), q_a=mk(64), p_a=mk(64), vi=True),
    ]).cuda()
    model = stag.models.StagModel(layers, kl_scaling=0.1)
    stag.manual_seed(5)
    full = model._forward_samples(g, x, S, sample_base=0).detach()
    assert not torch.equal(full[0], full[1])
    for world in (2, 4):
        parts = []
        for rank in range(world):
            base, n = P.shard_samples(S, rank, world)
            stag.manual_seed(5)
            parts.append(model._forward_samples(g, x, n, sample_base=base).detach())
        assert torch.equal(torch.cat(parts), full), world


@pytest.mark.parametrize("N,E,H,F", [(700, 9000, 4, 16), (300, 2000, 3, 7), (5000, 60000, 8, 32), (64, 0, 2, 8)])
def test_heads_aggregate_equals_the_expanded_weights_path(N, E, H, F):
    """The per-head aggregation of the GAT path (one launch per head, strided in place: ops.heads_aggregate) against the
    single launch over expanded [E, H F] weights and against a float64 scatter formulation: forward, d ft, d attention."""
    import stag_b200 as stag
    rng = np.random.default_rng(N + E + H)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    if E > 2000:
        dst[:1500] = 5          # a hub row
    g = stag.Graph(torch.from_numpy(src), torch.from_numpy(dst), N).to("cuda")
    ft0 = torch.from_numpy(rng.standard_normal((N, H, F)).astype(np.float32)).cuda()
    a0 = torch.from_numpy(rng.random((E, H)).astype(np.float32)).cuda()
    go = torch.from_numpy(rng.standard_normal((N, H, F)).astype(np.float32)).cuda()
    ft1, a1 = ft0.clone().requires_grad_(True), a0.clone().requires_grad_(True)
    out1 = stag.ops.heads_aggregate(g, ft1, a1)
    out1.backward(go)
    ft3, a3 = ft0.double().requires_grad_(True), a0.double().requires_grad_(True)
    msg = ft3[torch.from_numpy(src).cuda()] * a3.unsqueeze(-1)
    out3 = torch.zeros(N, H, F, dtype=torch.float64, device="cuda").index_add(0, torch.from_numpy(dst).cuda(), msg)
    out3.backward(go.double())
    rel = lambda u, v: float((u.double() - v).abs().max() / v.abs().max().clamp(min=1e-30))  # noqa: E731
    assert out1.shape == (N, H, F) and rel(out1, out3) < 1e-5
    if E:
        assert rel(ft1.grad, ft3.grad) < 1e-5 and rel(a1.grad, a3.grad) < 1e-5
        ft2, a2 = ft0.clone().requires_grad_(True), a0.clone().requires_grad_(True)
        aw = a2.unsqueeze(-1).expand(E, H, F).reshape(E, H * F)
        out2 = stag.ops.stochastic_aggregate(g, ft2.reshape(N, H * F), aw).view(N, H, F)
        out2.backward(go)
        assert rel(out1, out2) < 1e-6 and rel(ft1.grad, ft2.grad) < 1e-6 and rel(a1.grad, a2.grad) < 1e-5
    # only ft needs a gradient: the transposed pass alone
    ft4 = ft0.clone().requires_grad_(True)
    stag.ops.heads_aggregate(g, ft4, a0).backward(go)
    if E:
        assert rel(ft4.grad, ft3.grad) < 1e-5


def test_gat_layer_takes_the_per_head_route_on_larger_graphs(monkeypatch):
    """zoo.GAT above the 4 MB expansion threshold runs ops.heads_aggregate; same outputs and parameter gradients as the
    expanded single-launch route (forced here by patching the operator)."""
    import stag_b200 as stag
    from stag_b200 import ops
    rng = np.random.default_rng(2)
    N, E, H, F = 6000, 80000, 4, 8
    g = stag.Graph(torch.from_numpy(rng.integers(0, N, E)), torch.from_numpy(rng.integers(0, N, E)), N).to("cuda")
    x = torch.from_numpy(rng.standard_normal((N, 24)).astype(np.float32)).cuda()
    w = torch.from_numpy((1 + 0.3 * rng.standard_normal((E, H))).astype(np.float32)).cuda()
    go = torch.from_numpy(rng.standard_normal((N, H * F)).astype(np.float32)).cuda()
    torch.manual_seed(0)
    layer = stag.zoo.GAT(24, F, num_heads=H, allow_zero_in_degree=True).cuda()
    calls = []
    real = ops.heads_aggregate

    def spy(graph, ft, a):
        calls.append(tuple(a.shape))
        return real(graph, ft, a)

    def expanded(graph, ft, a):
        aw = a.unsqueeze(-1).expand(a.shape[0], H, F).reshape(a.shape[0], H * F)
        return ops.stochastic_aggregate(graph, ft.reshape(N, H * F), aw).view(N, H, F)

    res = []
    for fn in (spy, expanded):
        monkeypatch.setattr(ops, "heads_aggregate", fn)
        layer.zero_grad()
        out = layer(g, x, edge_weight=w)
        out.backward(go)
        res.append((out.detach().clone(), [p.grad.clone() for p in layer.parameters()]))
    assert calls == [(E, H)]
    rel = lambda u, v: float((u - v).abs().max() / v.abs().max().clamp(min=1e-30))  # noqa: E731
    assert rel(res[0][0], res[1][0]) < 1e-5
    for ga, gb in zip(res[0][1], res[1][1]):
        assert rel(ga, gb) < 1e-4


@pytest.mark.parametrize("N,E,H,with_w", [(500, 4000, 4, True), (300, 2500, 3, True), (2000, 30000, 8, False),
                                          (400, 3000, 1, True), (50, 0, 4, True), (900, 9000, 5, True)])
def test_attention_softmax_against_the_torch_formulation(N, E, H, with_w):
    """stag_attention_softmax (logits w * leaky_relu(el[u] + er[v]) computed inside the segmented softmax) against the
    reference's chain in float64 (stag/zoo/gat.py:113-122: u_add_v, leaky_relu, noise product, edge_softmax): the
    attention and its gradients w.r.t. el, er and the edge noise; nodes without in-edges, a hub, any number of heads."""
    import stag_b200 as stag
    rng = np.random.default_rng(N + E + H)
    src, dst = rng.integers(0, N, E), rng.integers(0, N - 7, E)       # the last 7 nodes have no in-edges
    if E > 5000:
        dst[:2000] = 3
    T = torch.from_numpy
    g = stag.Graph(T(src), T(dst), N).to("cuda")
    el0 = T(rng.standard_normal((N, H)).astype(np.float32)).cuda()
    er0 = T(rng.standard_normal((N, H)).astype(np.float32)).cuda()
    w0 = T((1 + 0.4 * rng.standard_normal((E, H))).astype(np.float32)).cuda() if with_w else None
    da = T(rng.standard_normal((E, H)).astype(np.float32)).cuda()
    el, er = el0.clone().requires_grad_(True), er0.clone().requires_grad_(True)
    w = None if w0 is None else w0.clone().requires_grad_(True)
    a = stag.ops.attention_softmax(g, el, er, w, 0.2)
    assert a.shape == (E, H)
    a.backward(da)
    # float64 scatter formulation
    eld, erd = el0.double().requires_grad_(True), er0.double().requires_grad_(True)
    wd = None if w0 is None else w0.double().requires_grad_(True)
    s_, d_ = T(src).cuda(), T(dst).cuda()
    lg = torch.nn.functional.leaky_relu(eld[s_] + erd[d_], 0.2)
    if wd is not None:
        lg = lg * wd
    if E:
        m = torch.full((N, H), -float("inf"), dtype=torch.float64, device="cuda").scatter_reduce(
            0, d_[:, None].expand(E, H), lg, "amax", include_self=True)
        ex = torch.exp(lg - m[d_])
        ref = ex / torch.zeros(N, H, dtype=torch.float64, device="cuda").index_add(0, d_, ex)[d_]
        ref.backward(da.double())
        rel = lambda u, v: float((u.double() - v).abs().max() / v.abs().max().clamp(min=1e-30))  # noqa: E731
        assert rel(a.detach(), ref.detach()) < 1e-5
        assert rel(el.grad, eld.grad) < 2e-5 and rel(er.grad, erd.grad) < 2e-5
        if w is not None:
            assert rel(w.grad, wd.grad) < 2e-5
        sums = torch.zeros(N, H, device="cuda").index_add(0, d_, a.detach())
        assert float((sums[: N - 7][torch.bincount(d_, minlength=N)[: N - 7] > 0] - 1).abs().max()) < 1e-5
        # run-to-run bitwise (no atomics)
        el2, er2 = el0.clone().requires_grad_(True), er0.clone().requires_grad_(True)
        stag.ops.attention_softmax(g, el2, er2, w0, 0.2).backward(da)
        assert torch.equal(el2.grad, el.grad) and torch.equal(er2.grad, er.grad)
    else:
        assert float(el.grad.abs().sum()) == 0.0 and float(er.grad.abs().sum()) == 0.0
