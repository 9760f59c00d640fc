"""TEST INFRASTRUCTURE ONLY -- mint the golden vectors under tests/golden/.

Runs the reference's OWN code (/root/reference/stag, unmodified, imported from where it
lies) on top of oracle/dgl_shim (DGL is third-party, un-vendored and not installable
here) with fixed seeds, feeding it EXTERNAL noise through the reference's own seam
(`StagLayer.rsample_noise` is the method the layer calls to obtain the [E,K] tensor,
stag/layers.py:96,115-129; `zoo.*.forward(edge_weight=)` is the operator boundary,
stag/layers.py:109-113) and records inputs, outputs and autograd gradients.

The reference has no golden vectors or numerical tests of its own (stag/tests are
shape-only), so these fixtures are the pin: tests/test_oracle_cpu.py checks oracle/ref_*.py
against them, tests/test_gpu_*.py check the CUDA path against them.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

This script only runs in the build container (it needs /root/reference); the GPU box uses
the committed .npz files.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("STAG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, os.path.join(HERE, "dgl_shim"))
    sys.path.insert(0, REFERENCE)
    import dgl  # noqa: F401  (the shim)
    import stag  # the reference
    assert os.path.realpath(stag.__file__).startswith(os.path.realpath(REFERENCE)), stag.__file__
    return stag, dgl


def np_(t):
    return None if t is None else t.detach().cpu().numpy()


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    arrays = {k: v for k, v in arrays.items() if v is not None}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote %-28s %s" % (name, {k: getattr(v, "shape", v) for k, v in arrays.items() if k in ("feat", "w")}))


def messy_graph(dgl, n, e, seed, hub=None):
    """Random multigraph with self loops, duplicate edges, zero-in-degree and
    zero-out-degree nodes; optionally one hub destination / source of degree `hub`."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n - 2, (e,), generator=g)      # node n-1, n-2 never a source
    dst = torch.randint(2, n, (e,), generator=g)          # node 0, 1 never a destination
    src[:3] = dst[:3]                                      # self loops
    src[3:6], dst[3:6] = src[6], dst[6]                    # duplicates
    if hub:
        dst[10:10 + hub] = 5
        src[20 + hub:20 + 2 * hub] = 7
    return dgl.graph((src, dst), num_nodes=n)


def powerlaw_graph(dgl, n, e, seed):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, n + 1) ** 0.9
    p /= p.sum()
    dst = rng.choice(n, size=e, p=p)
    src = rng.choice(n, size=e, p=p[::-1])
    perm = rng.permutation(n)
    return dgl.graph((torch.from_numpy(perm[src]), torch.from_numpy(perm[dst])), num_nodes=n)


def layer_case(stag, name, g, base, feat, w, relu=False, norm=False, w_grad=True, seed=0):
    """One StagLayer forward/backward of the reference with the noise tensor `w` supplied
    through rsample_noise (pre-relu, pre-in-norm, exactly where the reference draws it)."""
    layer = stag.layers.StagLayer(base, relu=relu, norm=norm)
    feat = feat.clone().requires_grad_(True)
    w_in = w.clone().requires_grad_(w_grad)
    layer.rsample_noise = lambda graph, sample_dimension: w_in
    out = layer(g, feat)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 99))
    out.backward(gout)
    src, dst = g.edges()
    params = {("p_" + k.replace(".", "__")): np_(v) for k, v in base.state_dict().items()}
    grads = {("g_" + k.replace(".", "__")): np_(v.grad) for k, v in base.named_parameters() if v.grad is not None}
    save(name, src=np_(src), dst=np_(dst), num_nodes=np.int64(g.number_of_nodes()), feat=np_(feat), w=np_(w),
         relu=np.int64(relu), in_norm=np.int64(norm), out=np_(out), gout=np_(gout), dfeat=np_(feat.grad),
         dw=np_(w_in.grad) if w_grad else None,
         w_used=np_(layer._edge_weight_sample) if (relu or norm) else None, **params, **grads)


def gat_cases(stag, dgl):
    """GAT (stag/zoo/gat.py:39-145): noise [E, num_heads] multiplies the leaky-relu logits before the segmented softmax.
    Every destination of these graphs has an in-edge (self loops added), as the reference scripts arrange."""
    T = torch
    rn = lambda *s, seed=0: T.randn(*s, generator=T.Generator().manual_seed(seed))  # noqa: E731
    gm = dgl.add_self_loop(messy_graph(dgl, 50, 300, 7))
    E = gm.number_of_edges()
    T.manual_seed(400)
    layer_case(stag, "gat_h4", gm, stag.zoo.GAT(24, 8, num_heads=4), rn(50, 24, seed=401), 1.0 + 0.4 * rn(E, 4, seed=402))
    T.manual_seed(403)
    layer_case(stag, "gat_last_residual", gm, stag.zoo.GAT(24, 6, num_heads=3, last=True, residual=True,
                                                             activation=T.nn.functional.elu),
               rn(50, 24, seed=404), 1.0 + 0.4 * rn(E, 3, seed=405))


def main():
    stag, dgl = import_reference()
    T = torch
    rn = lambda *s, seed=0: T.randn(*s, generator=T.Generator().manual_seed(seed))  # noqa: E731
    if "--gat-only" in sys.argv:   # added after the other fixtures were committed: leaves them byte-identical
        return gat_cases(stag, dgl)

    # --- the reference's own four test shapes (stag/tests/test_layers.py:13-54) --------------
    T.manual_seed(1234)
    g3 = dgl.rand_graph(3, 9)
    T.manual_seed(1)
    layer_case(stag, "t_r1_gcn", g3, stag.zoo.GCN(16, 32), rn(3, 16, seed=2), 1.0 + rn(9, 16, seed=3))
    T.manual_seed(2)
    layer_case(stag, "t_re_gcn", g3, stag.zoo.GCN(16, 32), rn(3, 16, seed=4), 1.0 + rn(9, 1, seed=5))

    # --- degree normalisations, zero-degree nodes, multi-edges, self loops ---------------------
    gm = messy_graph(dgl, 50, 300, 7)
    for i, norm in enumerate(("both", "right", "left", "none")):
        T.manual_seed(10 + i)
        layer_case(stag, "gcn_%s" % norm, gm, stag.zoo.GCN(24, 8, norm=norm), rn(50, 24, seed=20 + i),
                   1.0 + 0.4 * rn(300, 24, seed=30 + i))
    T.manual_seed(15)
    layer_case(stag, "gcn_k1", gm, stag.zoo.GCN(24, 8), rn(50, 24, seed=40), 1.0 + 0.4 * rn(300, 1, seed=41))
    T.manual_seed(16)
    layer_case(stag, "gcn_relu", gm, stag.zoo.GCN(24, 8), rn(50, 24, seed=42), rn(300, 24, seed=43), relu=True)
    T.manual_seed(17)  # unaligned width, no weight
    layer_case(stag, "gcn_d18_noweight", gm, stag.zoo.GCN(18, 18, weight=False, bias=False), rn(50, 18, seed=44),
               1.0 + rn(300, 18, seed=45))
    # Bernoulli + in-norm as in scripts/arxiv_mle/gcn/run.py:70-74 (rows whose weights sum to 0 keep scale 1)
    T.manual_seed(18)
    wb = (T.rand(300, 24, generator=T.Generator().manual_seed(46)) < 0.6).float()
    layer_case(stag, "gcn_bernoulli_innorm", gm, stag.zoo.GCN(24, 8), rn(50, 24, seed=47), wb, norm=True, w_grad=False)
    wb1 = (T.rand(300, 1, generator=T.Generator().manual_seed(48)) < 0.5).float()
    layer_case(stag, "gcn_bernoulli_innorm_k1", gm, stag.zoo.GCN(24, 8), rn(50, 24, seed=49), wb1, norm=True,
               w_grad=False)

    # --- GraphSAGE mean / gcn aggregators (stag/zoo/graph_sage.py:70-91) ------------------------
    for i, agg in enumerate(("mean", "gcn")):
        T.manual_seed(50 + i)
        layer_case(stag, "sage_%s" % agg, gm, stag.zoo.GraphSAGE(24, 8, aggregator_type=agg), rn(50, 24, seed=60 + i),
                   1.0 + 0.4 * rn(300, 24, seed=70 + i))
    # --- GIN / GatedGCN share the same aggregation -----------------------------------------------
    T.manual_seed(53)
    layer_case(stag, "gin", gm, stag.zoo.GIN(24, 8), rn(50, 24, seed=62), 1.0 + 0.4 * rn(300, 24, seed=72))

    # --- hubs: rows far above the kernel's hub threshold, both as destination and as source ----
    gh = messy_graph(dgl, 300, 1500, 9, hub=300)
    T.manual_seed(80)
    layer_case(stag, "hub_d128", gh, stag.zoo.GCN(128, 16), rn(300, 128, seed=81), 1.0 + 0.4 * rn(1500, 128, seed=82))
    T.manual_seed(83)
    layer_case(stag, "hub_d20_k1", gh, stag.zoo.GCN(20, 4), rn(300, 20, seed=84), 1.0 + 0.4 * rn(1500, 1, seed=85))
    gp = powerlaw_graph(dgl, 1000, 4000, 11)
    T.manual_seed(86)
    layer_case(stag, "powerlaw_d50", gp, stag.zoo.GCN(50, 12), rn(1000, 50, seed=87), 1.0 + 0.4 * rn(4000, 50, seed=88))

    # --- model level: StagModel.loss with vi=True, per-channel Normal ("rc"), S=2 ----------------
    T.manual_seed(100)
    N, E, D0, H, C, S = 60, 400, 12, 8, 4, 2
    g = messy_graph(dgl, N, E, 13)
    g = dgl.add_self_loop(g)
    E = g.number_of_edges()
    mk = lambda d, s: T.distributions.Normal(T.ones(d), s * T.ones(d))  # noqa: E731
    layers = T.nn.ModuleList([
        stag.layers.StagLayer(stag.zoo.GCN(D0, H, activation=T.relu), q_a=mk(D0, 0.3), p_a=mk(D0, 0.5), vi=True),
        stag.layers.StagLayer(stag.zoo.GCN(H, C, activation=lambda x: T.softmax(x, dim=-1)), q_a=mk(H, 0.2),
                              p_a=mk(H, 0.5), vi=True),
    ])
    model = stag.models.StagModel(layers, kl_scaling=0.5)
    eps = [rn(S, E, D0, seed=101), rn(S, E, H, seed=102)]
    counters = [0, 0]

    def patch(i, layer):
        def rsample(graph, sample_dimension):
            base = layer.q_a.base_distribution
            w = base.loc + eps[i][counters[i]] * base.scale   # torch/distributions/normal.py:82-85
            counters[i] += 1
            return w
        layer.rsample_noise = rsample
    for i, layer in enumerate(layers):
        patch(i, layer)
    feat = rn(N, D0, seed=103)
    y = T.randint(0, C, (N,), generator=T.Generator().manual_seed(104))
    mask = T.rand(N, generator=T.Generator().manual_seed(105)) < 0.7
    nll, reg = model.loss_terms(g, feat, y, mask=mask, n_samples=S)
    (nll + reg).backward()
    src, dst = g.edges()
    sd = {("p_" + k.replace(".", "__")): np_(v) for k, v in layers.state_dict().items()}
    gd = {("g_" + k.replace(".", "__")): np_(v.grad) for k, v in layers.named_parameters() if v.grad is not None}
    save("model_rc_vi", src=np_(src), dst=np_(dst), num_nodes=np.int64(N), feat=np_(feat), y=np_(y), mask=np_(mask),
         eps0=np_(eps[0]), eps1=np_(eps[1]), nll=np_(nll), reg=np_(reg), kl_scaling=np.float64(0.5), **sd, **gd)

    # --- amortised posteriors "re" [E,1] and "rec" [E,D] (stag/tests/test_layers.py:34-54), vi=True
    for tag, outf in (("re", 1), ("rec", 16)):
        T.manual_seed(200 + outf)
        ga = messy_graph(dgl, 40, 200, 17)
        Ea = ga.number_of_edges()
        q_a = stag.distributions.AmortizedDistribution(16, outf)
        layer = stag.layers.StagLayer(stag.zoo.GCN(16, 32), q_a=q_a, p_a=T.distributions.Normal(1.0, 1.0), vi=True)
        e_ = rn(Ea, 16, seed=201)

        def rsample(graph, sample_dimension, layer=layer, e_=e_):
            d = layer.q_a.expand([graph.number_of_edges(), sample_dimension])
            return d.loc + e_ * d.scale
        layer.rsample_noise = rsample
        feat = rn(40, 16, seed=202).requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            out = layer(ga, feat)
        kl = layer.kl_divergence()
        gout = rn(*out.shape, seed=203)
        ((out * gout).sum() + kl).backward()
        src, dst = ga.edges()
        sd = {("p_" + k.replace(".", "__")): np_(v) for k, v in layer.state_dict().items()}
        gd = {("g_" + k.replace(".", "__")): np_(v.grad) for k, v in layer.named_parameters() if v.grad is not None}
        save("amortized_%s" % tag, src=np_(src), dst=np_(dst), num_nodes=np.int64(40), feat=np_(feat), eps=np_(e_),
             out=np_(out), gout=np_(gout), kl=np_(kl), dfeat=np_(feat.grad), **sd, **gd)

    # --- readout (stag/layers.py:156-178) -----------------------------------------------------------
    gs = [messy_graph(dgl, n, 3 * n, 300 + n) for n in (5, 9, 4, 17)]
    bg = dgl.batch(gs)
    f = rn(bg.number_of_nodes(), 10, seed=301)
    save("readout", batch_num_nodes=np_(bg.batch_num_nodes()), feat=np_(f),
         sum=np_(stag.layers.SumNodes()(bg, f)), mean=np_(stag.layers.MeanNodes()(bg, f)))

    gat_cases(stag, dgl)


if __name__ == "__main__":
    main()
