"""GPU: the fused aggregation kernels, driven through the C ABI (ctypes, stag_b200.ops), agree
with the reference-made golden vectors and with the oracle under SHARED EXTERNAL NOISE to 1e-5
relative (north_star tolerance, fp32)."""
import numpy as np
import pytest
import torch

from conftest import golden, LAYER_CASES
from oracle import ref_layers, ref_spmm

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def close(a, b, rtol=RTOL, what=""):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    err = np.abs(a - b).max() / scale if b.size else 0.0
    assert err <= rtol, "%s: max err / max|ref| = %.3e > %.1e" % (what, err, rtol)
    # ... and ELEMENTWISE: a small entry (zero-in-degree neighbourhoods, hub tails) may not be wrong by
    # more than rtol of ITS OWN value plus the fp32 cancellation floor of a sum of terms of size `scale`
    # (same form as tests/test_oracle_cpu.py)
    if b.size:
        bad = np.abs(a - b) > rtol * np.abs(b) + 1e-6 * scale
        assert not bad.any(), "%s: %d entries beyond rtol %.1e elementwise (worst |a-b| = %.3e at |ref| = %.3e)" % (
            what, int(bad.sum()), rtol, np.abs(a - b)[bad].max(), np.abs(b)[bad][np.abs(a - b)[bad].argmax()])


def make_base(name, d):
    from stag_b200 import zoo
    kind = ref_layers.case_kind(name)
    D = d["feat"].shape[1]
    if kind == "gcn":
        has_w = "p_weight" in d
        out = d["p_weight"].shape[1] if has_w else D
        m = zoo.GCN(D, out, norm=ref_layers.gcn_norm_of(name), weight=has_w, bias="p_bias" in d)
    elif kind.startswith("sage"):
        m = zoo.GraphSAGE(D, d["p_fc_neigh__weight"].shape[0], aggregator_type=kind.split("_")[1])
    elif kind == "gat":
        _, H, F = d["p_attn_l"].shape
        last = name.startswith("gat_last")
        m = zoo.GAT(D, F, num_heads=H, last=last, residual="p_res_fc__weight" in d,
                    activation=torch.nn.functional.elu if last else None)
    else:
        m = zoo.GIN(D, d["p_apply_func__weight"].shape[0])
    sd = {k[2:].replace("__", "."): torch.from_numpy(d[k]) for k in d.files if k.startswith("p_")}
    m.load_state_dict(sd)
    return m.cuda()


@pytest.mark.parametrize("name", LAYER_CASES)
def test_layer_matches_reference_golden(name):
    """StagLayer(zoo.X) of this repo, fed the SAME noise tensor through the reference's seam
    (rsample_noise), reproduces the reference's outputs and gradients."""
    import stag_b200 as sb
    d = golden(name)
    g = sb.Graph(torch.from_numpy(d["src"]), torch.from_numpy(d["dst"]), int(d["num_nodes"])).to("cuda")
    base = make_base(name, d)
    layer = sb.layers.StagLayer(base, relu=bool(d["relu"]), norm=bool(d["in_norm"]))
    feat = torch.from_numpy(d["feat"]).cuda().requires_grad_(True)
    w = torch.from_numpy(d["w"]).cuda().requires_grad_("dw" in d)
    layer.rsample_noise = lambda graph, sample_dimension, **kw: w
    out = layer(g, feat)
    out.backward(torch.from_numpy(d["gout"]).cuda())
    close(out, d["out"], what="out")
    close(feat.grad, d["dfeat"], what="dfeat")
    if "dw" in d:
        close(w.grad, d["dw"], what="dw")
    if "w_used" in d:
        close(layer._edge_weight_sample, d["w_used"], what="w_used")
    for k, p in base.named_parameters():
        key = "g_" + k.replace(".", "__")
        if key in d:
            close(p.grad, d[key], what=k)


def rand_case(n, e, D, K, seed, hub=0):
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    if hub:
        dst[:hub] = n // 2
        src[hub:2 * hub] = n // 3
    x = rng.standard_normal((n, D)).astype(np.float32)
    w = (1 + 0.4 * rng.standard_normal((e, K))).astype(np.float32)
    gout = rng.standard_normal((n, D)).astype(np.float32)
    ss = rng.uniform(0.5, 1.5, n).astype(np.float32)
    ds = rng.uniform(0.5, 1.5, n).astype(np.float32)
    return src, dst, x, w, gout, ss, ds


@pytest.mark.parametrize("n,e,D,K,hub", [
    (1, 0, 4, 4, 0), (1, 3, 4, 4, 0), (5, 0, 7, 7, 0), (64, 500, 1, 1, 0), (64, 500, 3, 3, 0),
    (64, 500, 4, 1, 0), (200, 3000, 9, 9, 0), (200, 3000, 16, 16, 200), (200, 3000, 50, 50, 0),
    (300, 5000, 100, 100, 600), (300, 5000, 128, 128, 600), (300, 5000, 128, 1, 600),
    (100, 2000, 256, 256, 0), (60, 700, 1433, 1433, 150), (60, 700, 1433, 1, 150), (3000, 40000, 121, 121, 0),
])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_external_noise_fwd_bwd_vs_oracle(n, e, D, K, hub, reduce):
    import stag_b200 as sb
    src, dst, x, w, gout, ss, ds = rand_case(n, e, D, K, seed=n + e + D + K, hub=hub)
    T = torch.from_numpy
    # oracle (CPU, autograd)
    xo, wo = T(x).requires_grad_(True), T(w).requires_grad_(True)
    oo = ref_spmm.aggregate(T(src), T(dst), n, xo, wo, reduce=reduce, src_scale=T(ss), dst_scale=T(ds))
    oo.backward(T(gout))
    # CUDA
    g = sb.Graph(T(src), T(dst), n).to("cuda")
    xc, wc = T(x).cuda().requires_grad_(True), T(w).cuda().requires_grad_(True)
    oc = sb.ops.stochastic_aggregate(g, xc, wc, reduce=reduce, src_scale=T(ss).cuda(), dst_scale=T(ds).cuda())
    oc.backward(T(gout).cuda())
    close(oc, oo, what="out")
    close(xc.grad, xo.grad, what="dx")
    if e:
        close(wc.grad, wo.grad, what="dw")


def test_no_edge_weight_is_copy_u_sum():
    import stag_b200 as sb
    src, dst, x, w, gout, ss, ds = rand_case(500, 6000, 32, 32, seed=3, hub=300)
    T = torch.from_numpy
    g = sb.Graph(T(src), T(dst), 500).to("cuda")
    xc = T(x).cuda().requires_grad_(True)
    oc = sb.ops.stochastic_aggregate(g, xc, None)
    oc.backward(T(gout).cuda())
    xo = T(x).requires_grad_(True)
    oo = ref_spmm.aggregate(T(src), T(dst), 500, xo, None)
    oo.backward(T(gout))
    close(oc, oo)
    close(xc.grad, xo.grad)


def test_sample_batched_external_noise():
    """[S,E,K] noise with a shared X (first layer) and with per-sample X."""
    import stag_b200 as sb
    S = 3
    src, dst, x, w, gout, ss, ds = rand_case(150, 2000, 20, 20, seed=11, hub=200)
    rng = np.random.default_rng(12)
    ws = (1 + 0.3 * rng.standard_normal((S, 2000, 20))).astype(np.float32)
    gs = rng.standard_normal((S, 150, 20)).astype(np.float32)
    xs = rng.standard_normal((S, 150, 20)).astype(np.float32)
    T = torch.from_numpy
    g = sb.Graph(T(src), T(dst), 150).to("cuda")
    for shared in (True, False):
        xin = x if shared else xs
        xc, wc = T(xin).cuda().requires_grad_(True), T(ws).cuda().requires_grad_(True)
        oc = sb.ops.stochastic_aggregate(g, xc, wc, src_scale=T(ss).cuda(), dst_scale=T(ds).cuda(),
                                         n_samples=S)
        assert oc.shape == (S, 150, 20)
        oc.backward(T(gs).cuda())
        xo, wo = T(xin).requires_grad_(True), T(ws).requires_grad_(True)
        outs = [ref_spmm.aggregate(T(src), T(dst), 150, xo if shared else xo[s], wo[s], src_scale=T(ss),
                                   dst_scale=T(ds)) for s in range(S)]
        oo = torch.stack(outs)
        oo.backward(T(gs))
        close(oc, oo, what="out")
        close(xc.grad, xo.grad, what="dx")
        close(wc.grad, wo.grad, what="dw")


def test_cpu_tensor_is_rejected_loudly():
    import stag_b200 as sb
    g = sb.Graph(torch.tensor([0, 1]), torch.tensor([1, 0]), 2)
    with pytest.raises(sb.StagLibraryError):
        sb.ops.stochastic_aggregate(g, torch.zeros(2, 4), None)


def test_readout_golden():
    import stag_b200 as sb
    d = golden("readout")
    bnn = d["batch_num_nodes"]
    gs = [sb.Graph(torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64), int(n)) for n in bnn]
    bg = sb.batch(gs).to("cuda")
    f = torch.from_numpy(d["feat"]).cuda().requires_grad_(True)
    s = sb.layers.SumNodes()(bg, f)
    m = sb.layers.MeanNodes()(bg, f)
    close(s, d["sum"])
    close(m, d["mean"])
    (s.sum() + m.sum()).backward()
    want = 1.0 + 1.0 / np.repeat(bnn, bnn).astype(np.float64)
    close(f.grad, np.broadcast_to(want[:, None], d["feat"].shape).astype(np.float32))


def test_host_buffer_entry_point_vs_c_restatement():
    """stag_aggregate_host: the whole path (CSC/CSR build, fused forward, fused transposed pass) from HOST
    buffers through the bare C ABI -- no torch anywhere -- against the C restatement of the reference."""
    import ctypes
    from stag_b200 import _lib
    from oracle import ref_c
    lib = _lib.load()
    rng = np.random.default_rng(21)
    N, E, D, S = 3000, 40000, 128, 3
    src = rng.integers(0, N, E).astype(np.int64)
    dst = rng.integers(0, N, E).astype(np.int64)
    dst[:500] = 11
    x = rng.standard_normal((N, D)).astype(np.float32)
    dout = rng.standard_normal((S, N, D)).astype(np.float32)
    loc, scale = np.array([1.0], np.float32), np.array([0.4], np.float32)
    nz = _lib.StagNoise()
    nz.kind, nz.K, nz.param_shape = _lib.NOISE_NORMAL, D, _lib.PARAM_SCALAR
    nz.relu = nz.in_norm = nz.sample_base = 0
    nz.p0, nz.p1, nz.external = loc.ctypes.data, scale.ctypes.data, 0     # HOST pointers for this entry point
    nz.seed, nz.offset = 123, 4
    out = np.empty((S, N, D), np.float32)
    dx = np.empty((N, D), np.float32)
    rc = lib.stag_aggregate_host(0, src.ctypes.data, dst.ctypes.data, E, N, x.ctypes.data, dout.ctypes.data, D, S,
                                 ctypes.byref(nz), 1, out.ctypes.data, dx.ctypes.data)
    assert rc == 0, lib.stag_last_error()
    dx_ref = np.zeros((N, D), np.float64)
    for s in range(S):
        lp = ref_c.LayerPass(src, dst, N, x, dout[s], "normal", 1.0, 0.4, vi=False, gcn_norm=True)
        o, d = lp.run(sample=s, seed=123, offset=4)
        assert np.abs(out[s] - o).max() <= 2e-5 * np.abs(o).max()
        dx_ref += d
    assert np.abs(dx - dx_ref).max() <= 2e-5 * np.abs(dx_ref).max()
