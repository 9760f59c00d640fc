"""GPU: the fused counter-based RNG.  (1) the emitted noise equals the numpy restatement of the
generator (integer stream bit-exact -> uniforms / Bernoulli exact, normals to 1e-5); (2) the law
is right: moments and Kolmogorov-Smirnov against the named distributions; (3) the fused forward
and backward consume exactly the emitted noise (regeneration equality), so they agree with the
oracle fed that tensor."""
import numpy as np
import pytest
import torch
from scipy import stats

from oracle import ref_philox, ref_spmm

pytestmark = pytest.mark.gpu


def spec(kind, p0, p1, K, E, **kw):
    from stag_b200.ops import NoiseSpec
    t = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.float32).cuda()  # noqa: E731
    return NoiseSpec(kind, t(p0), t(p1), K, E, **kw)


@pytest.mark.parametrize("K", [1, 3, 16, 50])
def test_emit_matches_numpy_restatement(K):
    E, seed, off = 777, 0xDEADBEEFCAFE, (7 << 32) + 5
    for s_base in (0, 3):
        sp = spec("uniform", 0.25, 1.75, K, E, seed=seed, offset=off, sample_base=s_base)
        w = sp.materialize(n_samples=2).cpu().numpy()
        for s in range(2):
            ref = ref_philox.noise("uniform", E, K, s_base + s, seed, off, 0.25, 1.75)
            np.testing.assert_allclose(w[s], ref, rtol=0, atol=2e-7)
        sp = spec("bernoulli", 0.3, None, K, E, seed=seed, offset=off, sample_base=s_base)
        w = sp.materialize(n_samples=2).cpu().numpy()
        for s in range(2):
            assert np.array_equal(w[s], ref_philox.noise("bernoulli", E, K, s_base + s, seed, off, 0.3))
        sp = spec("normal", 1.0, 0.5, K, E, seed=seed, offset=off, sample_base=s_base)
        w = sp.materialize(n_samples=2).cpu().numpy()
        for s in range(2):
            ref = ref_philox.noise("normal", E, K, s_base + s, seed, off, 1.0, 0.5)
            np.testing.assert_allclose(w[s], ref, rtol=0, atol=2e-5)


def test_per_channel_and_per_edge_parameters():
    E, K = 500, 8
    rng = np.random.default_rng(0)
    loc, scale = rng.normal(1, 0.2, K).astype(np.float32), rng.uniform(0.1, 0.5, K).astype(np.float32)
    w = spec("normal", loc, scale, K, E, seed=1, offset=2).materialize().cpu().numpy()
    np.testing.assert_allclose(w, ref_philox.noise("normal", E, K, 0, 1, 2, loc, scale), atol=2e-5)
    loc_e, scale_e = rng.normal(1, 0.2, (E, 1)).astype(np.float32), rng.uniform(0.1, 0.5, (E, 1)).astype(np.float32)
    w = spec("normal", loc_e, scale_e, K, E, seed=1, offset=2).materialize().cpu().numpy()
    np.testing.assert_allclose(w, ref_philox.noise("normal", E, K, 0, 1, 2, loc_e, scale_e), atol=2e-5)
    loc_ec, scale_ec = rng.normal(1, 0.2, (E, K)).astype(np.float32), rng.uniform(0.1, 0.5, (E, K)).astype(np.float32)
    w = spec("normal", loc_ec, scale_ec, K, E, seed=1, offset=2).materialize().cpu().numpy()
    np.testing.assert_allclose(w, ref_philox.noise("normal", E, K, 0, 1, 2, loc_ec, scale_ec), atol=2e-5)


def test_moments_and_ks():
    E, K = 20000, 64   # 1.28e6 draws per law
    n = E * K
    w = spec("normal", 1.0, 0.4, K, E, seed=42, offset=0).materialize().cpu().numpy().astype(np.float64)
    z = (w - 1.0) / 0.4
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2.0 / n)
    assert abs(stats.skew(z.ravel())) < 5 * np.sqrt(6.0 / n)
    assert abs(stats.kurtosis(z.ravel())) < 5 * np.sqrt(24.0 / n)
    assert stats.kstest(z.ravel(), "norm").pvalue > 1e-3
    # per-channel moments within 5 sigma
    assert np.all(np.abs(z.mean(0)) < 5 / np.sqrt(E))
    assert np.all(np.abs(z.var(0) - 1) < 5 * np.sqrt(2.0 / E))
    # tails reach out: 1.28e6 draws should exceed 4 sigma a few dozen times
    assert 30 < (np.abs(z) > 4).sum() < 160
    lo, hi = 1 - 0.4 * np.sqrt(3), 1 + 0.4 * np.sqrt(3)
    u = spec("uniform", lo, hi, K, E, seed=42, offset=1).materialize().cpu().numpy().astype(np.float64)
    v = (u - lo) / (hi - lo)
    assert v.min() >= 0 and v.max() < 1
    assert abs(v.mean() - 0.5) < 5 * np.sqrt(1 / 12.0 / n)
    assert abs(v.var() - 1 / 12.0) < 5 * np.sqrt(1 / 180.0 / n)
    assert stats.kstest(v.ravel(), "uniform").pvalue > 1e-3
    p = 0.5 * (1 + np.sqrt(1 - 4 * 0.4 ** 2))   # scripts/arxiv_mle/gcn/run.py:71-72
    b = spec("bernoulli", p, None, K, E, seed=42, offset=2).materialize().cpu().numpy()
    assert set(np.unique(b)) <= {0.0, 1.0}
    k = int(b.sum())
    assert stats.binomtest(k, n, p).pvalue > 1e-3


def test_independence_across_edge_channel_sample_offset():
    E, K = 4096, 32
    a = spec("normal", 0.0, 1.0, K, E, seed=7, offset=0).materialize(n_samples=2).cpu().numpy().astype(np.float64)
    b = spec("normal", 0.0, 1.0, K, E, seed=7, offset=1).materialize().cpu().numpy().astype(np.float64)
    c = spec("normal", 0.0, 1.0, K, E, seed=8, offset=0).materialize().cpu().numpy().astype(np.float64)
    n = E * K
    lim = 5 / np.sqrt(n)
    corr = lambda x, y: np.corrcoef(x.ravel(), y.ravel())[0, 1]  # noqa: E731
    assert abs(corr(a[0], a[1])) < lim            # samples
    assert abs(corr(a[0], b)) < lim               # call offsets
    assert abs(corr(a[0], c)) < lim               # seeds
    assert abs(corr(a[0][:-1], a[0][1:])) < lim   # neighbouring edges
    assert abs(corr(a[0][:, :-1], a[0][:, 1:])) < 5 / np.sqrt(E * (K - 1))   # neighbouring channels
    assert abs(corr(a[0][:, 0::2], a[0][:, 1::2])) < 5 / np.sqrt(n / 2)      # Box-Muller pairs
    assert not np.array_equal(a[0], a[1])


@pytest.mark.parametrize("kind,p0,p1,relu,in_norm", [
    ("normal", 1.0, 0.4, False, False), ("normal", 0.2, 1.0, True, False),
    ("uniform", 0.3, 1.7, False, False), ("bernoulli", 0.6, None, False, True),
    ("bernoulli", 0.6, None, False, False), ("normal", 1.0, 0.4, False, True),
])
@pytest.mark.parametrize("D,K", [(20, 20), (20, 1), (128, 128), (50, 50), (100, 100), (192, 192), (256, 256),
                                 (384, 384), (640, 640)])
def test_fused_forward_backward_consume_the_emitted_noise(kind, p0, p1, relu, in_norm, D, K):
    """Fused (noise never stored) == oracle fed the emitted tensor; dX regenerates the same noise."""
    import stag_b200 as sb
    n, e, S = 300, 5000, 2
    rng = np.random.default_rng(D + K)
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    dst[:400] = 17
    src[400:800] = 19
    x = rng.standard_normal((n, D)).astype(np.float32)
    gout = rng.standard_normal((S, n, D)).astype(np.float32)
    T = torch.from_numpy
    g = sb.Graph(T(src), T(dst), n).to("cuda")
    sp = spec(kind, p0, p1, K, e, relu=relu, in_norm=in_norm, seed=99, offset=4, n_samples=S)
    w = sp.materialize(n_samples=S).cpu()          # relu applied, in-norm not
    xc = T(x).cuda().requires_grad_(True)
    out = sb.ops.stochastic_aggregate(g, xc, sp, n_samples=S)
    out.backward(T(gout).cuda())
    xo = T(x).requires_grad_(True)
    outs = []
    for s in range(S):
        ws = w[s]
        if in_norm:
            ws, _ = ref_spmm.in_norm(T(dst), n, ws)
        outs.append(ref_spmm.aggregate(T(src), T(dst), n, xo, ws))
    oo = torch.stack(outs)
    oo.backward(T(gout))
    err = (out.cpu() - oo).abs().max() / oo.abs().max()
    assert err < 1e-5, err
    err = (xc.grad.cpu() - xo.grad).abs().max() / xo.grad.abs().max()
    assert err < 1e-5, err


@pytest.mark.parametrize("kind", ["normal", "uniform"])
@pytest.mark.parametrize("pshape", ["scalar", "channel", "edge", "edge_channel"])
@pytest.mark.parametrize("D,K", [(24, 24), (24, 1), (128, 128)])
def test_fused_parameter_gradients(kind, pshape, D, K):
    """d(loc), d(scale) (d(low), d(high)) from the fused backward == autograd through
    loc + eps*scale on the emitted eps, for every parameter shape class."""
    import stag_b200 as sb
    from stag_b200 import _lib
    n, e = 200, 3000
    rng = np.random.default_rng(5)
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    dst[:300] = 3
    src[300:600] = 4
    shape = {"scalar": (), "channel": (K,), "edge": (e, 1), "edge_channel": (e, K)}[pshape]
    if pshape == "channel" and K == 1:
        pytest.skip("channel == scalar when K == 1")
    a = (1 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
    b = (0.3 + 0.1 * rng.uniform(size=shape)).astype(np.float32)
    if kind == "uniform":
        a, b = a - 1.0, b + 1.5
    x = rng.standard_normal((n, D)).astype(np.float32)
    gout = rng.standard_normal((n, D)).astype(np.float32)
    T = torch.from_numpy
    g = sb.Graph(T(src), T(dst), n).to("cuda")
    pa, pb = T(np.asarray(a)).cuda().requires_grad_(True), T(np.asarray(b)).cuda().requires_grad_(True)
    sp = sb.ops.NoiseSpec(kind, pa, pb, K, e, seed=5, offset=6)
    xc = T(x).cuda().requires_grad_(True)
    ss = T(rng.uniform(0.5, 1.5, n).astype(np.float32))
    ds = T(rng.uniform(0.5, 1.5, n).astype(np.float32))
    out = sb.ops.stochastic_aggregate(g, xc, sp, src_scale=ss.cuda(), dst_scale=ds.cuda())
    out.backward(T(gout).cuda())
    # oracle on the emitted raw variates
    lib = _lib.load()
    import ctypes
    raw = torch.empty((1, e, K), device="cuda")
    wbuf = torch.empty((1, e, K), device="cuda")
    nz = sb.ops._fill_noise(None, sb.ops._KIND[kind], K, pa.detach().contiguous(), pb.detach().contiguous(), None,
                            False, False, 0, 5, 6, sp.param_shape)
    _lib.check(lib.stag_noise_emit(ctypes.byref(nz), e, 1, wbuf.data_ptr(), raw.data_ptr(), 0))
    torch.cuda.synchronize()
    eps = raw[0].cpu()
    ao, bo = T(np.asarray(a)).requires_grad_(True), T(np.asarray(b)).requires_grad_(True)
    wo = ref_spmm.reparam_normal(ao, bo, eps) if kind == "normal" else ref_spmm.reparam_uniform(ao, bo, eps)
    wo = wo.expand(e, K)
    xo = T(x).requires_grad_(True)
    oo = ref_spmm.aggregate(T(src), T(dst), n, xo, wo, src_scale=ss, dst_scale=ds)
    oo.backward(T(gout))
    rel = lambda u, v: float((u.cpu() - v).abs().max() / v.abs().max().clamp(min=1e-30))  # noqa: E731
    assert rel(out, oo) < 1e-5
    assert rel(xc.grad, xo.grad) < 1e-5
    # scalar / per-channel gradients are sums of ~E*D (E) signed terms of magnitude ~1: allow the fp32
    # random-walk rounding of such a sum on top of the 1e-5 relative bound
    n_terms = {"scalar": e * D, "channel": e, "edge": D if K > 1 else 1, "edge_channel": 1}[pshape]
    slack = 1e-6 * np.sqrt(n_terms)

    def ok(u, v):
        return float((u.cpu() - v).abs().max()) <= 2e-5 * float(v.abs().max()) + slack
    assert ok(pa.grad, ao.grad), (pa.grad, ao.grad)
    assert ok(pb.grad, bo.grad), (pb.grad, bo.grad)


@pytest.mark.parametrize("kind", ["normal", "uniform"])
@pytest.mark.parametrize("pshape,D,K,relu", [("edge", 128, 128, False), ("edge", 24, 24, True), ("edge", 24, 1, False),
                                              ("edge_channel", 128, 128, False), ("edge_channel", 50, 50, True)])
def test_per_edge_parameter_gradients_for_batched_samples(kind, pshape, D, K, relu):
    """EDGE / EDGE_CHANNEL parameter gradients (the amortised posteriors) of S samples in ONE launch of the
    edge-parallel kernel == autograd through loc + eps * scale on the emitted variates, summed over the samples."""
    import stag_b200 as sb
    n, e, S = 150, 2500, 3
    rng = np.random.default_rng(11)
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    dst[:300] = 3
    src[300:600] = 4
    shape = (e, 1) if pshape == "edge" else (e, K)
    a = (1 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
    b = (0.3 + 0.1 * rng.uniform(size=shape)).astype(np.float32)
    if kind == "uniform":
        a, b = a - 1.0, b + 1.5
    x = rng.standard_normal((S, n, D)).astype(np.float32)
    gout = rng.standard_normal((S, n, D)).astype(np.float32)
    T = torch.from_numpy
    g = sb.Graph(T(src), T(dst), n).to("cuda")
    pa, pb = T(a).cuda().requires_grad_(True), T(b).cuda().requires_grad_(True)
    sp = sb.ops.NoiseSpec(kind, pa, pb, K, e, relu=relu, seed=5, offset=6, n_samples=S, batched=True)
    ss = T(rng.uniform(0.5, 1.5, n).astype(np.float32))
    ds = T(rng.uniform(0.5, 1.5, n).astype(np.float32))
    xc = T(x).cuda().requires_grad_(True)
    n0 = sb._lib.load().stag_launch_count()
    out = sb.ops.stochastic_aggregate(g, xc, sp, src_scale=ss.cuda(), dst_scale=ds.cuda(), n_samples=S)
    out.backward(T(gout).cuda())
    launches = sb._lib.load().stag_launch_count() - n0
    assert launches <= 12, launches          # not one backward launch per sample
    # oracle: per sample, autograd through the reparameterisation on the emitted raw variates
    plain = sb.ops.NoiseSpec(kind, pa.detach(), pb.detach(), K, e, seed=5, offset=6, n_samples=S,
                             generator="boxmuller")   # the stream of the gradient kernels (parameters with gradients)
    lib = sb._lib.load()
    import ctypes
    raw = torch.empty((S, e, K), device="cuda")
    wbuf = torch.empty((S, e, K), device="cuda")
    nz = sb.ops._fill_noise(None, plain.lib_kind, K, pa.detach().contiguous(), pb.detach().contiguous(), None,
                            False, False, 0, 5, 6, plain.param_shape)
    sb._lib.check(lib.stag_noise_emit(ctypes.byref(nz), e, S, wbuf.data_ptr(), raw.data_ptr(), 0))
    torch.cuda.synchronize()
    ao, bo = T(a).double().requires_grad_(True), T(b).double().requires_grad_(True)
    xo = T(x).double().requires_grad_(True)
    outs = []
    for s in range(S):
        eps = raw[s].cpu().double()
        wo = (ao + eps * bo) if kind == "normal" else (ao + eps * (bo - ao))
        wo = wo.expand(e, K)
        if relu:
            wo = wo.relu()
        outs.append(ref_spmm.aggregate(T(src), T(dst), n, xo[s], wo, src_scale=ss.double(), dst_scale=ds.double()))
    oo = torch.stack(outs)
    oo.backward(T(gout).double())
    rel = lambda u, v: float((u.cpu().double() - v).abs().max() / v.abs().max().clamp(min=1e-30))  # noqa: E731
    assert rel(out, oo) < 1e-5
    assert rel(xc.grad, xo.grad) < 1e-5
    n_terms = S * (D if (pshape == "edge" and K > 1) else 1)
    slack = 1e-6 * np.sqrt(n_terms)
    for got, ref in ((pa.grad, ao.grad), (pb.grad, bo.grad)):
        assert float((got.cpu().double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + slack
