"""GPU: the tensor-core normal generator (STAG_NOISE_NORMAL_HADAMARD, csrc/spmm_tc.cuh + spmm_wq.cuh).
(1) the emitted stream equals the numpy restatement BIT FOR BIT (every sum is exact in fp32);
(2) the tensor-core sums the fused kernel consumes equal the emitted stream bit for bit;
(3) the law: moments, Kolmogorov-Smirnov, independence across channel / edge / sample / offset;
(4) fused forward and transposed pass == oracle fed the emitted tensor, 1e-5 relative (north_star);
(5) the combinations the fused kernel does not take are refused, never silently re-routed."""
import ctypes

import numpy as np
import pytest
import torch
from scipy import stats

from oracle import ref_philox, ref_spmm

pytestmark = pytest.mark.gpu


def spec(p0, p1, K, E, **kw):
    from stag_b200.ops import NoiseSpec
    t = lambda v: torch.as_tensor(v, dtype=torch.float32).cuda()  # noqa: E731
    return NoiseSpec("normal", t(p0), t(p1), K, E, generator="hadamard", **kw)


@pytest.mark.parametrize("K", [128, 256, 384])
def test_emit_is_bit_exact_against_numpy(K):
    E, seed, off = 333, 0xDEADBEEFCAFE, (7 << 32) + 5
    for s_base in (0, 3):
        w = spec(1.0, 0.5, K, E, seed=seed, offset=off, sample_base=s_base).materialize(n_samples=2).cpu().numpy()
        for s in range(2):
            assert np.array_equal(w[s], ref_philox.noise("normal_hadamard", E, K, s_base + s, seed, off, 1.0, 0.5))
    rng = np.random.default_rng(0)
    loc_e = rng.normal(1, 0.2, (E, 1)).astype(np.float32)
    scale_e = rng.uniform(0.1, 0.5, (E, 1)).astype(np.float32)
    w = spec(loc_e, scale_e, K, E, seed=1, offset=2).materialize().cpu().numpy()
    np.testing.assert_allclose(w, ref_philox.noise("normal_hadamard", E, K, 0, 1, 2, loc_e, scale_e), rtol=0, atol=2e-7)


@pytest.mark.parametrize("K", [100, 97, 250])
def test_widths_within_a_third_of_a_group_run_padded(K):
    """K = 100 (products) runs at 128: the emitted stream is the first K columns of the 128-wide one, and 'auto' picks
    the tensor-core generator there."""
    from stag_b200 import _lib
    from stag_b200.ops import NoiseSpec
    E, Kp = 1 << 18, (K + 127) // 128 * 128
    assert spec(1.0, 0.5, K, E - 1).hadamard_width == 0          # short passes do not pay for the pad / slice launches
    sp = spec(1.0, 0.5, K, E, seed=5, offset=6)
    assert sp.hadamard_width == Kp and sp.lib_kind == _lib.NOISE_NORMAL_HADAMARD
    one = torch.ones((), device="cuda")
    assert NoiseSpec("normal", one, one, K, E).lib_kind == _lib.NOISE_NORMAL_HADAMARD          # generator=None -> auto
    assert NoiseSpec("normal", one, one, K, E, generator="boxmuller").lib_kind == _lib.NOISE_NORMAL
    w = sp.materialize(n_samples=2)[:, :4097].cpu().numpy()
    for s in range(2):
        assert np.array_equal(w[s], ref_philox.noise("normal_hadamard", 4097, Kp, s, 5, 6, 1.0, 0.5)[:, :K])


def test_tensor_core_sums_equal_the_emitted_stream_bitwise():
    """One in-edge per node, x = 1, loc = 0, scale = 1, no degree scales: out[v] is the noise row of v's edge."""
    import stag_b200 as sb
    N, D, S = 1000, 256, 2
    src, dst = torch.randperm(N), torch.arange(N)
    g = sb.Graph(src, dst, N).to("cuda")
    sp = spec(0.0, 1.0, D, N, seed=3, offset=9, n_samples=S, batched=True)
    out = sb.ops.stochastic_aggregate(g, torch.ones(N, D, device="cuda"), sp, n_samples=S)
    assert torch.equal(out, sp.materialize(n_samples=S))


def test_moments_ks_independence():
    E, K = 10000, 128   # 1.28e6 draws
    n = E * K
    a = spec(1.0, 0.4, K, E, seed=42, offset=0).materialize(n_samples=2).cpu().numpy().astype(np.float64)
    z = (a[0] - 1.0) / 0.4
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2.0 / n)
    assert abs(stats.skew(z.ravel())) < 5 * np.sqrt(6.0 / n)
    assert abs(stats.kurtosis(z.ravel())) < 5 * np.sqrt(24.0 / n)
    assert stats.kstest(z.ravel(), "norm").pvalue > 1e-3
    assert np.all(np.abs(z.mean(0)) < 5 / np.sqrt(E))
    assert np.all(np.abs(z.var(0) - 1) < 5 * np.sqrt(2.0 / E))
    assert 30 < (np.abs(z) > 4).sum() < 160          # 81 expected
    # per-channel law (each channel sees E independent byte rows)
    for c in (0, 1, 64, 127):
        assert stats.kstest(z[:, c], "norm").pvalue > 1e-4
    z1 = (a[1] - 1.0) / 0.4
    b = (spec(1.0, 0.4, K, E, seed=42, offset=1).materialize().cpu().numpy().astype(np.float64) - 1.0) / 0.4
    lim = 5 / np.sqrt(n)
    corr = lambda x, y: np.corrcoef(x.ravel(), y.ravel())[0, 1]  # noqa: E731
    assert abs(corr(z, z1)) < lim and abs(corr(z, b)) < lim
    assert abs(corr(z[:-1], z[1:])) < lim
    # channels of one group share their 128 bytes: uncorrelated (H is orthogonal), and so are their squares
    # (the input law has excess kurtosis 8e-4, which is what every fourth-order cross-cumulant scales with)
    C = np.corrcoef(z.T)
    assert np.abs(C - np.eye(K)).max() < 6 / np.sqrt(E)
    C2 = np.corrcoef((z ** 2).T)
    assert np.abs(C2 - np.eye(K)).max() < 6 / np.sqrt(E)


@pytest.mark.parametrize("N,E,D,S,shared,hub", [
    (500, 3000, 128, 3, True, 0), (500, 3000, 128, 3, False, 0), (3000, 40000, 256, 2, False, 9000),
    (300, 100, 128, 1, True, 0), (700, 9000, 384, 2, True, 2500), (5000, 70001, 128, 5, False, 300),
    (5000, 1 << 18, 100, 2, True, 0), (8000, 300000, 200, 1, False, 300)])   # padded to the next group (graphs of >= 2^18 edges)
def test_fused_forward_and_transposed_pass_consume_the_emitted_noise(N, E, D, S, shared, hub):
    import stag_b200 as sb
    rng = np.random.default_rng(N + E)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    if hub:
        dst[:hub] = 7
        src[hub:2 * hub] = 11
    T = torch.from_numpy
    g = sb.Graph(T(src), T(dst), N).to("cuda")
    x = rng.standard_normal((N, D) if shared else (S, N, D)).astype(np.float32)
    ss = T(rng.uniform(0.5, 1.5, N).astype(np.float32))
    ds = T(rng.uniform(0.5, 1.5, N).astype(np.float32))
    gout = rng.standard_normal((S, N, D)).astype(np.float32)
    sp = spec(1.0, 0.4, D, E, seed=11, offset=5, n_samples=S, batched=True)
    xc = T(x).cuda().requires_grad_(True)
    out = sb.ops.stochastic_aggregate(g, xc, sp, src_scale=ss.cuda(), dst_scale=ds.cuda(), n_samples=S)
    out.backward(T(gout).cuda())
    w = sp.materialize(n_samples=S).cpu().double()
    xo = T(x).double().requires_grad_(True)
    oo = torch.stack([ref_spmm.aggregate(T(src), T(dst), N, xo if shared else xo[s], w[s],
                                         src_scale=ss.double(), dst_scale=ds.double()) for s in range(S)])
    oo.backward(T(gout).double())
    rel = lambda u, v: float((u.cpu().double() - v).abs().max() / v.abs().max())  # noqa: E731
    assert rel(out, oo) < 1e-5
    assert rel(xc.grad, xo.grad) < 1e-5


def test_per_edge_parameters():
    import stag_b200 as sb
    N, E, D = 400, 5000, 128
    rng = np.random.default_rng(4)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    T = torch.from_numpy
    g = sb.Graph(T(src), T(dst), N).to("cuda")
    loc = rng.normal(1, 0.2, (E, 1)).astype(np.float32)
    scale = rng.uniform(0.1, 0.5, (E, 1)).astype(np.float32)
    sp = spec(loc, scale, D, E, seed=2, offset=3)
    x = T(rng.standard_normal((N, D)).astype(np.float32))
    out = sb.ops.stochastic_aggregate(g, x.cuda(), sp)
    oo = ref_spmm.aggregate(T(src), T(dst), N, x.double(), sp.materialize().cpu().double())
    assert float((out.cpu().double() - oo).abs().max() / oo.abs().max()) < 1e-5


def test_unsupported_combinations_are_refused():
    import stag_b200 as sb
    from stag_b200 import _lib
    from stag_b200.ops import NoiseSpec
    one = torch.ones((), device="cuda")
    for kw in (dict(K=64), dict(K=130), dict(K=128, relu=True), dict(K=128, in_norm=True)):
        K = kw.pop("K")
        with pytest.raises(ValueError):
            NoiseSpec("normal", one, one, K, 10, generator="hadamard", **kw).lib_kind
    with pytest.raises(ValueError):   # parameter gradients need the Box-Muller two-sum kernel
        NoiseSpec("normal", one.clone().requires_grad_(True), one, 128, 10, generator="hadamard").lib_kind
    # at the C ABI: a width the kernel does not take is STAG_EUNSUPPORTED / STAG_EINVAL, never another kernel
    lib = _lib.load()
    g = sb.Graph(torch.tensor([0, 1]), torch.tensor([1, 0]), 2).to("cuda")
    csc, _keep = g._s.csx(True)
    nz = sb.ops._fill_noise(None, _lib.NOISE_NORMAL_HADAMARD, 64, one.reshape(1), one.reshape(1), None,
                            False, False, 0, 1, 2, _lib.PARAM_SCALAR)
    x = torch.ones(2, 64, device="cuda")
    out = torch.empty(1, 2, 64, device="cuda")
    ws = torch.empty(max(lib.stag_spmm_workspace_bytes(ctypes.byref(csc), 64, 1), 256), dtype=torch.uint8, device="cuda")
    rc = lib.stag_spmm_fwd(ctypes.byref(csc), x.data_ptr(), 64, 0, 64, 1, ctypes.byref(nz), 0, 0, out.data_ptr(),
                           64, 128, 0, ws.data_ptr(), ws.numel(), 0)
    assert rc in (_lib.STAG_EINVAL, _lib.STAG_EUNSUPPORTED)
