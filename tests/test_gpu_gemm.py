"""GPU: stag_gemm_tcgen05 (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulator) against a float64
matmul: fp32-level accuracy (<= 1e-5 relative, max-norm), fused row scale / bias / relu epilogue, ragged
M / K / N, and gradients."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp(min=1e-30))


@pytest.mark.parametrize("M,K,N", [(1, 8, 16), (127, 16, 7), (128, 32, 16), (129, 50, 40), (1000, 128, 128),
                                   (5000, 128, 40), (300, 1433, 16), (4096, 256, 256), (777, 100, 121),
                                   (64, 9, 16), (2049, 33, 200)])
def test_matches_float64(M, K, N):
    from stag_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).cuda()
    w = torch.randn(K, N, generator=g).cuda()
    out = ops.dense_transform(a, w)
    ref = a.double() @ w.double()
    assert out.shape == (M, N)
    assert rel(out, ref) < 1e-5, rel(out, ref)


def test_epilogue_and_batched_samples():
    from stag_b200 import ops
    g = torch.Generator().manual_seed(0)
    S, N, K, C = 3, 500, 64, 40
    a = torch.randn(S, N, K, generator=g).cuda()
    w = torch.randn(K, C, generator=g).cuda()
    b = torch.randn(C, generator=g).cuda()
    rs = torch.rand(N, generator=g).cuda() + 0.5
    out = ops.dense_transform(a, w, row_scale=rs, bias=b, relu=True)
    ref = torch.relu((a.double() @ w.double()) * rs.double()[None, :, None] + b.double())
    assert out.shape == (S, N, C)
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("M,K,N", [(2708, 1433, 16), (2708, 1433, 7), (500, 999, 33), (1300, 640, 64), (130, 4100, 40)])
def test_split_k_cluster(M, K, N):
    """Few row tiles and a long K (Cora's first layer): K is split over a thread-block cluster and the partial
    accumulators are added in rank 0's shared memory in rank order -- fp32-level accuracy, the fused epilogue
    still applied once, bitwise run-to-run."""
    from stag_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).cuda()
    w = torch.randn(K, N, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    rs = (torch.rand(M, generator=g) + 0.5).cuda()
    out = ops.dense_transform(a, w, row_scale=rs, bias=b, relu=True)
    ref = torch.relu((a.double() @ w.double()) * rs.double()[:, None] + b.double())
    assert rel(out, ref) < 1e-5, rel(out, ref)
    assert torch.equal(out, ops.dense_transform(a, w, row_scale=rs, bias=b, relu=True))
    plain = ops.dense_transform(a, w)
    assert rel(plain, a.double() @ w.double()) < 1e-5


@pytest.mark.parametrize("M,K,N", [(300001, 128, 128), (300001, 128, 40), (250000, 160, 64), (200003, 100, 121),
                                   (190000, 96, 16), (150000, 32, 256), (400000, 64, 7), (169343 * 4, 128, 128)])
def test_persistent_tma_form_many_tiles_per_cta(M, K, N):
    """Row counts far beyond 148 x 128: every CTA of the persistent TMA form walks many row tiles, so its rings (sizes that
    do not divide the K blocks of a tile), both TMEM accumulators and every mbarrier phase wrap many times.  Ragged M,
    K tails filled by the TMA unit, N from one MMA column block to the 256-column maximum; fused epilogue; elementwise
    tolerance against float64; bitwise run-to-run."""
    from stag_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g, device="cuda")
    w = torch.randn(K, N, generator=g, device="cuda")
    b = torch.randn(N, generator=g, device="cuda")
    rs = torch.rand(M, generator=g, device="cuda") + 0.5
    out = ops.dense_transform(a, w, row_scale=rs, bias=b, relu=True)
    ref = torch.relu((a.double() @ w.double()) * rs.double()[:, None] + b.double())
    scale = float(ref.abs().max())
    assert bool(((out.double() - ref).abs() <= 1e-5 * ref.abs() + 2e-6 * scale).all())
    assert torch.equal(out, ops.dense_transform(a, w, row_scale=rs, bias=b, relu=True))
    del out, ref
    plain = ops.dense_transform(a, w)
    ref = a.double() @ w.double()
    assert bool(((plain.double() - ref).abs() <= 1e-5 * ref.abs() + 2e-6 * float(ref.abs().max())).all())


def test_gradients():
    from stag_b200 import ops
    g = torch.Generator().manual_seed(1)
    a = torch.randn(700, 48, generator=g).cuda().requires_grad_(True)
    w = torch.randn(48, 24, generator=g).cuda().requires_grad_(True)
    b = torch.randn(24, generator=g).cuda().requires_grad_(True)
    rs = (torch.rand(700, generator=g) + 0.5).cuda()
    go = torch.randn(700, 24, generator=g).cuda()
    ops.dense_transform(a, w, row_scale=rs, bias=b, relu=True).backward(go)
    a2, w2, b2 = [t.detach().double().requires_grad_(True) for t in (a, w, b)]
    torch.relu((a2 @ w2) * rs.double()[:, None] + b2).backward(go.double())
    assert rel(a.grad, a2.grad) < 1e-5 and rel(w.grad, w2.grad) < 1e-5 and rel(b.grad, b2.grad) < 1e-5


def test_unsupported_width_falls_to_cublas_in_python_only():
    """Nout > 256 is refused by the C entry point (STAG_EUNSUPPORTED) and routed to torch.matmul above it."""
    import ctypes
    from stag_b200 import _lib, ops
    lib = _lib.load()
    a = torch.randn(10, 8).cuda()
    wt = torch.randn(300, 8).cuda()
    out = torch.empty(10, 300).cuda()
    rc = lib.stag_gemm_tcgen05(a.data_ptr(), 8, wt.data_ptr(), 8, 10, 300, 8, 0, 0, 0, out.data_ptr(), 300, 0, 0, 0)
    assert rc == _lib.STAG_EUNSUPPORTED
    assert ops.dense_transform(a, wt.t().contiguous()).shape == (10, 300)
