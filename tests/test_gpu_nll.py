"""GPU: the fused likelihood epilogue (stag_nll) against torch.distributions, which is what the reference evaluates
(stag/likelihoods.py:13-38, stag/models.py:69-72): values and gradients, masks, clamped probabilities, sizes of
the named configurations (arxiv 40 classes, PPI 121 labels, Cora 7)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_nll(feats, y, mask, family):
    out = []
    for s in range(feats.shape[0]):
        v = -family(probs=feats[s], validate_args=False).log_prob(y)
        out.append((v if mask is None else v[mask]).mean())
    return torch.stack(out)


@pytest.mark.parametrize("S,N,C", [(1, 50, 7), (3, 1000, 40), (2, 777, 121), (4, 33, 1), (2, 5000, 33)])
@pytest.mark.parametrize("masked", [False, True])
def test_categorical(S, N, C, masked):
    from stag_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(S * N + C)
    feats = torch.rand((S, N, C), device="cuda", generator=g) * 3.0        # unnormalised, as a softmax-free head gives
    if C > 1:
        feats[0, :5, 0] = 0.0                                                # clamped terms (p_y = 0 for some rows)
    y = torch.randint(0, C, (N,), device="cuda", generator=g)
    y[:3] = 0
    mask = (torch.rand(N, device="cuda", generator=g) < 0.6) if masked else None
    a = feats.clone().requires_grad_(True)
    b = feats.clone().requires_grad_(True)
    out = ops.fused_nll(a, y, mask, "categorical")
    ref = reference_nll(b, y, mask, torch.distributions.Categorical)
    assert torch.allclose(out, ref, rtol=2e-6, atol=1e-6)
    w = torch.rand(S, device="cuda", generator=g) + 0.5
    (out * w).sum().backward()
    (ref * w).sum().backward()
    scale = b.grad.abs().max().clamp(min=1e-30)
    assert float((a.grad - b.grad).abs().max() / scale) < 1e-5


@pytest.mark.parametrize("S,N,C", [(1, 50, 7), (3, 1000, 121), (2, 4097, 16)])
@pytest.mark.parametrize("masked", [False, True])
def test_bernoulli(S, N, C, masked):
    from stag_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(S * N + C)
    feats = torch.rand((S, N, C), device="cuda", generator=g)
    feats[0, 0, :2] = torch.tensor([0.0, 1.0], device="cuda")               # clamped
    y = (torch.rand((N, C), device="cuda", generator=g) < 0.3).float()
    mask = (torch.rand(N, device="cuda", generator=g) < 0.5) if masked else None
    a = feats.clone().requires_grad_(True)
    b = feats.clone().requires_grad_(True)
    out = ops.fused_nll(a, y, mask, "bernoulli")
    ref = reference_nll(b, y, mask, torch.distributions.Bernoulli)
    assert torch.allclose(out, ref, rtol=5e-6, atol=1e-6)
    out.sum().backward()
    ref.sum().backward()
    scale = b.grad.abs().max().clamp(min=1e-30)
    assert float((a.grad - b.grad).abs().max() / scale) < 1e-5


def test_empty_mask_is_nan_and_model_loss_uses_the_fused_epilogue():
    import stag_b200 as stag
    from stag_b200 import _lib
    feats = torch.rand((2, 10, 4), device="cuda")
    y = torch.randint(0, 4, (10,), device="cuda")
    out = stag.ops.fused_nll(feats, y, torch.zeros(10, dtype=torch.bool, device="cuda"), "categorical")
    assert torch.isnan(out).all()
    # StagModel.loss_terms: same value through the fused epilogue and through torch.distributions
    g = stag.rand_graph(60, 400).to("cuda")
    layers = [stag.layers.StagLayer(stag.zoo.GCN(16, 16, activation=torch.relu)),
              stag.layers.StagLayer(stag.zoo.GCN(16, 5, activation=lambda x: x.softmax(-1)))]
    model = stag.models.StagModel(layers).cuda()
    for layer in layers:
        layer.cuda()
    x = torch.randn(60, 16, device="cuda")
    yy = torch.randint(0, 5, (60,), device="cuda")
    mask = torch.rand(60, device="cuda") < 0.5
    stag.manual_seed(3)
    n0 = _lib.load().stag_launch_count()
    nll, _ = model.loss_terms(g, x, yy, mask=mask, n_samples=4)
    assert _lib.load().stag_launch_count() > n0
    stag.manual_seed(3)
    outs = model._forward_samples(g, x, 4)
    ref = reference_nll(outs, yy, mask, torch.distributions.Categorical).mean()
    assert torch.allclose(nll, ref, rtol=1e-5)
