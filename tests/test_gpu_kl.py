"""GPU: the sample-based KL fallback of StagLayer.kl_divergence (stag/layers.py:139-141) evaluated by stag_noise_kl from
the regenerated variates, against the reference's own expression
    q.log_prob(w).sum(-1).mean() - p.log_prob(w).sum(-1).mean()
on the noise tensor emitted from the same Philox stream (values and gradients w.r.t. the parameters of q), for every
parameter shape class, Normal / Uniform posteriors, relu, Normal and mixture priors (the mixture of
scripts/citation_rec_contrastive/gcn/run.py:44-52) -- and without any [E,K] allocation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
td = torch.distributions


def mixture(dev, std=0.3):
    return td.MixtureSameFamily(
        mixture_distribution=td.Categorical(torch.ones(2, device=dev)),
        component_distribution=td.Normal(torch.tensor([0.0, 1.0], device=dev), torch.tensor([std, std], device=dev)))


def reference_expression(q, p, w):
    return q.log_prob(w).sum(dim=-1).mean() - p.log_prob(w).sum(dim=-1).mean()


@pytest.mark.parametrize("kind", ["normal", "uniform"])
@pytest.mark.parametrize("shape", ["scalar", "channel", "edge", "edge_channel"])
@pytest.mark.parametrize("K,S,relu,prior", [(16, 1, False, "mix"), (100, 3, False, "mix"), (256, 2, True, "normal"),
                                            (1433, 1, False, "mix")])
def test_fused_kl_equals_the_reference_expression_on_the_emitted_sample(kind, shape, K, S, relu, prior):
    from stag_b200 import ops
    if relu and kind == "uniform":
        pytest.skip("relu never clips a Uniform(0.3, 1.7) sample")
    dev = torch.device("cuda")
    E = 257 if K > 1000 else 1500
    g = torch.Generator(device=dev).manual_seed(K + S)
    pshape = {"scalar": (), "channel": (K,), "edge": (E, 1), "edge_channel": (E, K)}[shape]
    if kind == "normal":
        p0 = (1.0 + 0.2 * torch.randn(pshape, device=dev, generator=g)).requires_grad_(True)
        p1 = (0.2 + 0.3 * torch.rand(pshape, device=dev, generator=g)).requires_grad_(True)
    else:
        p0 = (0.3 + 0.1 * torch.rand(pshape, device=dev, generator=g)).requires_grad_(True)
        p1 = (1.5 + 0.2 * torch.rand(pshape, device=dev, generator=g)).requires_grad_(True)
    p = mixture(dev) if prior == "mix" else td.Normal(torch.tensor(0.5, device=dev), torch.tensor(0.7, device=dev))
    spec = ops.NoiseSpec(kind, p0, p1, K, E, relu=relu, seed=17, offset=3, sample_base=2, n_samples=S, batched=True)
    kl = ops.fused_kl_fallback(spec, ops.describe_prior(p))
    g0, g1 = torch.autograd.grad(kl, (p0, p1))
    # the reference's expression on the emitted tensor, in double precision
    a0 = p0.detach().double().requires_grad_(True)
    a1 = p1.detach().double().requires_grad_(True)
    eps_spec = ops.NoiseSpec(kind, torch.zeros((), device=dev), torch.ones((), device=dev), K, E, seed=17, offset=3,
                             sample_base=2, n_samples=S, batched=True, generator="boxmuller")
    raw = eps_spec.materialize(n_samples=S).double()      # Normal(0,1): eps; Uniform(0,1): u
    if kind == "normal":
        w = a0 + a1 * raw
        q = td.Normal(a0, a1, validate_args=False)
    else:
        w = a0 + (a1 - a0) * raw
        q = td.Uniform(a0, a1, validate_args=False)
    if relu:
        w = w.relu()
    pd = (td.MixtureSameFamily(td.Categorical(torch.ones(2, device=dev, dtype=torch.float64)),
                               td.Normal(torch.tensor([0.0, 1.0], device=dev, dtype=torch.float64),
                                         torch.tensor([0.3, 0.3], device=dev, dtype=torch.float64)))
          if prior == "mix" else td.Normal(torch.tensor(0.5, device=dev, dtype=torch.float64),
                                           torch.tensor(0.7, device=dev, dtype=torch.float64)))
    ref = reference_expression(q, pd, w)
    r0, r1 = torch.autograd.grad(ref, (a0, a1))
    assert abs(float(kl) - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
    for got, want in ((g0, r0), (g1, r1)):
        scale = float(want.abs().max())
        assert float((got.double() - want).abs().max()) <= 2e-4 * scale + 1e-9


def test_layer_fallback_allocates_no_noise_tensor():
    """StagLayer(GCN), vi=True, mixture prior: kl_divergence() takes the fused route (no [E,K] tensor is ever
    allocated: peak memory stays far below E*K*4 bytes) and equals the tensor route of the same layer."""
    import stag_b200 as stag
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    N, E, D = 2000, 200000, 128
    g = stag.Graph(torch.from_numpy(rng.integers(0, N, E)), torch.from_numpy(rng.integers(0, N, E)), N).to(dev)
    x = torch.randn(N, D, device=dev)
    layer = stag.layers.StagLayer(stag.zoo.GCN(D, 16), q_a=td.Normal(torch.ones(D), 0.3 * torch.ones(D)),
                                  p_a=mixture("cpu"), vi=True).to(dev)
    layer.p_a = mixture(dev)
    layer.p_a.base_distribution = layer.p_a
    stag.manual_seed(11)
    out = layer(g, x)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    kl = layer.kl_divergence()
    (out.sum() * 0 + kl).backward(retain_graph=True)
    torch.cuda.synchronize()
    assert torch.cuda.max_memory_allocated() - base < E * D * 4 // 8, "an [E,K]-sized tensor was allocated"
    g_fused = {k: p.grad.clone() for k, p in layer.q_a.named_parameters()}
    # tensor route of the same forward: materialise the sample, evaluate the reference's expression with torch
    for p_ in layer.parameters():
        p_.grad = None
    w = layer._edge_weight_sample
    kl_t = layer.q_a.log_prob(w).sum(dim=-1).mean() - layer.p_a.log_prob(w).sum(dim=-1).mean()
    kl_t.backward()
    assert abs(float(kl) - float(kl_t)) <= 2e-5 * abs(float(kl_t))
    for k, p_ in layer.q_a.named_parameters():
        assert float((g_fused[k] - p_.grad).abs().max()) <= 2e-4 * float(p_.grad.abs().max()) + 1e-9


def test_unsupported_priors_take_the_tensor_route():
    from stag_b200 import ops
    assert ops.describe_prior(td.Uniform(0.0, 2.0)) is None
    assert ops.describe_prior(td.Normal(torch.zeros(3), torch.ones(3))) is None
    assert ops.describe_prior(td.Normal(torch.tensor(0.0, requires_grad=True), torch.tensor(1.0))) is None
    w, loc, sc = ops.describe_prior(mixture("cpu"))
    assert w == [0.5, 0.5] and loc == [0.0, 1.0] and len(sc) == 2
