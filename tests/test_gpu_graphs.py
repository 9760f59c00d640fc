"""GPU: a step captured in a CUDA graph draws FRESH noise on every replay (StagNoise::counter, the device-side
addend to the Philox call counter; stag_b200.random.enable_device_counter / advance_device_counter), and replay i
equals the i-th eager step bit for bit -- forward and transposed pass, both normal generators.  This is what makes
the launch-bound configurations (a 32-molecule batch: ~16 kernels of a few microseconds per step) graph-capturable."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,generator", [(128, "hadamard"), (128, "boxmuller"), (20, None)])
def test_replays_equal_eager_steps(D, generator):
    import stag_b200 as sb
    from stag_b200.ops import NoiseSpec
    dev = torch.device("cuda")
    rng = np.random.default_rng(1)
    N, E, S = 300, 3000, 2
    g = sb.Graph(torch.from_numpy(rng.integers(0, N, E)), torch.from_numpy(rng.integers(0, N, E)), N).to(dev)
    g._s.csx(True), g._s.csx(False)                       # structure is built outside the capture (host syncs)
    x = torch.randn(N, D, device=dev)
    one, sg = torch.ones((), device=dev), torch.full((), 0.4, device=dev)
    ctr = sb.random.enable_device_counter(dev)

    def step():
        outs = []
        for _ in range(2):                                # two "layers": two offsets per step
            spec = NoiseSpec("normal", one, sg, D, E, n_samples=S, batched=True, generator=generator)
            outs.append(sb.ops.stochastic_aggregate(g, x, spec, n_samples=S))
        n = sb.random.advance_device_counter()
        assert n == 2
        return outs[0], outs[1]

    try:
        sb.manual_seed(31)
        ctr.zero_()
        eager = [tuple(t.clone() for t in step()) for _ in range(3)]
        assert not torch.equal(eager[0][0], eager[1][0])  # eager steps differ from one another ...
        sb.manual_seed(31)
        ctr.zero_()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # warm-up outside the capture (allocator, function attributes)
            step()
        torch.cuda.current_stream().wait_stream(side)
        sb.manual_seed(31)
        ctr.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out, dx = step()
        for i in range(3):                                # ... and replay i is eager step i, bit for bit
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, eager[i][0]) and torch.equal(dx, eager[i][1]), i
        assert int(ctr.item()) == 6
    finally:
        sb.random.disable_device_counter()


def test_transposed_pass_regenerates_the_noise_under_a_nonzero_device_counter():
    """<A_w x, y> == <x, A_w^T y> with the device counter at 5: forward and backward add the same word."""
    import stag_b200 as sb
    from stag_b200.ops import NoiseSpec
    dev = torch.device("cuda")
    rng = np.random.default_rng(2)
    N, E, D, S = 400, 5000, 128, 2
    g = sb.Graph(torch.from_numpy(rng.integers(0, N, E)), torch.from_numpy(rng.integers(0, N, E)), N).to(dev)
    ctr = sb.random.enable_device_counter(dev)
    try:
        ctr.fill_(5)
        x = torch.randn(N, D, device=dev, requires_grad=True)
        y = torch.randn(S, N, D, device=dev)
        one, sg = torch.ones((), device=dev), torch.full((), 0.4, device=dev)
        for gen in ("hadamard", "boxmuller"):
            x.grad = None
            spec = NoiseSpec("normal", one, sg, D, E, seed=9, offset=1, n_samples=S, batched=True, generator=gen)
            out = sb.ops.stochastic_aggregate(g, x, spec, n_samples=S)
            out.backward(y)
            lhs, rhs = (out.double() * y.double()).sum(), (x.detach().double() * x.grad.double()).sum()
            assert abs(float(lhs - rhs)) <= 1e-6 * float(out.detach().double().abs().mul(y.double().abs()).sum())
            ctr.zero_()
            other = sb.ops.stochastic_aggregate(g, x.detach(), spec, n_samples=S)
            ctr.fill_(5)
            assert not torch.equal(other, out.detach())       # the counter really enters the stream
            spec5 = NoiseSpec("normal", one, sg, D, E, seed=9, offset=6, n_samples=S, batched=True, generator=gen)
            ctr.zero_()
            assert torch.equal(sb.ops.stochastic_aggregate(g, x.detach(), spec5, n_samples=S), out.detach())  # offset + counter
            ctr.fill_(5)
    finally:
        sb.random.disable_device_counter()
