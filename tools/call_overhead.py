"""Per-call overhead of the aggregation on a launch-bound (molhiv-batch-sized) graph: raw C-ABI call vs the
autograd operator, host time and device time."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stag_b200 as sb
from stag_b200 import _lib
from stag_b200.ops import NoiseSpec

lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
N, E, D, S = 824, 1744, 16, 4
src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
g = sb.Graph(torch.from_numpy(src), torch.from_numpy(dst), N).to(dev)
st = g._s
csc, _k = st.csx(True)
csr, _k2 = st.csx(False)
x = torch.randn(S, N, D, device=dev)
out = torch.empty(S, N, D, device=dev)
ws = torch.empty(max(lib.stag_spmm_workspace_bytes(ctypes.byref(csc), D, S), 256), dtype=torch.uint8, device=dev)
one = torch.ones(1, device=dev); sg = torch.full((1,), 0.4, device=dev)
stream = torch.cuda.current_stream().cuda_stream
n = _lib.StagNoise()
n.kind, n.K, n.param_shape, n.relu, n.in_norm, n.sample_base = _lib.NOISE_NORMAL, D, _lib.PARAM_SCALAR, 0, 0, 0
n.p0, n.p1, n.external, n.seed, n.offset = one.data_ptr(), sg.data_ptr(), 0, 1, 2

def raw():
    _lib.check(lib.stag_spmm_fwd(ctypes.byref(csc), x.data_ptr(), D, N * D, D, S, ctypes.byref(n), 0, 0,
                                 out.data_ptr(), D, N * D, 0, ws.data_ptr(), ws.numel(), stream))

xs = x.clone().requires_grad_(True)
loc, scale = torch.ones((), device=dev), torch.full((), 0.4, device=dev)
gout = torch.randn(S, N, D, device=dev)

def op_fwd():
    spec = NoiseSpec("normal", loc, scale, D, E, n_samples=S, batched=True)
    return sb.ops.stochastic_aggregate(g, xs, spec, n_samples=S)

def op_fwd_bwd():
    o = op_fwd()
    o.backward(gout)
    xs.grad = None

def bench(fn, k=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(k): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / k * 1e6, a.elapsed_time(b) / k * 1e3

for name, fn in (("raw stag_spmm_fwd", raw), ("operator forward", op_fwd), ("operator forward + backward", op_fwd_bwd)):
    h, d = bench(fn)
    print("%-30s host %7.1f us/call   device-elapsed %7.1f us/call" % (name, h, d))
print("launches per raw call:", end=" ")
c0 = lib.stag_launch_count(); raw(); print(lib.stag_launch_count() - c0)
