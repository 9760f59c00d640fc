"""Device time of the Bernoulli + in-norm forward (arxiv shape, S = 16) for A/B runs of library variants."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import stag_b200 as sb
from stag_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
src, dst = bench.synth_graph()
g = sb.Graph(torch.from_numpy(src), torch.from_numpy(dst), bench.N_NODES).to(dev)
st = g._s
csc, _k1 = st.csx(True)
ss, ds = st.scale(False, "rsqrt"), st.scale(True, "rsqrt")
S, N, D = 16, bench.N_NODES, bench.WIDTH
x = torch.randn(S, N, D, device=dev)
out = torch.empty(S, N, D, device=dev)
nsc = torch.empty(S, N, D, device=dev)
ws = torch.empty(lib.stag_spmm_workspace_bytes(ctypes.byref(csc), D, S), dtype=torch.uint8, device=dev)
pb = torch.full((1,), 0.8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
n = _lib.StagNoise()
n.kind, n.K, n.param_shape, n.relu, n.in_norm, n.sample_base = _lib.NOISE_BERNOULLI, D, _lib.PARAM_SCALAR, 0, 1, 0
n.p0, n.p1, n.external, n.seed, n.offset = pb.data_ptr(), 0, 0, 42, 7

def fwd(ns):
    _lib.check(lib.stag_spmm_fwd(ctypes.byref(csc), x.data_ptr(), D, N * D, D, S, ctypes.byref(n), ss.data_ptr(), ds.data_ptr(),
                                 out.data_ptr(), D, N * D, ns, ws.data_ptr(), ws.numel(), stream))

def t(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k

print("%s in-norm %.3f ms (with scale output %.3f ms)" % (os.environ.get("STAG_B200_LIB", "default"), t(lambda: fwd(0)), t(lambda: fwd(nsc.data_ptr()))))
