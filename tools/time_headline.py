"""Device time per launch of the headline aggregation (arxiv shape, S = 16, Normal per-channel noise, scalar
parameters) and of the Uniform variant, for A/B runs of library variants (STAG_B200_LIB=...)."""
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
S, N, D, E = 16, bench.N_NODES, bench.WIDTH, bench.N_EDGES
x = torch.randn(S, N, D, device=dev)
out = torch.empty(S, N, D, device=dev)
ws = torch.empty(lib.stag_spmm_workspace_bytes(ctypes.byref(csc), D, S), dtype=torch.uint8, device=dev)
one = torch.ones(1, device=dev); sg = torch.full((1,), 0.4, device=dev)
lo = torch.full((1,), 0.3, device=dev); hi = torch.full((1,), 1.7, device=dev)
stream = torch.cuda.current_stream().cuda_stream

def noise(kind, p0, p1):
    n = _lib.StagNoise()
    n.kind, n.K, n.param_shape, n.relu, n.in_norm, n.sample_base = kind, D, _lib.PARAM_SCALAR, 0, 0, 0
    n.p0, n.p1, n.external = p0.data_ptr(), p1.data_ptr(), 0
    n.seed, n.offset = 42, 7
    return n

XSS = N * D

def fwd(nz):
    _lib.check(lib.stag_spmm_fwd(ctypes.byref(csc), x.data_ptr(), D, XSS, D, S, ctypes.byref(nz), ss.data_ptr(), ds.data_ptr(),
                                 out.data_ptr(), D, N * D, 0, ws.data_ptr(), ws.numel(), stream))

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

tag = os.environ.get("STAG_B200_LIB", "default")
if os.environ.get("SHARED_X"):
    XSS = 0
    tag += " shared-x"
print("%s normal %.3f ms uniform %.3f ms checksum %.6e" % (
    tag, t(lambda: fwd(noise(_lib.NOISE_NORMAL, one, sg))), t(lambda: fwd(noise(_lib.NOISE_UNIFORM, lo, hi))),
    float(out.double().abs().sum())))
