"""Device time per launch of the headline aggregation with the tensor-core normal generator (agg_wh_quad_kernel), per
kernel variant (STAG_WQ_VARIANT is read once per process: one process per variant), next to the Box-Muller kernel.
Usage: time_wq.py [shared | csr]"""
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
shared = len(sys.argv) > 1 and sys.argv[1] == "shared"
transposed = len(sys.argv) > 1 and sys.argv[1] == "csr"
csc, _k1 = st.csx(not transposed)
ss, ds = st.scale(False, "rsqrt"), st.scale(True, "rsqrt")
if transposed:
    ss, ds = ds, ss
S, N, D, E = 16, bench.N_NODES, bench.WIDTH, bench.N_EDGES
x = torch.randn((N, D) if shared else (S, N, D), device=dev)
out = torch.empty(S, N, D, device=dev)
ws = torch.empty(lib.stag_spmm_workspace_bytes(ctypes.byref(csc), D, S), dtype=torch.uint8, device=dev)
one = torch.ones(1, device=dev); sg = torch.full((1,), 0.4, device=dev)
stream = torch.cuda.current_stream().cuda_stream

def noise(kind):
    n = _lib.StagNoise()
    n.kind, n.K, n.param_shape, n.relu, n.in_norm, n.sample_base = kind, D, _lib.PARAM_SCALAR, 0, 0, 0
    n.p0, n.p1, n.external = one.data_ptr(), sg.data_ptr(), 0
    n.seed, n.offset = 42, 7
    return n

def fwd(nz):
    _lib.check(lib.stag_spmm_fwd(ctypes.byref(csc), x.data_ptr(), D, 0 if shared else N * D, D, S, ctypes.byref(nz),
                                 ss.data_ptr(), ds.data_ptr(), out.data_ptr(), D, N * D, 0, ws.data_ptr(), ws.numel(), stream))

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

tag = "variant=%s %s" % (os.environ.get("STAG_WQ_VARIANT", "0"), sys.argv[1] if len(sys.argv) > 1 else "per-sample")
th = t(lambda: fwd(noise(_lib.NOISE_NORMAL_HADAMARD)))
ch = float(out.double().abs().sum())
tb = t(lambda: fwd(noise(_lib.NOISE_NORMAL)))
print("%s: hadamard %.3f ms (checksum %.6e)  boxmuller %.3f ms" % (tag, th, ch, tb))
