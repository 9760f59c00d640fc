"""Device time per launch of the aggregation kernels for the noise variants of the path (arxiv-shaped
graph, S = 16, per-sample operand), as GB/s of algorithmic bytes and fraction of the measured HBM roof."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import stag_b200 as sb
from stag_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
src, dst = bench.synth_graph()
g = sb.Graph(torch.from_numpy(src), torch.from_numpy(dst), bench.N_NODES).to(dev)
st = g._s
csc, _k1 = st.csx(True)
csr, _k2 = st.csx(False)
ss, ds = st.scale(False, "rsqrt"), st.scale(True, "rsqrt")
S, N, D, E = 16, bench.N_NODES, bench.WIDTH, bench.N_EDGES
x = torch.randn(S, N, D, device=dev)
out = torch.empty(S, N, D, device=dev)
dxb = torch.empty(S, N, D, device=dev)
ws = torch.empty(max(lib.stag_spmm_workspace_bytes(ctypes.byref(csc), D, S), lib.stag_spmm_workspace_bytes(ctypes.byref(csr), D, S)),
                 dtype=torch.uint8, device=dev)
peak, _ = bench.peaks()
one = torch.ones(1, device=dev); sg = torch.full((1,), 0.4, device=dev)
onec = torch.ones(D, device=dev); sgc = torch.full((D,), 0.4, device=dev)
pb = torch.full((1,), 0.8, device=dev)
lo = torch.full((1,), 0.3, device=dev); hi = torch.full((1,), 1.7, device=dev)
dp = torch.zeros(2, D, device=dev)
loc_e = torch.ones(E, 1, device=dev); sg_e = torch.full((E, 1), 0.4, device=dev)
dpe = torch.zeros(2, E, device=dev)
stream = torch.cuda.current_stream().cuda_stream

def noise(kind, K, pshape, p0, p1, relu=0, in_norm=0):
    n = _lib.StagNoise()
    n.kind, n.K, n.param_shape, n.relu, n.in_norm, n.sample_base = kind, K, pshape, relu, in_norm, 0
    n.p0, n.p1, n.external = (p0.data_ptr() if p0 is not None else 0), (p1.data_ptr() if p1 is not None else 0), 0
    n.seed, n.offset = 42, 7
    return n

def fwd(nz):
    _lib.check(lib.stag_spmm_fwd(ctypes.byref(csc), x.data_ptr(), D, N * D, D, S, ctypes.byref(nz), ss.data_ptr(), ds.data_ptr(),
                                 out.data_ptr(), D, N * D, 0, ws.data_ptr(), ws.numel(), stream))

def bwd(nz):
    _lib.check(lib.stag_spmm_bwd(ctypes.byref(csr), x.data_ptr(), D, N * D, out.data_ptr(), D, N * D, D, S, ctypes.byref(nz),
                                 ss.data_ptr(), ds.data_ptr(), dxb.data_ptr(), D, N * D, dp[0].data_ptr(), dp[1].data_ptr(), 0,
                                 ws.data_ptr(), ws.numel(), stream))

def bwd_edge(nz):
    _lib.check(lib.stag_spmm_bwd(ctypes.byref(csr), x.data_ptr(), D, N * D, out.data_ptr(), D, N * D, D, S, ctypes.byref(nz),
                                 ss.data_ptr(), ds.data_ptr(), dxb.data_ptr(), D, N * D, dpe[0].data_ptr(), dpe[1].data_ptr(), 0,
                                 ws.data_ptr(), ws.numel(), stream))

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

L = _lib
cases = [
    ("no noise (copy_u, sum)", lambda: fwd(noise(L.NOISE_NONE, D, 0, None, None)), bench.bytes_fwd(S, False)),
    ("Normal, per-edge (K=1), scalar params", lambda: fwd(noise(L.NOISE_NORMAL, 1, L.PARAM_SCALAR, one, sg)), bench.bytes_fwd(S, False)),
    ("Normal, per-channel noise, scalar params [headline]", lambda: fwd(noise(L.NOISE_NORMAL, D, L.PARAM_SCALAR, one, sg)), bench.bytes_fwd(S, False)),
    ("Normal, per-channel noise, per-channel params", lambda: fwd(noise(L.NOISE_NORMAL, D, L.PARAM_CHANNEL, onec, sgc)), bench.bytes_fwd(S, False)),
    ("Normal + relu, scalar params", lambda: fwd(noise(L.NOISE_NORMAL, D, L.PARAM_SCALAR, one, sg, relu=1)), bench.bytes_fwd(S, False)),
    ("Uniform, per-channel noise, scalar params", lambda: fwd(noise(L.NOISE_UNIFORM, D, L.PARAM_SCALAR, lo, hi)), bench.bytes_fwd(S, False)),
    ("Bernoulli, per-channel noise", lambda: fwd(noise(L.NOISE_BERNOULLI, D, L.PARAM_SCALAR, pb, None)), bench.bytes_fwd(S, False)),
    ("Bernoulli + in-norm (arxiv Bernoulli config)", lambda: fwd(noise(L.NOISE_BERNOULLI, D, L.PARAM_SCALAR, pb, None, in_norm=1)), bench.bytes_fwd(S, False)),
    ("backward dX + d(loc,scale), per-channel params (vi)", lambda: bwd(noise(L.NOISE_NORMAL, D, L.PARAM_CHANNEL, onec, sgc)), bench.bytes_bwd(S, True, False)),
    ("backward dX + d(loc,scale), scalar params (vi)", lambda: bwd(noise(L.NOISE_NORMAL, D, L.PARAM_SCALAR, one, sg)), bench.bytes_bwd(S, True, False)),
    ("Normal, per-channel noise, per-edge params [E,1] (amortised re)", lambda: fwd(noise(L.NOISE_NORMAL, D, L.PARAM_EDGE, loc_e, sg_e)), bench.bytes_fwd(S, False) + 8 * E),
    ("backward dX + d(loc,scale)[E,1], 16 samples in one call (amortised re)", lambda: bwd_edge(noise(L.NOISE_NORMAL, D, L.PARAM_EDGE, loc_e, sg_e)), bench.bytes_bwd(S, True, False) + 16 * E),
    ("Normal via tensor-core Hadamard generator (default where eligible)", lambda: fwd(noise(L.NOISE_NORMAL_HADAMARD, D, L.PARAM_SCALAR, one, sg)), bench.bytes_fwd(S, False)),
]
rows = []
for name, fn, nbytes in cases:
    ms = t(fn)
    gbs = nbytes / ms / 1e6
    rows.append({"case": name, "ms_per_launch": ms, "algorithmic_GB_per_s": gbs, "frac_of_hbm_roof": gbs / peak,
                 "GEdge_samples_per_s": E * S / ms / 1e6})
    print("%-58s %7.3f ms  %7.0f GB/s  %5.1f%% of HBM roof  %6.2f GEdge-samples/s" % (name, ms, gbs, 100 * gbs / peak, E * S / ms / 1e6))
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "modes.json"), "w"), indent=1)
