"""Host and device cost of one fused aggregation call on a small (launch-bound) graph: Cora-shaped layer 2
(N 2708, E 10556, D 16, S 4) through stag_b200.ops.stochastic_aggregate and through the bare C ABI."""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stag_b200 as sb  # noqa: E402
from stag_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
N, E, D, S = 2708, 10556, 16, 4
g0 = torch.Generator().manual_seed(0)
g = sb.Graph(torch.randint(0, N, (E,), generator=g0), torch.randint(0, N, (E,), generator=g0), N).to(dev)
x = torch.randn(S, N, D, device=dev)
one, sg = torch.ones((), device=dev), torch.full((), 0.4, device=dev)
ss, ds = g._s.scale(False, "rsqrt"), g._s.scale(True, "rsqrt")


def call_ops():
    sp = sb.ops.NoiseSpec("normal", one, sg, D, E, n_samples=S, batched=True)
    return sb.ops.stochastic_aggregate(g, x, sp, src_scale=ss, dst_scale=ds, n_samples=S)


lib = _lib.load()
csc, _keep = g._s.csx(True)
out = torch.empty(S, N, D, device=dev)
ws = torch.empty(max(lib.stag_spmm_workspace_bytes(ctypes.byref(csc), D, S), 256), dtype=torch.uint8, device=dev)
nz = _lib.StagNoise()
nz.kind, nz.K, nz.param_shape = _lib.NOISE_NORMAL, D, _lib.PARAM_SCALAR
nz.p0, nz.p1, nz.seed, nz.offset = one.data_ptr(), sg.data_ptr(), 42, 0
stream = torch.cuda.current_stream().cuda_stream


def call_abi():
    nz.offset += 1
    lib.stag_spmm_fwd(ctypes.byref(csc), x.data_ptr(), D, N * D, D, S, ctypes.byref(nz), ss.data_ptr(), ds.data_ptr(),
                      out.data_ptr(), D, N * D, 0, ws.data_ptr(), ws.numel(), stream)


for name, fn in (("ops.stochastic_aggregate", call_ops), ("C ABI stag_spmm_fwd", call_abi)):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    n0 = lib.stag_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(500):
        fn()
    b.record()
    host = (time.perf_counter() - t0) / 500 * 1e6
    torch.cuda.synchronize()
    print("%-28s host issue %.1f us/call, device span %.1f us/call, %d launches/call"
          % (name, host, a.elapsed_time(b) / 500 * 1e3, (lib.stag_launch_count() - n0) // 500))
