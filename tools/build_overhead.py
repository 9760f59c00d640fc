"""Time of the on-device CSC + CSR build (stag_csx_build) for minibatch-sized and full-sized graphs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stag_b200 as sb

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for N, E in ((824, 1744), (3308, 6986), (5718, 164106), (169343, 1166243)):
    src = torch.from_numpy(rng.integers(0, N, E)).to(dev)
    dst = torch.from_numpy(rng.integers(0, N, E)).to(dev)
    for _ in range(3):
        g = sb.Graph(src, dst, N); g._s.csx(True); g._s.csx(False)
    torch.cuda.synchronize()
    k = 20
    t0 = time.perf_counter()
    for _ in range(k):
        g = sb.Graph(src, dst, N); g._s.csx(True); g._s.csx(False)
    torch.cuda.synchronize()
    print("N %7d E %8d: CSC + CSR build %8.1f us" % (N, E, (time.perf_counter() - t0) / k * 1e6))
