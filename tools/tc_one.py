"""A few launches of the tensor-core noise kernel at the arxiv shape (ncu target)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
src, dst = bench.synth_graph()
path = bench.Path(dev, src, dst, 16, 0, False, normal=(sys.argv[1] if len(sys.argv) > 1 else "hadamard"))
for layer in range(3):
    path.fwd(layer)
torch.cuda.synchronize()
print("ok")
