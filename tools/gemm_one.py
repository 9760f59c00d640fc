import sys, os
sys.path.insert(0, os.getcwd())
import torch
from stag_b200 import ops
M, K, N = 16 * 169343, 128, int(os.environ.get("GN", "128"))
a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda"); bias = torch.randn(N, device="cuda")
for _ in range(3): ops.dense_transform(a, w, bias=bias, relu=True)
torch.cuda.synchronize()
