"""Quick device timing of stag_gemm_tcgen05 against torch.matmul (cuBLAS fp32) at the arxiv layer shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stag_b200 import ops

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def tg(fn, n=20):
    """the same call replayed from a CUDA graph: device time without the host-side call overhead"""
    fn(); torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
    torch.cuda.synchronize()
    return t(g.replay, 5) / n

for M, K, N in [(16 * 169343, 128, 128), (16 * 169343, 128, 40), (169343, 128, 128), (2708, 1433, 16)]:
    a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda"); bias = torch.randn(N, device="cuda")
    ms_ours = t(lambda: ops.dense_transform(a, w, bias=bias, relu=True))
    ms_cublas = t(lambda: torch.relu(a @ w + bias))
    ms_mm = t(lambda: a @ w)
    if M < 200000:
        print("   graph replay: ours %.4f ms, cuBLAS matmul %.4f ms, +bias+relu %.4f ms"
              % (tg(lambda: ops.dense_transform(a, w, bias=bias, relu=True)), tg(lambda: a @ w), tg(lambda: torch.relu(a @ w + bias))))
    gb = (M * K + M * N) * 4 / 1e9
    print("M=%d K=%d N=%d  tcgen05 3xTF32 %.3f ms (%.0f GB/s, %.1f TFLOP/s eff)   cuBLAS fp32 matmul %.3f ms, +bias+relu %.3f ms"
          % (M, K, N, ms_ours, gb / ms_ours * 1e3, 2.0 * M * K * N / ms_ours / 1e9, ms_mm, ms_cublas))
