import torch
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
N, D, H = 1224515, 100, 1166755
dx = torch.randn(N, D, device="cuda"); back = torch.randn(H, D, device="cuda")
idx = torch.randperm(N, device="cuda")[:H].sort().values
ext = torch.randn(N + H, D, device="cuda")
print("clone own      %.3f ms" % t(lambda: ext[:N].clone()))
print("contiguous halo %.3f ms" % t(lambda: ext[N:].contiguous()))
print("index_add_     %.3f ms" % t(lambda: dx.index_add_(0, idx, back)))
print("index_select   %.3f ms" % t(lambda: dx.index_select(0, idx)))
g = dx.index_select(0, idx)
print("add_           %.3f ms" % t(lambda: g.add_(back)))
print("index_copy_    %.3f ms" % t(lambda: dx.index_copy_(0, idx, g)))
print("index_put      %.3f ms" % t(lambda: dx.index_put_((idx,), g)))
def f(): dx[idx] += back
print("dx[idx] += back %.3f ms" % t(f))
