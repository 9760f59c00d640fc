"""TEST INFRASTRUCTURE ONLY (see dgl/__init__.py).  dgl.function builtins the
reference uses: stag/layers.py:13-14, stag/zoo/gcn.py:59,63,95,
stag/zoo/graph_sage.py:53,57,72,86,94, stag/zoo/gat.py:114,125-126,
stag/zoo/gated_gcn.py:30-36,42."""
import torch


def _bcast(u, e):
    # DGL broadcasts message operands over trailing dims ((E,D) x (E,1), (E,D) x (E,)).
    while e.dim() < u.dim():
        e = e.unsqueeze(-1)
    while u.dim() < e.dim():
        u = u.unsqueeze(-1)
    return u, e


class _Message:
    def __init__(self, kind, a, b, out):
        self.kind, self.a, self.b, self.out = kind, a, b, out

    def __call__(self, g):
        if self.kind == "copy_e":
            return g.edata[self.a]
        if self.kind == "copy_u":
            return g.ndata[self.a][g._src]
        if self.kind == "u_mul_e":
            u, e = _bcast(g.ndata[self.a][g._src], g.edata[self.b])
            return u * e
        if self.kind == "u_add_v":
            return g.ndata[self.a][g._src] + g.ndata[self.b][g._dst]
        raise NotImplementedError(self.kind)


class _Reduce:
    def __init__(self, kind, msg, out):
        self.kind, self.msg, self.out = kind, msg, out

    def __call__(self, g, m):
        n = g.number_of_nodes()
        if self.kind in ("sum", "mean"):
            out = torch.zeros((n,) + m.shape[1:], dtype=m.dtype, device=m.device)
            out = out.index_add(0, g._dst, m)
            if self.kind == "mean":
                deg = g.in_degrees().to(m).clamp(min=1)
                out = out / deg.reshape((-1,) + (1,) * (m.dim() - 1))
            return out
        if self.kind == "max":
            out = torch.full((n,) + m.shape[1:], float("-inf"), dtype=m.dtype, device=m.device)
            idx = g._dst.reshape((-1,) + (1,) * (m.dim() - 1)).expand_as(m)
            out = out.scatter_reduce(0, idx, m, reduce="amax", include_self=True)
            return torch.where(torch.isinf(out), torch.zeros_like(out), out)
        raise NotImplementedError(self.kind)


def copy_edge(e, out):
    return _Message("copy_e", e, None, out)


copy_e = copy_edge


def copy_src(u, out):
    return _Message("copy_u", u, None, out)


copy_u = copy_src


def u_mul_e(u, e, out):
    return _Message("u_mul_e", u, e, out)


def u_add_v(u, v, out):
    return _Message("u_add_v", u, v, out)


def sum(msg, out):  # noqa: A001
    return _Reduce("sum", msg, out)


def mean(msg, out):
    return _Reduce("mean", msg, out)


def max(msg, out):  # noqa: A001
    return _Reduce("max", msg, out)
