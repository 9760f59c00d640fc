"""TEST INFRASTRUCTURE ONLY -- a pure-PyTorch stand-in for the parts of DGL that
/root/reference/stag touches, so that the reference's own Python code can be
executed unmodified on CPU (DGL itself is third-party, un-vendored, unpinned and
not installable here: no network).  Nothing under ``stag_b200/`` may import this.

What is restated (DGL semantics; upstream source is NOT under /root/reference, so
this is "published behaviour", anchored on the reference's own call sites):

* ``DGLGraph`` surface used by the reference:
  ``local_var/local_scope``            stag/layers.py:9,86  stag/zoo/gcn.py:58
  ``ndata/edata/srcdata/dstdata``      stag/layers.py:11,18,30,34  stag/zoo/gcn.py:62,94,96
  ``number_of_nodes/edges/dst_nodes``  stag/layers.py:118  stag/zoo/gcn.py:61
  ``in_degrees/out_degrees``           stag/layers.py:21  stag/zoo/gcn.py:68,101
  ``update_all / apply_edges``         stag/layers.py:12-15,33  stag/zoo/gcn.py:95
  ``is_block, to, edges, batch_num_nodes``
* ``dgl.function``: copy_edge/copy_e, copy_src/copy_u, u_mul_e, u_add_v, sum, mean, max
* ``dgl.nn``: GraphConv, SAGEConv, GATConv, GINConv constructors/attributes, edge_softmax
* ``dgl.rand_graph/graph/batch/add_self_loop/remove_self_loop/add_reverse_edges/
  sum_nodes/mean_nodes``

Reduction order: ``sum`` is ``zeros.index_add_(0, dst, msg)`` which on CPU adds
messages sequentially in edge-id order -- the same order a stable COO->CSC
conversion (DGL's) visits the in-edges of a node.
"""
import contextlib
import torch

from . import function  # noqa: F401
from . import base, utils  # noqa: F401


class _Frame(dict):
    def update(self, other=(), **kw):  # keep dict semantics, explicit for clarity
        super().update(other, **kw)


class _EdgeBatch:
    def __init__(self, g, src, dst):
        self.src = {k: v[src] for k, v in g.ndata.items()}
        self.dst = {k: v[dst] for k, v in g.ndata.items()}
        self.data = dict(g.edata)


class DGLGraph:
    is_block = False

    def __init__(self, src, dst, num_nodes, batch_num_nodes=None, batch_num_edges=None):
        self._src = src.to(torch.int64)
        self._dst = dst.to(torch.int64)
        self._n = int(num_nodes)
        self.ndata = _Frame()
        self.edata = _Frame()
        self._bnn = batch_num_nodes
        self._bne = batch_num_edges

    # --- frames -----------------------------------------------------------
    @property
    def srcdata(self):
        return self.ndata

    @property
    def dstdata(self):
        return self.ndata

    def local_var(self):
        g = DGLGraph(self._src, self._dst, self._n, self._bnn, self._bne)
        g.ndata = _Frame(self.ndata)
        g.edata = _Frame(self.edata)
        return g

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = _Frame(self.ndata), _Frame(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    # --- structure --------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes
    number_of_dst_nodes = number_of_nodes
    num_dst_nodes = number_of_nodes
    number_of_src_nodes = number_of_nodes
    num_src_nodes = number_of_nodes

    def number_of_edges(self):
        return int(self._src.shape[0])

    num_edges = number_of_edges

    def edges(self):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self._src, minlength=self._n)

    def batch_num_nodes(self):
        if self._bnn is None:
            return torch.tensor([self._n], dtype=torch.int64)
        return self._bnn

    def batch_num_edges(self):
        if self._bne is None:
            return torch.tensor([self.number_of_edges()], dtype=torch.int64)
        return self._bne

    @property
    def device(self):
        return self._src.device

    def to(self, device):
        g = DGLGraph(self._src.to(device), self._dst.to(device), self._n,
                     self._bnn, self._bne)
        g.ndata = _Frame({k: v.to(device) for k, v in self.ndata.items()})
        g.edata = _Frame({k: v.to(device) for k, v in self.edata.items()})
        return g

    # --- message passing --------------------------------------------------
    def apply_edges(self, func):
        if isinstance(func, function._Message):
            self.edata[func.out] = func(self)
        else:
            out = func(_EdgeBatch(self, self._src, self._dst))
            for k, v in out.items():
                self.edata[k] = v

    def update_all(self, message_func, reduce_func):
        msg = message_func(self)
        self.ndata[reduce_func.out] = reduce_func(self, msg)


def graph(data, num_nodes=None, idtype=None, device=None):
    src, dst = data
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    if num_nodes is None:
        num_nodes = int(max(src.max().item(), dst.max().item())) + 1 if src.numel() else 0
    return DGLGraph(src, dst, num_nodes)


def rand_graph(num_nodes, num_edges, idtype=None, device=None):
    # dgl.rand_graph draws edge ids uniformly; multi-edges and self loops may occur.
    eids = torch.randint(0, num_nodes * num_nodes, (num_edges,))
    return DGLGraph(eids // num_nodes, eids % num_nodes, num_nodes)


def batch(graphs):
    off, srcs, dsts = 0, [], []
    for g in graphs:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        off += g._n
    out = DGLGraph(torch.cat(srcs), torch.cat(dsts), off,
                   torch.tensor([g._n for g in graphs], dtype=torch.int64),
                   torch.tensor([g.number_of_edges() for g in graphs], dtype=torch.int64))
    keys = set(graphs[0].ndata.keys())
    for k in keys:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in set(graphs[0].edata.keys()):
        out.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return out


def add_self_loop(g):
    loop = torch.arange(g._n, dtype=torch.int64)
    out = DGLGraph(torch.cat([g._src, loop]), torch.cat([g._dst, loop]), g._n)
    out.ndata = _Frame(g.ndata)
    return out


def remove_self_loop(g):
    keep = g._src != g._dst
    out = DGLGraph(g._src[keep], g._dst[keep], g._n)
    out.ndata = _Frame(g.ndata)
    return out


def add_reverse_edges(g):
    out = DGLGraph(torch.cat([g._src, g._dst]), torch.cat([g._dst, g._src]), g._n)
    out.ndata = _Frame(g.ndata)
    return out


def _segment(g, feat, op):
    bnn = g.batch_num_nodes()
    seg = torch.repeat_interleave(torch.arange(len(bnn)), bnn)
    out = torch.zeros((len(bnn),) + feat.shape[1:], dtype=feat.dtype, device=feat.device)
    out.index_add_(0, seg.to(feat.device), feat)
    if op == "mean":
        shape = (-1,) + (1,) * (feat.dim() - 1)
        out = out / bnn.to(feat).clamp(min=1).reshape(shape)
    return out


def sum_nodes(g, name):
    return _segment(g, g.ndata[name], "sum")


def mean_nodes(g, name):
    return _segment(g, g.ndata[name], "mean")


from . import nn  # noqa: E402,F401
