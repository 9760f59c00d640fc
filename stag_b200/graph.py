"""Graph object accepted by the stag_b200 layers.

It duck-types the part of ``dgl.DGLGraph`` the reference touches:
``local_var/local_scope``, ``ndata/edata/srcdata/dstdata``, ``number_of_nodes/edges``,
``in_degrees/out_degrees``, ``is_block``, ``update_all``, ``apply_edges``, ``to``,
``edges``, ``batch_num_nodes`` -- reference call sites stag/layers.py:9,11-14,21,33,86,
118,165,177; stag/zoo/gcn.py:58-68,94-101; stag/distributions.py:222-231 -- plus the
module-level helpers the reference scripts use (``graph``, ``rand_graph``, ``batch``,
``add_self_loop``, ``remove_self_loop``, ``add_reverse_edges``, ``sum_nodes``,
``mean_nodes``; scripts/arxiv_mle/gcn/run.py:53-55).  A real ``DGLGraph`` is accepted
through :func:`as_graph` when dgl is importable.

The compressed adjacencies (CSC by destination, CSR by source) are built ON DEVICE by
``stag_csx_build`` the first time an aggregation needs them and cached on the
structure, which ``local_var`` copies share.
"""
import contextlib
import ctypes

import torch

from . import _lib


class _Structure:
    """COO edge list + lazily built device CSC/CSR, shared by all views of a graph."""

    def __init__(self, src, dst, num_nodes, batch_num_nodes=None, batch_num_edges=None, eid_map=None, num_src=None):
        # eid_map [E]: id under which each local edge draws its Philox noise / indexes an external
        # noise tensor (row-partitioned graphs keep the edge ids of the unpartitioned graph)
        # num_src: rows of the gathered operand when they differ from the destination rows (the local
        # piece of a row-partitioned graph: owned destinations x owned + halo sources); default = num_nodes
        self.eid_map = eid_map
        self.src = src
        self.dst = dst
        self.num_nodes = int(num_nodes)
        self.num_src = self.num_nodes if num_src is None else int(num_src)
        self.num_edges = int(src.shape[0])
        self.batch_num_nodes = batch_num_nodes
        self.batch_num_edges = batch_num_edges
        self._csx = {}
        self._deg = {}
        self._scale = {}
        self._node_ptr = None
        self._ws = None

    @property
    def device(self):
        return self.src.device

    def csx(self, by_dst):
        """(StagGraph ctypes struct, keep-alive tensors) for the CSC (by_dst) or CSR."""
        key = bool(by_dst)
        if key not in self._csx:
            g, keep = build_csx(self.src, self.dst, self.num_nodes if key else self.num_src, key,
                                num_cols=self.num_src if key else self.num_nodes)
            if self.eid_map is not None and self.num_edges:
                m = self.eid_map.to(device=self.device, dtype=torch.int32)
                keep["eid"].copy_(m[keep["eid"].long()])
                last = keep["eidf"] < 0
                keep["eidf"].copy_(torch.where(last, keep["eid"] | -2147483648, keep["eid"]))
            self._csx[key] = (g, keep)
        return self._csx[key]

    def degrees(self, in_deg):
        key = bool(in_deg)
        if key not in self._deg:
            if self.device.type == "cuda":
                indptr = self.csx(key)[1]["indptr"]
                self._deg[key] = (indptr[1:] - indptr[:-1]).to(torch.int64)
            else:
                idx = self.dst if in_deg else self.src
                self._deg[key] = torch.bincount(idx, minlength=self.num_nodes if in_deg else self.num_src)
        return self._deg[key]

    def scale(self, in_deg, kind):
        """Cached degree scalings: kind 'rsqrt' = clamp(deg,1)^-1/2 (GCN norm='both',
        stag/zoo/gcn.py:68-70,101-103), 'inv' = 1/clamp(deg,1) ('left'/'right' and
        fn.mean), 'inv1' = 1/(deg+1) (SAGE gcn aggregator, stag/zoo/graph_sage.py:89)."""
        key = (bool(in_deg), kind)
        if key not in self._scale:
            d = self.degrees(in_deg).to(torch.float32)
            if kind == "rsqrt":
                v = torch.pow(d.clamp(min=1), -0.5)
            elif kind == "inv":
                v = 1.0 / d.clamp(min=1)
            elif kind == "inv1":
                v = 1.0 / (d + 1)
            else:
                raise KeyError(kind)
            self._scale[key] = v.contiguous()
        return self._scale[key]

    def node_ptr(self):
        if self._node_ptr is None:
            bnn = self.batch_num_nodes
            if bnn is None:
                bnn = torch.tensor([self.num_nodes], dtype=torch.int64)
            ptr = torch.zeros(len(bnn) + 1, dtype=torch.int64)
            ptr[1:] = torch.cumsum(bnn.cpu(), 0)
            self._node_ptr = ptr.to(torch.int32).to(self.device)
        return self._node_ptr

    def workspace(self, nbytes):
        """Grow-only scratch buffer handed to the library (owned by torch's allocator)."""
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        return self._ws


def build_csx(src, dst, num_nodes, by_dst, num_cols=None):
    """Run ``stag_csx_build`` on the current stream.  Returns (StagGraph, tensors).  `num_nodes` = rows of the
    compressed structure (destinations for the CSC, sources for the CSR), `num_cols` = rows of the gathered operand."""
    if src.device.type != "cuda":
        raise _lib.StagLibraryError(
            "stag_b200: graph structure must live on a CUDA device (got %s); there is no CPU path"
            % src.device)
    lib = _lib.load()
    E, N = int(src.shape[0]), int(num_nodes)
    dev = src.device
    src = src.contiguous()
    dst = dst.contiguous()
    with torch.cuda.device(dev):
        thr = lib.stag_hub_threshold()
        indptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        indices = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        eid = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        hub_rows = torch.empty(E // thr + 1, dtype=torch.int32, device=dev)
        hub_seg_ptr = torch.empty(E // thr + 2, dtype=torch.int32, device=dev)
        row_order = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
        items = torch.empty((lib.stag_csx_items_capacity(E, N), 4), dtype=torch.int32, device=dev)
        erow = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        eidf = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        ws_bytes = lib.stag_csx_workspace_bytes(E, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        counts = (ctypes.c_int32 * 3)()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.stag_csx_build(
            src.data_ptr(), dst.data_ptr(), E, N, 1 if by_dst else 0,
            indptr.data_ptr(), indices.data_ptr(), eid.data_ptr(),
            hub_rows.data_ptr(), hub_seg_ptr.data_ptr(), row_order.data_ptr(), items.data_ptr(), erow.data_ptr(), eidf.data_ptr(), counts,
            ws.data_ptr(), ws_bytes, stream))
    g = _lib.StagGraph()
    g.num_rows, g.num_cols, g.num_edges = N, (N if num_cols is None else int(num_cols)), E
    g.indptr, g.indices, g.eid = indptr.data_ptr(), indices.data_ptr(), eid.data_ptr()
    g.num_hubs, g.num_hub_segs = int(counts[0]), int(counts[1])
    g.hub_rows, g.hub_seg_ptr = hub_rows.data_ptr(), hub_seg_ptr.data_ptr()
    g.row_order = row_order.data_ptr() if N > 0 else None
    g.items, g.num_items = (items.data_ptr(), int(counts[2])) if N > 0 else (None, 0)
    g.erow, g.eidf = erow.data_ptr(), eidf.data_ptr()
    keep = {"indptr": indptr, "indices": indices[:E], "eid": eid[:E], "row_order": row_order[:N], "items": items[:int(counts[2])], "erow": erow[:E], "eidf": eidf[:E],
            "hub_rows": hub_rows[:g.num_hubs], "hub_seg_ptr": hub_seg_ptr[:g.num_hubs + 1]}
    return g, keep


class _EdgeBatch:
    def __init__(self, g):
        src, dst = g.edges()
        self.src = {k: v[src] for k, v in g.ndata.items()}
        self.dst = {k: v[dst] for k, v in g.ndata.items()}
        self.data = dict(g.edata)


class Graph:
    """Homogeneous directed multigraph (COO, int64 ids) with DGL-style data frames."""

    is_block = False

    def __init__(self, src=None, dst=None, num_nodes=None, _structure=None, eid_map=None, num_src=None):
        if _structure is None:
            src = torch.as_tensor(src, dtype=torch.int64)
            dst = torch.as_tensor(dst, dtype=torch.int64).to(src.device)
            if num_nodes is None:
                num_nodes = int(max(src.max().item(), dst.max().item())) + 1 if src.numel() else 0
            _structure = _Structure(src, dst, num_nodes, eid_map=eid_map, num_src=num_src)
        self._s = _structure
        self.ndata = {}
        self.edata = {}

    # frames ---------------------------------------------------------------------------
    @property
    def srcdata(self):
        return self.ndata

    @property
    def dstdata(self):
        return self.ndata

    def local_var(self):
        g = Graph(_structure=self._s)
        g.ndata = dict(self.ndata)
        g.edata = dict(self.edata)
        return g

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = dict(self.ndata), dict(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    # structure ------------------------------------------------------------------------
    @property
    def device(self):
        return self._s.device

    def number_of_nodes(self):
        return self._s.num_nodes

    def number_of_src_nodes(self):
        return self._s.num_src

    num_nodes = number_of_nodes
    number_of_dst_nodes = number_of_nodes
    num_src_nodes = number_of_src_nodes
    num_dst_nodes = number_of_nodes

    def number_of_edges(self):
        return self._s.num_edges

    num_edges = number_of_edges

    def edges(self):
        return self._s.src, self._s.dst

    def in_degrees(self):
        return self._s.degrees(True)

    def out_degrees(self):
        return self._s.degrees(False)

    def batch_num_nodes(self):
        if self._s.batch_num_nodes is None:
            return torch.tensor([self._s.num_nodes], dtype=torch.int64, device=self.device)
        return self._s.batch_num_nodes.to(self.device)

    def batch_num_edges(self):
        if self._s.batch_num_edges is None:
            return torch.tensor([self._s.num_edges], dtype=torch.int64, device=self.device)
        return self._s.batch_num_edges.to(self.device)

    @property
    def batch_size(self):
        return 1 if self._s.batch_num_nodes is None else len(self._s.batch_num_nodes)

    def to(self, device):
        device = torch.device(device)
        if device == self.device:
            s = self._s
        else:
            s = _Structure(self._s.src.to(device), self._s.dst.to(device), self._s.num_nodes,
                           self._s.batch_num_nodes, self._s.batch_num_edges, self._s.eid_map, self._s.num_src)
        g = Graph(_structure=s)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device) for k, v in self.edata.items()}
        return g

    def adj_tensors(self, fmt):
        """('csc'|'csr') -> (indptr, indices, eid) like DGL; int32 on device."""
        _, t = self._s.csx(fmt == "csc")
        return t["indptr"], t["indices"], t["eid"]

    # message passing ------------------------------------------------------------------
    def apply_edges(self, func):
        from . import function as fn
        if isinstance(func, fn.Message):
            self.edata[func.out] = func.edgewise(self)
        else:
            out = func(_EdgeBatch(self))
            for k, v in out.items():
                self.edata[k] = v

    def update_all(self, message_func, reduce_func):
        """DGL-style message passing with the builtin functions of ``stag_b200.function``
        (u_mul_e / copy_u / copy_e with sum / mean), executed by the fused CUDA kernels."""
        from . import function as fn
        if not isinstance(message_func, fn.Message) or not isinstance(reduce_func, fn.Reduce):
            raise NotImplementedError("update_all supports stag_b200.function builtins only")
        self.ndata[reduce_func.out] = fn.run(self, message_func, reduce_func)

    def __repr__(self):
        return "Graph(num_nodes=%d, num_edges=%d, device=%s)" % (
            self.number_of_nodes(), self.number_of_edges(), self.device)


def as_graph(g):
    """Accept a stag_b200 Graph or (when dgl is importable) a DGLGraph."""
    if isinstance(g, Graph):
        return g
    if hasattr(g, "edges") and hasattr(g, "number_of_nodes"):
        cached = getattr(g, "_stag_b200_graph", None)
        if cached is not None:
            return cached
        src, dst = g.edges()
        out = Graph(src.to(torch.int64), dst.to(torch.int64), g.number_of_nodes())
        try:
            bnn = g.batch_num_nodes()
            if len(bnn) > 1:
                out._s.batch_num_nodes = bnn.to(torch.int64)
            g._stag_b200_graph = out
        except Exception:
            pass
        return out
    raise TypeError("expected a stag_b200.Graph or a DGLGraph, got %r" % type(g))


# dgl-like constructors / transforms ---------------------------------------------------
def graph(data, num_nodes=None, idtype=None, device=None):
    src, dst = data
    g = Graph(src, dst, num_nodes)
    return g.to(device) if device is not None else g


def rand_graph(num_nodes, num_edges, idtype=None, device=None):
    eids = torch.randint(0, num_nodes * num_nodes, (num_edges,))
    g = Graph(eids // num_nodes, eids % num_nodes, num_nodes)
    return g.to(device) if device is not None else g


def batch(graphs):
    """Block-diagonal batching (dgl.batch); records batch_num_nodes / batch_num_edges.  Runs on the device the
    graphs live on: one concatenation of the edge lists and one repeat_interleave of the node offsets, no per-graph
    arithmetic (a molhiv minibatch is 32-128 graphs of ~25 nodes: scripts/molhiv_mle/run.py builds one per step)."""
    graphs = [as_graph(g) for g in graphs]
    dev = graphs[0].device if graphs else torch.device("cpu")
    nn_ = torch.tensor([g.number_of_nodes() for g in graphs], dtype=torch.int64)
    ne_ = torch.tensor([g.number_of_edges() for g in graphs], dtype=torch.int64)
    off = torch.cumsum(nn_, 0) - nn_                                  # first node id of every graph
    eoff = torch.repeat_interleave(off.to(dev), ne_.to(dev))          # [E] offset of every edge's graph
    srcs = torch.cat([g._s.src for g in graphs]) + eoff if graphs else torch.empty(0, dtype=torch.int64)
    dsts = torch.cat([g._s.dst for g in graphs]) + eoff if graphs else torch.empty(0, dtype=torch.int64)
    st = _Structure(srcs, dsts, int(nn_.sum()), nn_, ne_)
    out = Graph(_structure=st)
    for k in graphs[0].ndata.keys():
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata.keys():
        out.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return out


def _derived(g, src, dst):
    out = Graph(src, dst, g.number_of_nodes())
    out.ndata = dict(g.ndata)
    return out


def add_self_loop(g):
    g = as_graph(g)
    s, d = g.edges()
    loop = torch.arange(g.number_of_nodes(), dtype=torch.int64, device=s.device)
    return _derived(g, torch.cat([s, loop]), torch.cat([d, loop]))


def remove_self_loop(g):
    g = as_graph(g)
    s, d = g.edges()
    keep = s != d
    return _derived(g, s[keep], d[keep])


def add_reverse_edges(g):
    g = as_graph(g)
    s, d = g.edges()
    return _derived(g, torch.cat([s, d]), torch.cat([d, s]))


def sum_nodes(g, name):
    from . import ops
    return ops.segment_reduce(as_graph(g), g.ndata[name], mean=False)


def mean_nodes(g, name):
    from . import ops
    return ops.segment_reduce(as_graph(g), g.ndata[name], mean=True)
