"""Multi-GPU plumbing for the three axes along which the stochastic aggregation shards
(SURVEY.md 8(e)); one process per GPU, ``torch.distributed`` (NCCL over NVLink on the box,
gloo in the CPU tests).  The reference has no distributed code at all (single ``cuda:0``,
scripts/*/run.py ``model.cuda()``), so nothing here replaces a reference file.

* MC samples        -- graph and parameters replicated; rank r draws the Philox sample indices
                       ``sample_base + [0, S_local)``.  Inference: all-reduce(sum) of the [N,C]
                       probabilities.  Training: one flat-bucket all-reduce of the gradients.
* graph minibatches -- disjoint block-diagonal batches per rank (no cross edges), DDP-style
                       gradient averaging with the same flat bucket.
* rows of one graph -- 1-D block row partition of the destinations: rank r owns the in-edges
                       (CSC rows) of nodes [lo_r, hi_r) and the matching X row block.  Forward:
                       halo exchange (only the source rows a rank's edges reference, one
                       variable-size all-to-all), local aggregation into the owned rows.
                       Backward: the partial gradients of the halo rows travel back to their
                       owners along the same lists (`RowPartition`).
"""
import torch
import torch.distributed as dist


def shard_samples(n_samples, rank, world):
    """(sample_base, n_local): contiguous split of the global MC sample indices."""
    base, rem = divmod(int(n_samples), int(world))
    n_local = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, n_local


def shard_items(n_items, rank, world):
    """Indices of the minibatches / graphs owned by `rank` (round-robin)."""
    return list(range(rank, int(n_items), int(world)))


def allreduce_gradients(parameters, group=None, average=True):
    """One flat bucket, one all-reduce (the gradient sets here are KB-MB: latency-bound)."""
    params = [p for p in parameters if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()


class GradBucket:
    """The gradients of `parameters` as views of ONE flat buffer, so that the all-reduce of a step is a single
    collective on memory autograd has already written: no concatenation before it and no copies after it (the
    per-parameter copies of ``allreduce_gradients`` cost 0.38 ms of a 2 ms PPI minibatch step, measured)."""

    def __init__(self, parameters, group=None):
        self.params = [p for p in parameters if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else None
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad = self.flat[off:off + k].view_as(p)   # backward accumulates in place into the bucket
            off += k

    def zero(self):
        self.flat.zero_()

    def allreduce(self, average=True):
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat /= dist.get_world_size(self.group)
        return self.flat.numel()


def mc_mean(local_sum, n_samples_total, group=None):
    """Monte-Carlo predictive mean from per-rank partial sums of the [N,C] outputs
    (stag/models.py:46-55 computes the mean of the stacked per-sample outputs)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM, group=group)
    return local_sum / float(n_samples_total)


def row_blocks(num_nodes, world):
    """Block boundaries [world+1]; every block has ceil(N/world) rows except the last ones
    (equal-sized blocks are what all_gather / reduce_scatter need; the tail is padded)."""
    per = (int(num_nodes) + world - 1) // world
    return [min(r * per, num_nodes) for r in range(world + 1)], per


def edge_balanced_blocks(dst, num_nodes, world):
    """Block boundaries [world+1] such that every block of destination rows holds about E / world in-edges (the
    work of the fused pass is per edge; on a power-law graph equal ROW blocks leave the ranks up to 10 % apart, and
    every exchange then waits for the slowest).  Deterministic: every rank computes the same cuts from `dst`."""
    deg = torch.bincount(dst, minlength=int(num_nodes))
    cum = torch.cumsum(deg, 0)
    E = int(dst.numel())
    targets = torch.tensor([E * r // world for r in range(1, world)], dtype=cum.dtype, device=cum.device)
    cuts = (torch.searchsorted(cum, targets, right=False) + 1).clamp_(max=int(num_nodes)).tolist() if world > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [int(num_nodes)]
    for i in range(1, len(bounds)):                      # monotone even on degenerate inputs
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds, max(bounds[i + 1] - bounds[i] for i in range(world))


def _all_to_all_rows(out, inp, recv_counts, send_counts, group=None, async_op=False):
    """Variable-size row exchange: rows [sum(send_counts[:q]), +send_counts[q]) of `inp` go to rank q, the rows
    received from rank q land at [sum(recv_counts[:q]), +recv_counts[q]) of `out`.  NCCL: one
    all_to_all_single; gloo (CPU tests) has no all-to-all: point-to-point pairs."""
    if dist.get_backend(group) != "gloo":
        return dist.all_to_all_single(out, inp, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                                      group=group, async_op=async_op)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    so = ro = 0
    reqs = []
    for q in range(world):
        sv, rv = inp[so:so + send_counts[q]], out[ro:ro + recv_counts[q]]
        if q == rank:
            rv.copy_(sv)
        else:
            if send_counts[q]:
                reqs.append(dist.isend(sv.contiguous(), q, group=group))
            if recv_counts[q]:
                reqs.append(("recv", rv, q))
        so += send_counts[q]
        ro += recv_counts[q]
    for r in reqs:
        if isinstance(r, tuple):
            buf = torch.empty_like(r[1])
            dist.recv(buf, r[2], group=group)
            r[1].copy_(buf)
    for r in reqs:
        if not isinstance(r, tuple):
            r.wait()
    return None


class RowPartition:
    """1-D row partition of one large graph over the ranks of `group`.

    Rank r owns the destination rows [lo_r, hi_r) -- their in-edges and the matching X row block.  Two forms of the
    feature exchange:

    * ``halo=True`` (default): every rank receives ONLY the source rows its own edges reference.  The needed-row
      lists are exchanged once (``setup_halo``); per layer ``exchange`` gathers the requested rows of the owned
      block into a send buffer and moves them with one variable-size all-to-all (NCCL ``all_to_all_single``).
      The local graph is bipartite: destinations = owned rows, sources = [owned rows | halo rows sorted by global
      id] (``stag_b200.Graph(..., num_src=)``), so nothing is computed or stored for rows a rank does not own.
      Backward: the partial gradients of the halo rows travel back along the same lists (``exchange_back``) and
      are added into the owners' blocks -- the reduce-scatter of dX restricted to the rows that were used.
    * ``halo=False``: all-gather of whole X row blocks / reduce-scatter of a full [N,D] partial dX (round 1).

    ``balance="edges"`` (halo form) cuts the rows so that every rank holds about E / world in-edges instead of N / world
    rows: the fused pass costs per edge, and every exchange waits for the slowest rank.

    Either way the local graph keeps the edge ids of the unpartitioned graph (``eid_map``): the Philox noise of every
    edge, and the order in which a row's in-edges are summed, are those of the single-GPU run -- the forward is
    bitwise identical.  ``aggregate(graph, feat, edge_weight, ...)`` is ``stag_b200.ops.stochastic_aggregate`` on
    the GPU; the CPU tests inject the oracle.
    """

    def __init__(self, src, dst, num_nodes, rank, world, group=None, halo=True, balance="rows"):
        self.rank, self.world, self.group = rank, world, group
        self.num_nodes = int(num_nodes)
        if balance == "edges":
            if not halo:
                raise ValueError("balance='edges' needs the halo form (all_gather / reduce_scatter want equal row blocks)")
            self.bounds, self.per = edge_balanced_blocks(dst, num_nodes, world)
        elif balance == "rows":
            self.bounds, self.per = row_blocks(num_nodes, world)
        else:
            raise ValueError("balance must be 'rows' or 'edges'")
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_own = self.hi - self.lo
        own = (dst >= self.lo) & (dst < self.hi)
        self.edge_ids = torch.nonzero(own, as_tuple=False).reshape(-1)   # global edge ids, increasing
        self.src = src[own]
        self.dst = dst[own]
        self.padded = self.per * world
        self.halo = bool(halo)
        if self.halo:
            self._plan_halo()

    def _owner(self, rows):
        """Owning rank of global rows (int64 tensor)."""
        cuts = torch.tensor(self.bounds[1:-1], dtype=rows.dtype, device=rows.device)
        return torch.bucketize(rows, cuts, right=True)

    # ---- halo plan ------------------------------------------------------------------------------------------
    def _plan_halo(self):
        src = self.src
        remote = (src < self.lo) | (src >= self.hi)
        self.need = torch.unique(src[remote])            # sorted global ids: grouped by owner, increasing
        owner = self._owner(self.need)
        self.recv_counts = torch.bincount(owner, minlength=self.world).tolist()
        self.n_halo = int(self.need.numel())
        self.n_ext = self.n_own + self.n_halo
        # extended source index: owned rows first, then the halo rows in `need` order
        ext = torch.where(remote, self.n_own + torch.searchsorted(self.need, src), src - self.lo)
        self.src_ext = ext
        self.dst_loc = self.dst - self.lo
        self.send_counts = None
        self.send_idx = None

    def setup_halo(self):
        """Tell every owner which of its rows this rank needs (once per graph).  Collective."""
        dev = self.need.device
        if self.world == 1:
            self.send_counts, self.send_idx = [0], torch.empty(0, dtype=torch.int64, device=dev)
            return self
        rc = torch.tensor(self.recv_counts, dtype=torch.int64, device=dev)
        sc = torch.empty_like(rc)
        if dist.get_backend(self.group) == "gloo":
            got = [None] * self.world
            dist.all_gather_object(got, self.recv_counts, group=self.group)
            sc = torch.tensor([got[q][self.rank] for q in range(self.world)], dtype=torch.int64)
        else:
            dist.all_to_all_single(sc, rc, group=self.group)
        self.send_counts = sc.tolist()
        owner_lo = torch.tensor(self.bounds[:-1], dtype=torch.int64, device=dev)
        owner = self._owner(self.need)
        req = (self.need - owner_lo[owner]).contiguous()     # row indices inside the owner's block
        self.send_idx = torch.empty(int(sum(self.send_counts)), dtype=torch.int64, device=dev)
        _all_to_all_rows(self.send_idx, req, self.send_counts, self.recv_counts, group=self.group)
        return self

    def halo_bytes(self, D, itemsize=4):
        """(received, sent) bytes of one exchange of [*, D] rows."""
        return self.n_halo * D * itemsize, int(sum(self.send_counts or [0])) * D * itemsize

    def local_graph(self, graph_cls):
        """Local structure: bipartite (owned destinations x owned + halo sources) with ``halo=True``, else over
        all N nodes (only owned destinations have in-edges).  Edge ids are those of the unpartitioned graph."""
        if self.halo:
            return graph_cls(self.src_ext, self.dst_loc, self.n_own, eid_map=self.edge_ids, num_src=self.n_ext)
        return graph_cls(self.src, self.dst, self.num_nodes, eid_map=self.edge_ids)

    def exchange(self, x_block, out=None, async_op=False):
        """Owned rows [n_own, D] (or [S, n_own, D]) -> extended operand [n_ext, D] ([S, n_ext, D]): own rows, then
        the halo rows received from their owners.  With ``async_op`` returns (x_ext, work); the collective runs on
        the communicator's stream and the caller overlaps it with work that does not read the halo rows."""
        lead = x_block.shape[:-2]
        D = x_block.shape[-1]
        if out is None:
            out = torch.empty(lead + (self.n_ext, D), dtype=x_block.dtype, device=x_block.device)
        out[..., : self.n_own, :].copy_(x_block)
        if self.world == 1 or (self.n_halo == 0 and sum(self.send_counts) == 0):
            return (out, None) if async_op else out
        if len(lead) == 0:
            send = x_block.index_select(0, self.send_idx)
            recv = out[self.n_own:]
            work = _all_to_all_rows(recv, send, self.recv_counts, self.send_counts, group=self.group, async_op=async_op)
        else:   # [S, rows, D]: rows outermost on the wire so that one all-to-all moves every sample
            S = lead[0]
            send = x_block.index_select(1, self.send_idx).transpose(0, 1).contiguous().reshape(-1, S * D)
            recv = torch.empty((self.n_halo, S * D), dtype=x_block.dtype, device=x_block.device)
            work = _all_to_all_rows(recv, send, self.recv_counts, self.send_counts, group=self.group, async_op=False)
            out[:, self.n_own:, :].copy_(recv.reshape(self.n_halo, S, D).transpose(0, 1))
        return (out, work) if async_op else out

    def exchange_back(self, dx_ext):
        """Partial gradient of the extended operand [n_ext, D] -> gradient of the owned rows [n_own, D]: the halo
        part travels back to the owners and is added there (index_add in peer order: deterministic)."""
        D = dx_ext.shape[-1]
        dx = dx_ext[: self.n_own].clone()
        if self.world == 1:
            return dx
        back = torch.empty((int(sum(self.send_counts)), D), dtype=dx_ext.dtype, device=dx_ext.device)
        _all_to_all_rows(back, dx_ext[self.n_own:].contiguous(), self.send_counts, self.recv_counts, group=self.group)
        # a peer asks for every row at most once, so its segment is a plain gather-add-scatter (no atomics: an
        # index_add_ over the 1.2 M x 100 halo of the products shape took 9 of the 10.7 ms of this call); segments
        # in peer order: deterministic
        off = 0
        for c in self.send_counts:
            if c:
                idx = self.send_idx[off:off + c]
                dx.index_copy_(0, idx, dx.index_select(0, idx).add_(back[off:off + c]))
            off += c
        return dx

    # ---- whole-block form (halo=False) --------------------------------------------------------------------------
    def gather_features(self, x_block):
        """all-gather of the owned X row blocks -> full X [N,D] (the halo exchange)."""
        D = x_block.shape[-1]
        pad = torch.zeros((self.per, D), dtype=x_block.dtype, device=x_block.device)
        pad[: x_block.shape[0]] = x_block
        if self.world == 1:
            return pad[: self.num_nodes]
        full = torch.empty((self.padded, D), dtype=x_block.dtype, device=x_block.device)
        dist.all_gather_into_tensor(full, pad, group=self.group)
        return full[: self.num_nodes]

    def scatter_gradients(self, dx_full):
        """reduce-scatter(sum) of the per-rank partial dX [N,D] -> owned block [hi-lo, D]."""
        D = dx_full.shape[-1]
        if self.world == 1:
            return dx_full[self.lo:self.hi]
        pad = torch.zeros((self.padded, D), dtype=dx_full.dtype, device=dx_full.device)
        pad[: self.num_nodes] = dx_full
        out = torch.empty((self.per, D), dtype=dx_full.dtype, device=dx_full.device)
        if dist.get_backend(self.group) == "gloo":   # gloo has no reduce_scatter_tensor
            dist.all_reduce(pad, op=dist.ReduceOp.SUM, group=self.group)
            out = pad[self.rank * self.per:(self.rank + 1) * self.per]
        else:
            dist.reduce_scatter_tensor(out, pad, op=dist.ReduceOp.SUM, group=self.group)
        return out[: self.hi - self.lo]
