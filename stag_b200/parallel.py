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
                       all-gather of the X row blocks (halo), local aggregation into the owned
                       rows.  Backward: local transposed aggregation gives a partial dX for ALL
                       nodes, reduce-scatter(sum) returns each rank its block.
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


class RowPartition:
    """1-D row partition of one large graph over the ranks of `group`.

    ``aggregate(graph, feat, edge_weight, src_scale, dst_scale, n_samples)`` is the local
    operator (``stag_b200.ops.stochastic_aggregate`` on the GPU; the CPU tests inject the
    oracle).  The local graph keeps GLOBAL source ids (it gathers from the all-gathered X)
    and global destination ids restricted to the owned block, so edge ids -- and therefore the
    Philox noise of every edge -- are those of the unpartitioned graph.
    """

    def __init__(self, src, dst, num_nodes, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.num_nodes = int(num_nodes)
        self.bounds, self.per = row_blocks(num_nodes, world)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        own = (dst >= self.lo) & (dst < self.hi)
        self.edge_ids = torch.nonzero(own, as_tuple=False).reshape(-1)   # global edge ids, increasing
        self.src = src[own]
        self.dst = dst[own]
        self.padded = self.per * world

    def local_graph(self, graph_cls):
        """Local structure over all N nodes (only owned destinations have in-edges)."""
        return graph_cls(self.src, self.dst, self.num_nodes, eid_map=self.edge_ids)

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
