"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the stochastic neighbour
aggregation (the reference's hot path), dependency-free (torch CPU ops only).

Each function cites the reference lines it follows.  Everything is written with
differentiable torch ops so that ``torch.autograd`` on this restatement is the
gradient oracle (the reference's backward is autograd through DGL's GSpMM, i.e.
the transposed aggregation + SDDMM).

Parity pin: tests/golden/*.npz were produced by the reference's OWN code
(/root/reference/stag, unmodified, over oracle/dgl_shim) with
oracle/make_golden.py; tests/test_oracle_cpu.py checks this file against them and,
when /root/reference is present, against the live reference.  The DGL layer under
the reference is itself a restatement (DGL absent) -- "DGL boundary unpinned".
"""
import torch


def apply_relu(w, relu):
    """stag/layers.py:98-99."""
    return w.relu() if relu else w


def in_norm(dst, num_nodes, w):
    """stag/layers.py:8-36 (_in_norm): rescale edge noise so that for every node and
    channel the in-edge weights sum to the in-degree; nodes whose weights sum to 0
    keep scale 1.  Returns (w_scaled [E,K], node_scaling [N,K])."""
    cur = torch.zeros((num_nodes,) + w.shape[1:], dtype=w.dtype).index_add(0, dst, w)  # :12-18
    indeg = torch.bincount(dst, minlength=num_nodes).unsqueeze(-1)                     # :21
    scale = torch.where(cur != 0.0, indeg / cur, torch.ones_like(cur))                 # :24-28
    return w * scale[dst], scale                                                       # :33-35


def u_mul_e_sum(src, dst, num_nodes, x, w):
    """graph.update_all(fn.u_mul_e('h','_edge_weight','m'), fn.sum('m','h'))
    stag/zoo/gcn.py:63,95 -- message x[src]*w (w broadcast over trailing dims),
    summed into dst in edge-id order."""
    m = x[src]
    if w is not None:
        ww = w
        while ww.dim() < m.dim():
            ww = ww.unsqueeze(-1)
        m = m * ww
    return torch.zeros((num_nodes,) + m.shape[1:], dtype=m.dtype).index_add(0, dst, m)


def aggregate(src, dst, num_nodes, x, w=None, reduce="sum", src_scale=None, dst_scale=None):
    """The fused operator's contract (include/stag_b200.h: stag_spmm_fwd):

        out[v,c] = dst_scale[v] * sum_{e:(u->v)} w[e,c] * (src_scale[u] * x[u,c])

    ``reduce='mean'`` divides by clamp(in_degree,1) (dgl fn.mean,
    stag/zoo/graph_sage.py:72)."""
    xs = x if src_scale is None else x * src_scale.reshape((-1,) + (1,) * (x.dim() - 1))
    out = u_mul_e_sum(src, dst, num_nodes, xs, w)
    if reduce == "mean":
        deg = torch.bincount(dst, minlength=num_nodes).to(out.dtype).clamp(min=1)
        out = out / deg.reshape((-1,) + (1,) * (out.dim() - 1))
    if dst_scale is not None:
        out = out * dst_scale.reshape((-1,) + (1,) * (out.dim() - 1))
    return out


def gcn_norms(src, dst, num_nodes, norm, dtype=torch.float32):
    """Degree scalings of stag/zoo/gcn.py:67-75 (source side, out-degree) and
    :100-108 (destination side, in-degree).  Returns (src_scale|None, dst_scale|None)."""
    s = d = None
    if norm in ("left", "both"):
        degs = torch.bincount(src, minlength=num_nodes).to(dtype).clamp(min=1)
        s = torch.pow(degs, -0.5) if norm == "both" else 1.0 / degs
    if norm in ("right", "both"):
        degs = torch.bincount(dst, minlength=num_nodes).to(dtype).clamp(min=1)
        d = torch.pow(degs, -0.5) if norm == "both" else 1.0 / degs
    return s, d


def gcn_forward(src, dst, num_nodes, feat, edge_weight, weight, bias, norm="both", activation=None):
    """stag/zoo/gcn.py:58-116 with the (only live) aggregate-then-transform branch."""
    s, d = gcn_norms(src, dst, num_nodes, norm, feat.dtype)
    feat_src = feat if s is None else feat * s.reshape((-1,) + (1,) * (feat.dim() - 1))  # :67-75
    rst = u_mul_e_sum(src, dst, num_nodes, feat_src, edge_weight)                       # :94-96
    if weight is not None:
        rst = torch.matmul(rst, weight)                                                  # :97-98
    if d is not None:
        rst = rst * d.reshape((-1,) + (1,) * (rst.dim() - 1))                            # :100-108
    if bias is not None:
        rst = rst + bias                                                                 # :110-111
    if activation is not None:
        rst = activation(rst)                                                            # :113-114
    return rst


def sage_forward(src, dst, num_nodes, feat, edge_weight, fc_self_w, fc_neigh_w, bias,
                 aggregator_type="mean", activation=None):
    """stag/zoo/graph_sage.py:44-119, 'mean' (:70-75) and 'gcn' (:76-91) aggregators.
    Linear weights are torch.nn.Linear layout [out, in]."""
    if aggregator_type == "mean":
        h_neigh = aggregate(src, dst, num_nodes, feat, edge_weight, reduce="mean")
        h_neigh = h_neigh @ fc_neigh_w.t()
        rst = feat @ fc_self_w.t() + h_neigh                                             # :107
    elif aggregator_type == "gcn":
        neigh = u_mul_e_sum(src, dst, num_nodes, feat, edge_weight)
        degs = torch.bincount(dst, minlength=num_nodes).to(feat)
        h_neigh = (neigh + feat) / (degs.unsqueeze(-1) + 1)                              # :89
        rst = h_neigh @ fc_neigh_w.t()
    else:
        raise KeyError(aggregator_type)
    if bias is not None:
        rst = rst + bias
    if activation is not None:
        rst = activation(rst)
    return rst


def reparam_normal(loc, scale, eps):
    """torch/distributions/normal.py:82-85 (rsample): loc + eps*scale, all expanded
    to [E,K] (stag/layers.py:117-124)."""
    return loc + eps * scale


def reparam_uniform(low, high, u):
    """torch/distributions/uniform.py:85-88 (rsample): low + u*(high-low)."""
    return low + u * (high - low)


def bernoulli_from_uniform(probs, u):
    """torch/distributions/bernoulli.py:116-119 (sample): 1 where u < p.  The
    reference draws through torch.bernoulli; the threshold form is the same law."""
    return (u < probs).to(u.dtype)


def kl_normal(loc_q, scale_q, loc_p, scale_p):
    """torch/distributions/kl.py:468-471 followed by .mean() (stag/layers.py:136-139)."""
    var_ratio = (scale_q / scale_p).pow(2)
    t1 = ((loc_q - loc_p) / scale_p).pow(2)
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).mean()


def kl_fallback(logq, logp):
    """stag/layers.py:141-143: sum over channels, mean over edges, difference."""
    return logq.sum(dim=-1).mean() - logp.sum(dim=-1).mean()


def stag_layer_gcn(src, dst, num_nodes, feat, w, weight, bias, norm="both", activation=None,
                   relu=False, in_norm_flag=False):
    """StagLayer.forward (stag/layers.py:84-113) around zoo.GCN, noise ``w`` [E,K]
    supplied externally (the shared-noise parity seam)."""
    w = apply_relu(w, relu)
    if in_norm_flag:
        w, _ = in_norm(dst, num_nodes, w)
    return gcn_forward(src, dst, num_nodes, feat, w, weight, bias, norm, activation)


def readout(feat, batch_num_nodes, op="sum"):
    """SumNodes / MeanNodes (stag/layers.py:156-178 -> dgl.sum_nodes/mean_nodes)."""
    seg = torch.repeat_interleave(torch.arange(len(batch_num_nodes)), batch_num_nodes)
    out = torch.zeros((len(batch_num_nodes),) + feat.shape[1:], dtype=feat.dtype).index_add(0, seg, feat)
    if op == "mean":
        out = out / batch_num_nodes.to(feat).clamp(min=1).reshape((-1,) + (1,) * (feat.dim() - 1))
    return out
