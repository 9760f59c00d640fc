"""``GAT`` -- constructor and forward of the reference's ``stag.zoo.GAT`` (stag/zoo/gat.py:7-149).
The noise ``[E,H]`` multiplies the pre-softmax logits (:117-119).  Attention: the logits
``edge_weight * leaky_relu(el[u] + er[v])`` are computed INSIDE the segmented softmax over the in-edges of each node
(``stag_attention_softmax``, forward and a deterministic backward, csrc/edge_softmax.cu: no ``[E,H]`` logits, no atomics).
The weighted aggregation ``update_all(u_mul_e('ft','a'), sum)`` (:125-126) runs on ``stag_spmm_fwd`` with the attention as
external weights: all heads in one launch with expanded ``[E, H*F]`` weights on small (launch-bound) graphs (<= 4 MB),
else ``ops.heads_aggregate`` -- one launch per head, every head reading / writing its F columns in place.
``accepts_noise_spec`` is False: the noise multiplies the LOGITS, not the messages, so ``StagLayer`` hands this
layer a tensor emitted from the library's Philox stream (``stag_noise_emit``, K = num_heads).
"""
import torch
from torch import nn

from .. import ops
from ..graph import as_graph


class GAT(nn.Module):
    accepts_noise_spec = False

    def __init__(self, in_feats, out_feats, last=False, num_heads=4, feat_drop=0.0, attn_drop=0.0,
                 negative_slope=0.2, residual=False, activation=None, allow_zero_in_degree=False, bias=True):
        super().__init__()
        self._num_heads = num_heads
        self._in_src_feats = self._in_dst_feats = in_feats
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        if bias:
            self.bias = nn.Parameter(torch.empty(num_heads * out_feats))
        else:
            self.register_buffer("bias", None)
        if residual:
            if in_feats != out_feats * num_heads:
                self.res_fc = nn.Linear(in_feats, num_heads * out_feats, bias=False)
            else:
                self.res_fc = nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        self.activation = activation
        self.last = last
        self.sample_dimension = num_heads
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc.weight, gain=gain)
        nn.init.xavier_uniform_(self.attn_l, gain=gain)
        nn.init.xavier_uniform_(self.attn_r, gain=gain)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_uniform_(self.res_fc.weight, gain=gain)

    def forward(self, graph, feat, get_attention=False, edge_weight=None):
        g = as_graph(graph)
        src, dst = g.edges()
        N, H, F = g.number_of_nodes(), self._num_heads, self._out_feats
        h = self.feat_drop(feat)
        ft = self.fc(h).view(N, H, F)
        el = (ft * self.attn_l).sum(dim=-1)
        er = (ft * self.attn_r).sum(dim=-1)
        # attention: softmax over the in-edges of  edge_weight * leaky_relu(el[u] + er[v])  in one fused pass per direction
        # (stag_attention_softmax: the [E,H] logits and their autograd graph never exist)
        a = self.attn_drop(ops.attention_softmax(g, el, er, edge_weight, self.leaky_relu.negative_slope))   # [E,H]
        if a.shape[0] * H * F * 4 <= (4 << 20):
            # small graphs are launch-bound: ONE fused launch over the H*F channels, the attention of head k expanded to the
            # weight of channels k*F .. (k+1)*F - 1 (external per-channel weights [E, H*F], at most 4 MB; autograd sums the
            # SDDMM gradient back over the F channels of a head)
            aw = a.unsqueeze(-1).expand(a.shape[0], H, F).reshape(a.shape[0], H * F)
            rst = ops.stochastic_aggregate(g, ft.reshape(N, H * F), aw).view(N, H, F)
        else:
            # one launch per head on the per-edge-weight kernel, every head reading / writing its F columns in place (row
            # stride H*F): no [E, H*F] expansion, no per-head copies
            rst = ops.heads_aggregate(g, ft, a)
        if self.res_fc is not None:
            rst = rst + self.res_fc(h).view(N, -1, F)
        if self.bias is not None:
            rst = rst + self.bias.view(1, H, F)
        rst = rst.mean(-2) if self.last else rst.flatten(-2, -1)
        if self.activation:
            rst = self.activation(rst)
        if get_attention:
            return rst, a.unsqueeze(-1)
        return rst
