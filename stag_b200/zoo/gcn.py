"""``GCN`` -- same constructor, attributes and forward signature as the reference's
``stag.zoo.GCN`` (stag/zoo/gcn.py:8-116, a ``dgl.nn.GraphConv`` subclass).

Reference forward: src-norm (out_deg^-1/2, :67-75) -> ``update_all(u_mul_e, sum)`` (:63,94-96)
-> ``@ W`` (:97-98) -> dst-norm (in_deg^-1/2, :100-108) -> bias (:110-111) -> activation
(:113-114).  Here the two degree scalings and the noise live inside the aggregation kernel
(``stag_spmm_fwd``), so one launch replaces the reference's ~10.  The dst-norm is a per-row
scalar and commutes with ``@ W``; it is applied in the aggregation epilogue.  ``@ W`` + bias + relu
run on the tcgen05 tensor-core kernel (``stag_gemm_tcgen05``, 3xTF32).
"""
import torch
from torch import nn

from .. import ops
from ..graph import as_graph


class DGLError(Exception):
    pass


class GCN(nn.Module):
    accepts_noise_spec = True
    accepts_sample_batch = True   # [S,N,D] features with one sample-batched NoiseSpec (StagModel._can_batch)

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm not in ("none", "both", "right", "left"):
            raise DGLError('Invalid norm value. Must be either "none", "both", "right" or "left".'
                           ' But got "{}".'.format(norm))
        self._in_feats = in_feats
        self._out_feats = out_feats
        self._norm = norm
        self._allow_zero_in_degree = allow_zero_in_degree
        if weight:
            self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        else:
            self.register_parameter("weight", None)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        """Glorot-uniform weight, zero bias (dgl.nn.GraphConv.reset_parameters)."""
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, weight=None, edge_weight=None):
        """``feat`` [N,D_in] (or [S,N,D_in] for S batched Monte-Carlo samples);
        ``edge_weight``: None | tensor [E,D_in] / [E,1] | :class:`stag_b200.ops.NoiseSpec`."""
        g = as_graph(graph)
        st = g._s
        if edge_weight is not None and not isinstance(edge_weight, ops.NoiseSpec):
            assert edge_weight.shape[-2 if edge_weight.dim() == 3 else 0] == g.number_of_edges()
        src_scale = dst_scale = None
        if self._norm in ("left", "both"):
            src_scale = st.scale(False, "rsqrt" if self._norm == "both" else "inv")
        if self._norm in ("right", "both"):
            dst_scale = st.scale(True, "rsqrt" if self._norm == "both" else "inv")
        if weight is not None:
            if self.weight is not None:
                raise DGLError("External weight is provided while at the same time the"
                               " module has defined its own weight parameter. Please"
                               " create the module with flag weight=False.")
        else:
            weight = self.weight
        n_samples = ops.spec_samples(edge_weight, feat)
        rst = ops.stochastic_aggregate(g, feat, edge_weight, reduce="sum", src_scale=src_scale,
                                       dst_scale=dst_scale, n_samples=n_samples)
        if weight is not None:
            # agg @ W + bias (+ relu) in one tcgen05 kernel; other activations are applied after it
            relu = self._activation in (torch.relu, torch.nn.functional.relu) or isinstance(self._activation, nn.ReLU)
            rst = ops.dense_transform(rst, weight, bias=self.bias, relu=relu)
            if self._activation is not None and not relu:
                rst = self._activation(rst)
            return rst
        if self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst

    def extra_repr(self):
        return "in={_in_feats}, out={_out_feats}, normalization={_norm}".format(**self.__dict__)
