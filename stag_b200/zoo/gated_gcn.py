"""``GatedGCN`` -- the reference's ``stag.zoo.GatedGCN`` (stag/zoo/gated_gcn.py:6-61):
``A h + sum_e w * h_u`` (or ``sum_e B h_u`` without edge weights, :30-36) -> BatchNorm ->
ReLU -> residual -> dropout."""
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..graph import as_graph


class GatedGCN(nn.Module):
    accepts_noise_spec = True
    accepts_sample_batch = False  # BatchNorm statistics are per pass: StagModel runs the samples sequentially

    def __init__(self, input_dim, output_dim, dropout=0.0, batch_norm=True, residual=False):
        super().__init__()
        self.in_channels = input_dim
        self.out_channels = output_dim
        self.dropout = dropout
        self.batch_norm = batch_norm
        self.residual = residual
        if input_dim != output_dim:
            self.residual = False
        self.A = nn.Linear(input_dim, output_dim, bias=True)
        self.B = nn.Linear(input_dim, output_dim, bias=True)
        self.bn_node_h = nn.BatchNorm1d(output_dim)

    def forward(self, g=None, h=None, edge_weight=None, graph=None, feat=None):
        # the reference names the arguments (g, h) (stag/zoo/gated_gcn.py:26) while StagLayer calls
        # base_layer.forward(graph=, feat=, edge_weight=) (stag/layers.py:109-113): accept both
        g = graph if g is None else g
        h = feat if h is None else h
        if isinstance(edge_weight, ops.NoiseSpec) and edge_weight.batched:
            raise NotImplementedError("GatedGCN: BatchNorm makes sample batching per-sample; "
                                      "StagModel(batch_samples=False)")
        g = as_graph(g)
        h_in = h
        if edge_weight is not None:
            sum_h = ops.stochastic_aggregate(g, h, edge_weight, reduce="sum")
        else:
            sum_h = ops.stochastic_aggregate(g, self.B(h), None, reduce="sum")
        h = self.A(h) + sum_h
        if self.batch_norm:
            h = self.bn_node_h(h)
        h = F.relu(h)
        if self.residual:
            h = h_in + h
        h = F.dropout(h, self.dropout)
        return h

    def __repr__(self):
        return "{}(in_channels={}, out_channels={})".format(
            self.__class__.__name__, self.in_channels, self.out_channels)
