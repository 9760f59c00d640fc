"""Graph-convolution layers with the constructors of the reference's ``stag.zoo``
(stag/zoo/__init__.py:1-5), all running their neighbour aggregation on the fused CUDA kernels."""
from . import gat, gated_gcn, gcn, gin, graph_sage

GCN = gcn.GCN
GraphSAGE = graph_sage.GraphSAGE
GAT = gat.GAT
GIN = gin.GIN
GatedGCN = gated_gcn.GatedGCN

__all__ = ["GCN", "GraphSAGE", "GAT", "GIN", "GatedGCN"]
