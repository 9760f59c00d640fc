"""Graph-convolution layers with the reference's constructors (stag/zoo/__init__.py:1-5)."""
from .gcn import GCN
from .graph_sage import GraphSAGE
from .gat import GAT
from .gin import GIN
from .gated_gcn import GatedGCN

__all__ = ["GCN", "GraphSAGE", "GAT", "GIN", "GatedGCN"]
