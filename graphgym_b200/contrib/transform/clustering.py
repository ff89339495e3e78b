"""Node clustering coefficient on the GPU — the label of the reference's headline node task
(``node_clustering_coefficient``: nx.clustering, ref: graphgym/models/feature_augment.py:81-82, used by
config/idgcn_tf and the other ``*_tf`` node configs).

For a simple undirected graph  c_i = 2 T_i / (d_i (d_i - 1))  with T_i the triangles through i, and the closed
walks of length three at i count every triangle twice: c_i = diag(A^3)_i / (d_i (d_i - 1)), 0 when d_i < 2.
diag(A^3) comes exact (int64) from the ID-GNN Fast cycle kernel (``closed_walk_counts``, csrc/cycle.cu), so the
label falls out of the same sparse propagation instead of a Python loop over nodes.
"""
import torch

from graphgym_b200 import ops
from graphgym_b200.contrib.transform.identity import closed_walk_counts


def clustering_coefficient(edge_index, n, graph_ptr=None):
    """[n] float64, equal to ``list(nx.clustering(G).values())`` for a simple undirected graph given as a symmetric
    directed edge list (both directions of every edge, no self loops, no duplicates — the DeepSNAP convention)."""
    ops._need_cuda(edge_index)
    walks, overflow = closed_walk_counts(edge_index, n, 3, symmetric=True, graph_ptr=graph_ptr)
    if overflow:
        raise OverflowError('closed-walk counts exceeded int64')
    deg = ops.segment_degree(ops.layout_build(edge_index, n, ops.LOOPS_REMOVE, ops.BY_TARGET)).double()
    pairs = deg * (deg - 1.0)
    # label preprocessing, off the per-step path: one elementwise division
    return torch.where(pairs > 0, walks[:, 2].double() / pairs.clamp(min=1.0), torch.zeros_like(pairs))
