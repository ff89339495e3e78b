"""CPU restatement of the TensorFlow ID-GNN layers of TfgIDLayer.py (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

The reference's second front door (main_zd.py) drives Keras layers written on tf_geometric.  TensorFlow / tf_geometric are
not installable here, so these functions restate the layers' arithmetic in torch from the reference's own source:
  gcn_id   TfgIDLayer.py:478-525 with gcn_norm_adj :528-566 (SparseAdj.add_self_loop = append N loops, sparse_adj.py:58-63)
  IDSAGE   TfgIDLayer.py:74-120   (half the units each for [self | neighbour mean], in-layer activation)
  IDGIN    TfgIDLayer.py:143-167
  gat_id   TfgIDLayer.py:269-388  (scaled dot-product scorer, softmax over the row's edges, sparse_adj.py:136-151)
tf_geometric helper semantics restated from its documentation ("parity unpinned" for those: `segment_softmax` =
exp(z - max) / sum, `add_self_loop_edge` = append (i, i) for every i with `fill_weight`, `mean_reducer` = segment mean over
the rows that occur).  Convention: Tfg aggregates at row = edge_index[0] and gathers col = edge_index[1].
"""
import torch

from .layers import id_transform


def _add_self_loop(edge_index, n, w=None, fill=1.0):
    loop = torch.arange(n, dtype=torch.long).unsqueeze(0).repeat(2, 1)
    ei = torch.cat([edge_index, loop], 1)
    if w is None:
        w = torch.ones(edge_index.size(1), dtype=torch.float64)
    return ei, torch.cat([w, w.new_full((n,), fill)])


def _spmm(ei, w, h, n):
    """sparse_adj @ h: out[row] += w * h[col] (sparse_adj.py:91-97)."""
    return torch.zeros((n, h.size(1)), dtype=h.dtype).index_add_(0, ei[0], h[ei[1]] * w.view(-1, 1).to(h.dtype))


def gcn_id(x, edge_index, ids, kernel, kernel_id, bias=None, activation=torch.relu):
    n = x.size(0)
    ei, w = _add_self_loop(edge_index, n)                       # renorm=True: self loops BEFORE the degree
    w = w.to(x.dtype)
    deg = torch.zeros(n, dtype=x.dtype).index_add_(0, ei[0], w)
    dis = deg.pow(-0.5)
    dis = torch.where(torch.isinf(dis) | torch.isnan(dis), torch.zeros_like(dis), dis)
    norm = dis[ei[0]] * w * dis[ei[1]]
    h = _spmm(ei, norm, id_transform(x, ids, kernel, kernel_id), n)
    if bias is not None:
        h = h + bias
    return activation(h) if activation is not None else h


def id_sage(x, edge_index, ids, self_kernel, id_kernel, neighbor_kernel, bias=None, activation=torch.relu, concat=True):
    n = x.size(0)
    row, col = edge_index
    s = torch.zeros((n, x.size(1)), dtype=x.dtype).index_add_(0, row, x[col])
    cnt = torch.zeros(n, dtype=x.dtype).index_add_(0, row, torch.ones(row.numel(), dtype=x.dtype))
    mean = s / cnt.clamp(min=1).view(-1, 1)
    h = id_transform(x, ids, self_kernel, id_kernel)
    nb = mean @ neighbor_kernel
    h = torch.cat([h, nb], 1) if concat else h + nb
    if bias is not None:
        h = h + bias
    return activation(h) if activation is not None else h


def id_gin(x, edge_index, ids, mlp, mlp_id, eps=0.0):
    """mlp / mlp_id: callables (TfgIDLayer.py:143-167; no self-loop removal there, unlike idconv.py:370)."""
    n = x.size(0)
    ones = torch.ones(edge_index.size(1), dtype=x.dtype)
    h = x * (1.0 + eps) + _spmm(edge_index, ones, x, n)
    return mlp(h).index_add(0, ids, mlp_id(h.index_select(0, ids)))


def gat_id(x, edge_index, ids, query_kernel, query_bias, key_kernel, key_bias, kernel, kernel_id, bias=None,
           activation=torch.relu):
    """num_heads = 1, query / key activation relu (the IDGAT defaults, TfgIDLayer.py:170-182)."""
    n = x.size(0)
    ei, _ = _add_self_loop(edge_index, n)
    row, col = ei
    q = torch.relu(x @ query_kernel + query_bias)[row]
    k = torch.relu(x @ key_kernel + key_bias)[col]
    v = id_transform(x, ids, kernel, kernel_id)
    score = (q * k).sum(-1) / (q.size(-1) ** 0.5)
    mx = torch.full((n,), float('-inf'), dtype=x.dtype).scatter_reduce(0, row, score, 'amax', include_self=True)
    e = (score - mx[row]).exp()
    alpha = e / torch.zeros(n, dtype=x.dtype).index_add_(0, row, e)[row]
    h = _spmm(ei, alpha, v, n)
    if bias is not None:
        h = h + bias
    return activation(h) if activation is not None else h
