"""ID-GNN layers behind the reference's registry names ``idconv / gcnidconv / sageidconv /
gatidconv / ginidconv`` (ref: graphgym/contrib/layer/idconv.py:444-448).

Same constructor signature, parameter names (``weight``, ``weight_id``, ``bias``, ``att``), init
(glorot / zeros) and ``forward(batch) -> batch`` contract as the reference; the math underneath is
the gg_* CUDA path:

    heterogeneous transform  X W + onehot(id) X W_id   -> one multi-segment GEMM   (functional.id_linear)
    self-loop edit + D^-1/2 A D^-1/2                   -> cached CSR/CSC layout     (graph.GraphLayout)
    propagate (gather, scale, scatter)                 -> CSR SpMM, CSC SpMM bwd    (functional.aggregate)
"""
import math

import torch
import torch.nn as nn
from torch.nn import Parameter

from graphgym_b200 import functional as F_
from graphgym_b200 import ops
from graphgym_b200.config import cfg
from graphgym_b200.graph import get_id_index, get_layout
from graphgym_b200.register import register_layer


def glorot_(t):
    """PyG ``glorot``: U(-a, a), a = sqrt(6 / (size(-2) + size(-1)))."""
    if t is not None:
        a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        with torch.no_grad():
            t.uniform_(-a, a)


def zeros_(t):
    if t is not None:
        with torch.no_grad():
            t.zero_()


class _IDBase(nn.Module):
    """weight / weight_id / bias parameters shared by the four weight-matrix ID layers."""

    def _make_params(self, fan_in, fan_out, bias):
        self.weight = Parameter(torch.empty(fan_in, fan_out))
        self.weight_id = Parameter(torch.empty(fan_in, fan_out))
        if bias:
            self.bias = Parameter(torch.empty(fan_out))
        else:
            self.register_parameter('bias', None)

    def reset_parameters(self):
        glorot_(self.weight)
        glorot_(self.weight_id)
        zeros_(self.bias)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class _NormalisedIDConv(_IDBase):
    """Common body of GeneralIDConvLayer and GCNIDConvLayer: transform, optional GCN norm, propagate."""

    def _setup(self, in_channels, out_channels, improved, cached, bias, normalize, aggr):
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached, self.normalize, self.aggr = improved, cached, normalize, aggr
        self._make_params(in_channels, out_channels, bias)
        self.reset_parameters()
        self.cached_num_edges = None

    def forward(self, x, edge_index, id, edge_weight=None):
        if edge_weight is not None or self.improved:
            raise NotImplementedError('edge_weight / improved are never set by the GraphGym wrappers '
                                      '(ref: idconv.py:388-404); not on the accelerated path')
        if self.cached and self.cached_num_edges is not None and \
                edge_index.size(1) != self.cached_num_edges:
            raise RuntimeError(
                'Cached {} number of edges, but found {}. Please disable the caching behavior of '
                'this layer by removing the `cached=True` argument in its constructor.'.format(
                    self.cached_num_edges, edge_index.size(1)))
        self.cached_num_edges = edge_index.size(1)
        n = x.size(0)
        h = F_.id_linear(x, self.weight, self.weight_id, get_id_index(id, n))
        if self.normalize:
            # add_remaining_self_loops + deg over edge_index[0] (ref: idconv.py:139-148)
            layout = get_layout(edge_index, n, ops.LOOPS_ADD_REMAINING)
            kind = 'gcn_src_mean' if self.aggr == 'mean' else 'gcn_src'
        else:
            layout, kind = get_layout(edge_index, n, ops.LOOPS_KEEP), \
                ('mean' if self.aggr == 'mean' else 'sum')
        return F_.aggregate(h, layout, kind, 0.0, self.bias)


class GeneralIDConvLayer(_NormalisedIDConv):
    """ref: idconv.py:16-101 — aggr and normalisation come from ``cfg.gnn`` at construction."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, bias=True, **kwargs):
        super().__init__()
        if cfg.gnn.agg not in ('add', 'mean'):
            raise NotImplementedError("cfg.gnn.agg = {!r}: 'add' and 'mean' are on the accelerated "
                                      "path".format(cfg.gnn.agg))
        self._setup(in_channels, out_channels, improved, cached, bias, cfg.gnn.normalize_adj,
                    cfg.gnn.agg)


class GCNIDConvLayer(_NormalisedIDConv):
    """ref: idconv.py:104-189."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, bias=True,
                 normalize=True, **kwargs):
        super().__init__()
        self._setup(in_channels, out_channels, improved, cached, bias, normalize, 'add')


class SAGEIDConvLayer(_IDBase):
    """ref: idconv.py:192-263 — mean aggregation, then [x | mean] (W, W_id on centres)."""

    def __init__(self, in_channels, out_channels, normalize=False, concat=False, bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.concat = normalize, concat
        self._make_params(2 * in_channels if concat else in_channels, out_channels, bias)
        self.reset_parameters()

    def forward(self, x, edge_index, id, edge_weight=None, size=None, res_n_id=None):
        if edge_weight is not None or size is not None:
            raise NotImplementedError('weighted / bipartite SAGEID is not used by GraphGym')
        n, k = x.size(0), self.in_channels
        info = get_id_index(id, n)
        if self.concat:
            mean = F_.aggregate(x, get_layout(edge_index, n, ops.LOOPS_KEEP), 'mean')
            # [x | mean] W + onehot(id) [x | mean] W_id without materialising the concat
            out = F_.seg_linear([x, mean], [self.weight, self.weight_id],
                                [(0, 0, False), (1, 0, False), (0, 1, True), (1, 1, True)], info,
                                self.bias, w_rows=((0, k), (k, 2 * k), (0, k), (k, 2 * k)))
        else:
            mean = F_.aggregate(x, get_layout(edge_index, n, ops.LOOPS_ADD_REMAINING), 'mean')
            out = F_.id_linear(mean, self.weight, self.weight_id, info, self.bias)
        if self.normalize:
            out = torch.nn.functional.normalize(out, p=2, dim=-1)
        return out


class GATIDConvLayer(_IDBase):
    """ref: idconv.py:266-347 — heterogeneous transform, then additive attention with the edge-softmax
    fused into the aggregation (functional.gat_aggregate).  Dropout on alpha is p=0 in GraphGym."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0,
                 bias=True, **kwargs):
        super().__init__()
        if not concat or dropout != 0:
            raise NotImplementedError('concat=False / attention dropout are never used by GraphGym '
                                      '(ref: idconv.py:421); not on the accelerated path')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.heads, self.concat, self.negative_slope, self.dropout = heads, concat, negative_slope, dropout
        self._make_params(in_channels, heads * out_channels, bias)
        self.att = Parameter(torch.empty(1, heads, 2 * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        super().reset_parameters()
        if hasattr(self, 'att'):
            glorot_(self.att)

    def forward(self, x, edge_index, id, size=None):
        if size is not None:
            raise NotImplementedError('bipartite GATID is not used by GraphGym')
        n = x.size(0)
        h = F_.id_linear(x, self.weight, self.weight_id, get_id_index(id, n))
        # remove_self_loops + add_self_loops (ref: idconv.py:302-304)
        layout = get_layout(edge_index, n, ops.LOOPS_REMOVE_ADD)
        return F_.gat_aggregate(h, self.att, self.bias, layout, self.heads, self.negative_slope)

    def __repr__(self):
        return '{}({}, {}, heads={})'.format(self.__class__.__name__, self.in_channels,
                                             self.out_channels, self.heads)


class GINIDConvLayer(nn.Module):
    """ref: idconv.py:350-382 — z = (1+eps) x + sum_j x_j on the loop-free graph; nn(z) everywhere,
    nn_id(z[id]) added on the centre rows."""

    def __init__(self, nn, nn_id, eps=0, train_eps=False, **kwargs):
        super().__init__()
        self.nn, self.nn_id = nn, nn_id
        self.initial_eps = eps
        if train_eps:
            raise NotImplementedError('train_eps=True is never used by GraphGym (ref: idconv.py:436)')
        self.register_buffer('eps', torch.Tensor([eps]))

    def forward(self, x, edge_index, id):
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        n = x.size(0)
        z = F_.aggregate(x, get_layout(edge_index, n, ops.LOOPS_REMOVE), 'sum',
                         self_scale=1.0 + float(self.initial_eps))
        ids = get_id_index(id, n).ids
        out = _mlp(self.nn, z)
        out_id = _mlp(self.nn_id, F_.gather_rows(z, ids))
        return F_.scatter_add_rows(out, ids, out_id)

    def __repr__(self):
        return '{}(nn={})'.format(self.__class__.__name__, self.nn)


def _mlp(seq, x):
    """Sequential(Linear, ReLU, Linear, ...) with each ReLU fused into the preceding GEMM epilogue."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if not isinstance(m, nn.Linear):
            raise NotImplementedError('GIN MLPs are Linear/ReLU stacks (ref: idconv.py:432-435)')
        relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
        x = F_.linear(x, m.weight, m.bias, ops.ACT_RELU if relu else ops.ACT_NONE)
        i += 2 if relu else 1
    return x


# ---- batch wrappers: forward(batch) -> batch (ref: idconv.py:385-441) ----------------------------
class _BatchWrapper(nn.Module):
    def forward(self, batch):
        batch.node_feature = self.model(batch.node_feature, batch.edge_index, batch.node_id_index)
        return batch


class GeneralIDConv(_BatchWrapper):
    def __init__(self, dim_in, dim_out, bias=False, **kwargs):
        super().__init__()
        self.model = GeneralIDConvLayer(dim_in, dim_out, bias=bias)


class GCNIDConv(_BatchWrapper):
    def __init__(self, dim_in, dim_out, bias=False, **kwargs):
        super().__init__()
        self.model = GCNIDConvLayer(dim_in, dim_out, bias=bias)


class SAGEIDConv(_BatchWrapper):
    def __init__(self, dim_in, dim_out, bias=False, **kwargs):
        super().__init__()
        self.model = SAGEIDConvLayer(dim_in, dim_out, bias=bias, concat=True)


class GATIDConv(_BatchWrapper):
    def __init__(self, dim_in, dim_out, bias=False, **kwargs):
        super().__init__()
        self.model = GATIDConvLayer(dim_in, dim_out, bias=bias)


class GINIDConv(_BatchWrapper):
    def __init__(self, dim_in, dim_out, bias=False, **kwargs):
        super().__init__()

        def mlp():
            return nn.Sequential(nn.Linear(dim_in, dim_out), nn.ReLU(), nn.Linear(dim_out, dim_out))

        self.model = GINIDConvLayer(mlp(), mlp())


register_layer('idconv', GeneralIDConv)
register_layer('gcnidconv', GCNIDConv)
register_layer('sageidconv', SAGEIDConv)
register_layer('gatidconv', GATIDConv)
register_layer('ginidconv', GINIDConv)
