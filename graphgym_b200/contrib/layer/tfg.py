"""The TensorFlow ID-GNN layers of the reference's second front door (TfgIDLayer.py, driven by main_zd.py:52-242) on the
accelerated path, with THEIR semantics — which differ from the PyG-side layers of idconv.py:

  TfgIDGCN   gcn_id (TfgIDLayer.py:478-566): append N self loops (no removal), degree over the rows, in-layer activation
  TfgIDSAGE  IDSAGE (TfgIDLayer.py:15-120): [x W_self (+ W_id on the centres) | mean_j(x_j) W_nb], units / 2 each, activation
  TfgIDGIN   IDGIN (TfgIDLayer.py:123-167): (1 + eps) x + sum_j x_j WITHOUT removing self loops, mlp / mlp_id
  TfgIDGAT   gat_id (TfgIDLayer.py:269-388): scaled dot-product attention <relu(x Wq + bq)_i, relu(x Wk + bk)_j> / sqrt(d),
             softmax over the row's edges (self loops appended), values with the ID transform, in-layer activation

Tfg aggregates at row = edge_index[0] and gathers col = edge_index[1] (sparse_adj.py:91-97) — the transposed flow of the
PyG layers; the layouts are therefore built on the flipped edge list.  Call shapes: ``layer([x, edge_index, id_index])``
(the Keras call of TfgIDLayer.py:74-84) and ``layer(batch)`` (GraphGym's).  Registered as ``Tfg-idgcn / Tfg-idsage /
Tfg-idgin / Tfg-idgat`` (main_zd.py:299-308).  num_heads = 1, no attention dropout, no edge_weight (main_zd never sets them).
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
from torch.nn import Parameter

from graphgym_b200 import functional as F_
from graphgym_b200 import ops
from graphgym_b200.contrib.layer.idconv import _mlp, glorot_, zeros_
from graphgym_b200.graph import get_id_index, get_layout
from graphgym_b200.register import register_layer

_FLIPPED = OrderedDict()


def _flipped(edge_index):
    """edge_index with its two rows swapped (row <-> col), cached per tensor identity + version."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, edge_index.device.index)
    hit = _FLIPPED.get(key)
    if hit is not None and hit[0] is edge_index:
        return hit[1]
    if edge_index.dtype != torch.int64:
        edge_index = edge_index.long()           # Tfg passes int32 edge lists (Graph.cast_edge_index)
    out = torch.stack([edge_index[1], edge_index[0]]).contiguous()
    _FLIPPED[key] = (edge_index, out)
    while len(_FLIPPED) > 8:
        _FLIPPED.popitem(last=False)
    return out


class _TfgBase(nn.Module):
    def forward(self, inputs, cache=None, training=None, mask=None):
        if isinstance(inputs, (list, tuple)):
            if len(inputs) == 4 and inputs[3] is not None:
                raise NotImplementedError('edge_weight is never passed by main_zd.py; not on the accelerated path')
            return self._call(inputs[0], inputs[1], inputs[2])
        batch = inputs
        batch.node_feature = self._call(batch.node_feature, batch.edge_index, batch.node_id_index)
        return batch

    @staticmethod
    def _act(h, activation):
        if activation is None:
            return h
        if activation == 'relu':
            return F_.post_ops(h, None, True, ops.ACT_RELU, 0.0, False)
        raise NotImplementedError('activation {!r}: main_zd.py uses relu'.format(activation))


class TfgIDGCN(_TfgBase):
    def __init__(self, dim_in, units, bias=True, activation='relu', **kwargs):
        super().__init__()
        self.kernel = Parameter(torch.empty(dim_in, units))
        self.kernel_id = Parameter(torch.empty(dim_in, units))
        self.bias = Parameter(torch.zeros(units)) if bias else None
        self.activation = activation
        glorot_(self.kernel); glorot_(self.kernel_id)

    def _call(self, x, edge_index, ids):
        n = x.size(0)
        h = F_.id_linear(x, self.kernel, self.kernel_id, get_id_index(ids.long(), n))
        # rows = edge_index[0]: on the flipped list that is the layout's target side; 'gcn_tgt' = degree over the rows
        layout = get_layout(_flipped(edge_index), n, ops.LOOPS_ADD)
        return self._act(F_.aggregate(h, layout, 'gcn_tgt', 0.0, self.bias), self.activation)


class TfgIDSAGE(_TfgBase):
    def __init__(self, dim_in, units, bias=True, activation='relu', concat=True, **kwargs):
        super().__init__()
        if concat and units % 2 != 0:
            raise Exception('units must be a event number if concat is True')      # TfgIDLayer.py:42-43
        ku = units // 2 if concat else units
        self.concat, self.activation, self.units = concat, activation, units
        self.self_kernel = Parameter(torch.empty(dim_in, ku))
        self.id_kernel = Parameter(torch.empty(dim_in, ku))
        self.neighbor_kernel = Parameter(torch.empty(dim_in, ku))
        self.bias = Parameter(torch.zeros(units)) if bias else None
        for p in (self.self_kernel, self.id_kernel, self.neighbor_kernel):
            glorot_(p)

    def _call(self, x, edge_index, ids):
        n = x.size(0)
        info = get_id_index(ids.long(), n)
        mean = F_.aggregate(x, get_layout(_flipped(edge_index), n, ops.LOOPS_KEEP), 'mean')
        if self.concat:
            ku = self.units // 2
            left = F_.seg_linear([x], [self.self_kernel, self.id_kernel], [(0, 0, False), (0, 1, True)], info,
                                 self.bias[:ku] if self.bias is not None else None)
            right = F_.seg_linear([mean], [self.neighbor_kernel], [(0, 0, False)], None,
                                  self.bias[ku:] if self.bias is not None else None)
            h = torch.cat([left, right], 1)            # tf.concat of the two halves (TfgIDLayer.py:109-110)
        else:
            h = F_.seg_linear([x, mean], [self.self_kernel, self.id_kernel, self.neighbor_kernel],
                              [(0, 0, False), (0, 1, True), (1, 2, False)], info, self.bias)
        return self._act(h, self.activation)


class TfgIDGIN(_TfgBase):
    def __init__(self, mlp_model, mlpid_model, eps=0.0, train_eps=False, **kwargs):
        super().__init__()
        if train_eps:
            raise NotImplementedError('train_eps is never set by main_zd.py')
        self.mlp_model, self.mlp_id, self.eps = mlp_model, mlpid_model, float(eps)

    def _call(self, x, edge_index, ids):
        n = x.size(0)
        h = F_.aggregate(x, get_layout(_flipped(edge_index), n, ops.LOOPS_KEEP), 'sum', self_scale=1.0 + self.eps)
        out = _mlp(self.mlp_model, h)
        ids = ids.long()
        return F_.scatter_add_rows(out, ids, _mlp(self.mlp_id, F_.gather_rows(h, ids)))


class _DotAttention(torch.autograd.Function):
    """out[i] = sum_j softmax_i(<Q_i, K_j> / sqrt(d)) V_j over the layout's rows (+ bias)."""

    @staticmethod
    def forward(ctx, q, k, v, bias, layout):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        csr = layout.csr
        scale = 1.0 / math.sqrt(q.size(1))
        z = ops.sddmm(csr, k, q)                                      # z[s] = <q[row(s)], k[nbr[s]]>
        alpha = ops.segment_softmax(csr, z, scale)
        out = ops.spmm(csr, v, alpha, ops.SUM, None, 0.0, bias)
        ctx.layout, ctx.scale, ctx.has_bias = layout, scale, bias is not None
        ctx.save_for_backward(q, k, v, alpha)
        return out

    @staticmethod
    def backward(ctx, g):
        q, k, v, alpha = ctx.saved_tensors
        lay = ctx.layout
        csr, csc, m = lay.csr, lay.csc, lay.csc2csr
        g = g.contiguous()
        dalpha = ops.sddmm(csr, v, g)
        dz = ops.segment_softmax_bwd(csr, alpha, dalpha, ctx.scale)
        alpha_t, _ = ops.gat_csc_gather(csc, m, alpha, dz)
        dz_t, _ = ops.gat_csc_gather(csc, m, dz, dz)
        dv = ops.spmm(csc, g, alpha_t)
        dq = ops.spmm(csr, k, dz)
        dk = ops.spmm(csc, q, dz_t)
        gb = ops.colsum(g) if (ctx.has_bias and ctx.needs_input_grad[3]) else None
        return dq, dk, dv, gb, None


class TfgIDGAT(_TfgBase):
    def __init__(self, dim_in, units, attention_units=None, bias=True, activation='relu', num_heads=1, **kwargs):
        super().__init__()
        if num_heads != 1:
            raise NotImplementedError('num_heads > 1 is never set by main_zd.py')
        au = units if attention_units is None else attention_units
        self.query_kernel = Parameter(torch.empty(dim_in, au))
        self.query_bias = Parameter(torch.zeros(au))
        self.key_kernel = Parameter(torch.empty(dim_in, au))
        self.key_bias = Parameter(torch.zeros(au))
        self.kernel = Parameter(torch.empty(dim_in, units))
        self.kernel_id = Parameter(torch.empty(dim_in, units))
        self.bias = Parameter(torch.zeros(units)) if bias else None
        self.activation = activation
        for p in (self.query_kernel, self.key_kernel, self.kernel, self.kernel_id):
            glorot_(p)

    def _call(self, x, edge_index, ids):
        n = x.size(0)
        q = F_.linear(x, self.query_kernel, self.query_bias, ops.ACT_RELU, w_trans=False)
        k = F_.linear(x, self.key_kernel, self.key_bias, ops.ACT_RELU, w_trans=False)
        v = F_.id_linear(x, self.kernel, self.kernel_id, get_id_index(ids.long(), n))
        layout = get_layout(_flipped(edge_index), n, ops.LOOPS_ADD)
        return self._act(_DotAttention.apply(q, k, v, self.bias, layout), self.activation)


def _gin_mlp(dim_in, dim_out):
    return nn.Sequential(nn.Linear(dim_in, dim_out), nn.ReLU(), nn.Linear(dim_out, dim_out))


class _Wrap(nn.Module):
    """GraphGym constructor shape (dim_in, dim_out, bias) around a Tfg layer."""

    def forward(self, inputs, *a, **k):
        return self.model(inputs, *a, **k)


class TfgIDGCNConv(_Wrap):
    def __init__(self, dim_in, dim_out, bias=True, **kwargs):
        super().__init__()
        self.model = TfgIDGCN(dim_in, dim_out, bias=bias)


class TfgIDSAGEConv(_Wrap):
    def __init__(self, dim_in, dim_out, bias=True, **kwargs):
        super().__init__()
        self.model = TfgIDSAGE(dim_in, dim_out, bias=bias)


class TfgIDGINConv(_Wrap):
    def __init__(self, dim_in, dim_out, bias=True, **kwargs):
        super().__init__()
        self.model = TfgIDGIN(_gin_mlp(dim_in, dim_out), _gin_mlp(dim_in, dim_out))


class TfgIDGATConv(_Wrap):
    def __init__(self, dim_in, dim_out, bias=True, **kwargs):
        super().__init__()
        self.model = TfgIDGAT(dim_in, dim_out, bias=bias)


register_layer('Tfg-idgcn', TfgIDGCNConv)
register_layer('Tfg-idsage', TfgIDSAGEConv)
register_layer('Tfg-idgin', TfgIDGINConv)
register_layer('Tfg-idgat', TfgIDGATConv)
