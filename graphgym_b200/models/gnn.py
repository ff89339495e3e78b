"""The callers either side of the message-passing stage, as GraphGym stacks them (ref: graphgym/models/gnn.py:21-81,
139-174, graphgym/models/head.py:15-43): ``GNNPreMP`` (linear layers), ``GNNStackStage`` (cfg.gnn.layers_mp GeneralLayers
of cfg.gnn.layer_type, L2-normalised at the end of the stage when cfg.gnn.l2norm) and ``GNNNodeHead`` (MLP on the rows
``node_id_index`` selects).  Kept thin: everything heavy is the layers' gg_* path; the stage's final L2 normalisation is
folded into the last layer's fused post-op pass instead of running as a separate F.normalize.
"""
import torch
import torch.nn as nn

from graphgym_b200 import functional as F_
from graphgym_b200 import ops
from graphgym_b200.config import cfg
from graphgym_b200.models.layer import TFG_ALIASES, GeneralLayer


class GNNStackStage(nn.Module):
    """ref: gnn.py:65-81 — ``layer{i}`` children, then F.normalize(p=2, dim=-1) if cfg.gnn.l2norm."""

    def __init__(self, dim_in, dim_out, num_layers, layer_type=None):
        super().__init__()
        name = layer_type or cfg.gnn.layer_type
        name = TFG_ALIASES.get(name, name)
        for i in range(num_layers):
            d_in = dim_in if i == 0 else dim_out
            last = i == num_layers - 1
            self.add_module('layer{}'.format(i), GeneralLayer(name, d_in, dim_out, has_act=True,
                                                              has_l2norm=bool(cfg.gnn.l2norm) and last))
        self.dim_out = dim_out

    def forward(self, batch):
        for layer in self.children():
            batch = layer(batch)
        return batch


class _Linear(nn.Module):
    """``linear`` of the reference's layer_dict (ref: models/layer.py:73-86) on the gg GEMM."""

    def __init__(self, dim_in, dim_out, bias=False, **kwargs):
        super().__init__()
        self.model = nn.Linear(dim_in, dim_out, bias=bias)

    def forward(self, batch):
        if isinstance(batch, torch.Tensor):
            return F_.linear(batch, self.model.weight, self.model.bias)
        batch.node_feature = F_.linear(batch.node_feature, self.model.weight, self.model.bias)
        return batch


class GNNPreMP(nn.Module):
    """ref: gnn.py:40-62 (GeneralMultiLayer('linear', ...)): Linear -> BN -> act per layer, fused post-ops."""

    def __init__(self, dim_in, dim_out, num_layers=1):
        super().__init__()
        from graphgym_b200.models import layer as L
        L.layer_dict.setdefault('linear', _Linear)
        for i in range(num_layers):
            self.add_module('Layer_{}'.format(i), GeneralLayer('linear', dim_in if i == 0 else dim_out, dim_out, True))
        self.dim_out = dim_out

    def forward(self, batch):
        for layer in self.children():
            batch = layer(batch)
        return batch


class GNNNodeHead(nn.Module):
    """ref: head.py:15-43 — the prediction MLP on the rows ``node_id_index`` selects (the centre copies of the ego-nets)
    or on all rows; returns (pred, label)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.layer_post_mp = nn.Linear(dim_in, dim_out, bias=True)

    def forward(self, batch):
        h = batch.node_feature
        ids = getattr(batch, 'node_id_index', None)
        if ids is not None:
            h = F_.gather_rows(h, ids)
        pred = F_.linear(h, self.layer_post_mp.weight, self.layer_post_mp.bias)
        return pred, getattr(batch, 'node_label', None)


class GNN(nn.Module):
    """ref: gnn.py:139-174 — pre_mp -> mp stage -> post_mp head, dimensions from cfg.gnn."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        d = cfg.gnn.dim_inner
        if cfg.gnn.layers_pre_mp > 0:
            self.pre_mp = GNNPreMP(dim_in, d, cfg.gnn.layers_pre_mp)
            dim_in = d
        if cfg.gnn.layers_mp > 0:
            self.mp = GNNStackStage(dim_in, d, cfg.gnn.layers_mp)
            dim_in = d
        self.post_mp = GNNNodeHead(dim_in, dim_out)

    def forward(self, batch):
        for module in self.children():
            batch = module(batch)
        return batch
