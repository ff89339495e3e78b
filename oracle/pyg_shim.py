"""Stand-ins for the third-party modules the reference imports on the hot path, so that the
reference's OWN files can be executed in the build container to generate golden vectors
(tests/golden/make_golden.py).  TEST INFRASTRUCTURE ONLY; never imported by graphgym_b200.

What is shimmed (all restated from PyG 1.x docs, see oracle/pyg_utils.py):
  torch_scatter.scatter_add
  torch_geometric.nn.conv.MessagePassing   (<= 1.5 API: x_i/x_j, edge_index_i, size_i kwargs routing,
                                            flow source_to_target, aggr add/mean)
  torch_geometric.utils.{add_remaining_self_loops, remove_self_loops, add_self_loops, softmax,
                         negative_sampling (never called on this path)}
  torch_geometric.nn.inits.{glorot, zeros, reset}
  graphgym.config.cfg                      (only cfg.gnn.agg / cfg.gnn.normalize_adj are read)
What is NOT shimmed: graphgym/register.py, graphgym/contrib/layer/idconv.py,
graphgym/contrib/transform/identity.py, graphgym/models/transform.py — loaded from the reference tree.
"""
import importlib.util
import inspect
import os
import sys
import types

import torch

from . import pyg_utils

_SPECIAL = ('edge_index', 'edge_index_i', 'edge_index_j', 'size', 'size_i', 'size_j')


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target', node_dim=0):
        super().__init__()
        assert aggr in ('add', 'mean') and flow == 'source_to_target' and node_dim == 0
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._msg_args = list(inspect.signature(self.message).parameters)
        self._upd_args = list(inspect.signature(self.update).parameters)[1:]

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = 1, 0  # aggregate at edge_index[1], gather from edge_index[0]
        size = [None, None] if size is None else list(size)
        for name in self._msg_args:
            if name[-2:] in ('_i', '_j') and name not in _SPECIAL:
                data = kwargs.get(name[:-2])
                if torch.is_tensor(data):
                    idx = i if name.endswith('_i') else j
                    if size[idx] is None:
                        size[idx] = data.size(0)
        size[0] = size[1] if size[0] is None else size[0]
        size[1] = size[0] if size[1] is None else size[1]
        args = []
        for name in self._msg_args:
            if name == 'edge_index':
                args.append(edge_index)
            elif name in ('edge_index_i', 'edge_index_j'):
                args.append(edge_index[i if name.endswith('_i') else j])
            elif name == 'size':
                args.append(size)
            elif name in ('size_i', 'size_j'):
                args.append(size[i if name.endswith('_i') else j])
            elif name[-2:] in ('_i', '_j'):
                data = kwargs.get(name[:-2])
                if data is None:
                    args.append(None)
                else:
                    assert torch.is_tensor(data), 'bipartite inputs are not on the hot path'
                    args.append(data.index_select(0, edge_index[i if name.endswith('_i') else j]))
            else:
                args.append(kwargs.get(name))
        out = self.message(*args)
        out = pyg_utils.propagate(edge_index, out, size[i], self.aggr)
        kwargs = dict(kwargs, edge_index=edge_index, size=size)
        return self.update(out, *[kwargs[name] for name in self._upd_args])

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _Cfg(types.SimpleNamespace):
    pass


def install(reference_root='/root/reference', agg='add', normalize_adj=False):
    """Install the shims and import the reference's hot-path files.  Returns a namespace with
    ``idconv``, ``identity``, ``transform``, ``register`` (the reference modules) and ``cfg``."""
    if not os.path.isdir(reference_root):
        raise FileNotFoundError(reference_root)

    def _no_negative_sampling(*a, **k):
        raise NotImplementedError('negative_sampling is not on the hot path')

    _module('torch_scatter', scatter_add=pyg_utils.scatter_add)
    tg = _module('torch_geometric')
    tg.nn = _module('torch_geometric.nn')
    tg.nn.conv = _module('torch_geometric.nn.conv', MessagePassing=MessagePassing)
    tg.nn.inits = _module('torch_geometric.nn.inits', glorot=pyg_utils.glorot, zeros=pyg_utils.zeros,
                          reset=pyg_utils.reset)
    tg.utils = _module('torch_geometric.utils',
                       add_remaining_self_loops=pyg_utils.add_remaining_self_loops,
                       remove_self_loops=pyg_utils.remove_self_loops,
                       add_self_loops=pyg_utils.add_self_loops, softmax=pyg_utils.softmax,
                       negative_sampling=_no_negative_sampling)
    cfg = _Cfg(gnn=_Cfg(agg=agg, normalize_adj=normalize_adj, self_msg='concat', att_heads=1, att_final_linear=False, att_final_linear_bn=False))
    gg = _module('graphgym')
    gg.__path__ = []  # a package, but nothing resolves through the filesystem
    gg.config = _module('graphgym.config', cfg=cfg)
    root = os.path.join(reference_root, 'graphgym')
    reg = _load('graphgym.register', os.path.join(root, 'register.py'))
    reg.layer_dict.clear()
    ns = types.SimpleNamespace(cfg=cfg, register=reg)
    ns.idconv = _load('graphgym.contrib.layer.idconv', os.path.join(root, 'contrib/layer/idconv.py'))
    ns.identity = _load('graphgym.contrib.transform.identity',
                        os.path.join(root, 'contrib/transform/identity.py'))
    ns.transform = _load('graphgym.models.transform', os.path.join(root, 'models/transform.py'))
    return ns
