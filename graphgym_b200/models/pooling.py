"""Graph-level pooling on the GPU (ref: graphgym/models/pooling.py:12-41).

``global_add_pool / global_mean_pool / global_max_pool(x, batch, id=None, size=None)`` keep the reference's
signature and meaning — ``scatter(x, batch, dim=0, dim_size=size, reduce=...)``, applied to the centre rows
``x[id], batch[id]`` when ``cfg.dataset.transform == 'ego'`` (ID-GNN Full) — over ``gg_segment_pool_f32``
(csrc/pool.cu): ``batch`` is non-decreasing in a DeepSNAP batch, so segments come from a binary search and every
(graph, column) is reduced in row order (deterministic; torch_scatter's atomics are not).  Differentiable.
"""
import torch

from graphgym_b200 import ops
from graphgym_b200.config import cfg
from graphgym_b200.ops import _ptr, _stream, check, lib

_MODES = {'add': 0, 'mean': 1, 'max': 2}


class _SegmentPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, key, row_index, size, mode):
        ops._need_cuda(x, key, row_index)
        x, ldx = ops._rows(x, 'x')
        key = key.contiguous().long()
        m, f = key.numel(), x.size(1)
        dev = x.device
        L = lib()
        seg_ptr = torch.empty(size + 1, dtype=torch.int32, device=dev)
        bad = torch.empty(1, dtype=torch.int32, device=dev)
        check(L.gg_segment_bounds_i64(_ptr(key), m, size, _ptr(seg_ptr), _ptr(bad), _stream()), 'gg_segment_bounds_i64')
        if int(bad.item()):
            raise ValueError('pooling: `batch` must be non-decreasing with values in [0, size) '
                             '(a block-diagonal batch); got an unsorted or out-of-range index')
        out = torch.empty((size, f), dtype=torch.float32, device=dev)
        argmax = torch.empty((size, f), dtype=torch.int32, device=dev) if mode == 2 else None
        if row_index is not None:
            row_index = row_index.contiguous().long()
        check(L.gg_segment_pool_f32(_ptr(x), ldx, _ptr(row_index), _ptr(seg_ptr), size, f, mode, _ptr(out), max(f, 1),
                                    _ptr(argmax), _stream()), 'gg_segment_pool_f32')
        ctx.saved = (key, row_index, seg_ptr, argmax, x.size(0), mode)
        return out

    @staticmethod
    def backward(ctx, g):
        key, row_index, seg_ptr, argmax, n, mode = ctx.saved
        g, ldg = ops._rows(g.contiguous(), 'g')
        f = g.size(1)
        gx = torch.zeros((n, f), dtype=torch.float32, device=g.device)   # rows outside `id` get no gradient
        check(lib().gg_segment_pool_bwd_f32(_ptr(g), ldg, _ptr(row_index), _ptr(key), _ptr(seg_ptr), key.numel(), f,
                                            mode, _ptr(argmax), _ptr(gx), max(f, 1), _stream()),
              'gg_segment_pool_bwd_f32')
        return gx, None, None, None, None


def _pool(x, batch, id, size, mode):
    size = int(batch.max().item()) + 1 if size is None else int(size)
    row_index = None
    if cfg.dataset.transform == 'ego':            # ref: pooling.py:15-17 — pool the centre nodes only
        row_index = id
        batch = torch.index_select(batch, dim=0, index=id)
    return _SegmentPool.apply(x, batch, row_index, size, _MODES[mode])


def global_add_pool(x, batch, id=None, size=None):
    return _pool(x, batch, id, size, 'add')


def global_mean_pool(x, batch, id=None, size=None):
    return _pool(x, batch, id, size, 'mean')


def global_max_pool(x, batch, id=None, size=None):
    return _pool(x, batch, id, size, 'max')


pooling_dict = {'add': global_add_pool, 'mean': global_mean_pool, 'max': global_max_pool}
