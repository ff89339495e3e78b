"""Balanced label binning on the GPU.

The reference turns a scalar node statistic (the node task's label: ``node_clustering_coefficient``) into class labels with
numpy on the host, over the whole dataset: ``_get_bin_edges(..., 'balanced')`` takes ``feature_dim`` order statistics of
the sorted values and de-duplicates them (ref: graphgym/models/feature_augment.py:219-231), ``_bin_features`` labels every
node with ``np.digitize(arr, bin_edges) - 1`` (ref: feature_augment.py:134-143).  Both run here in float64 on the device
(csrc/binning.cu over the radix sort of the layout build); only the ``feature_dim`` edge values cross to the host.
"""
import numpy as np
import torch

from graphgym_b200 import ops
from graphgym_b200.ops import _ptr, _stream, check, lib


def argsort_f64(x):
    """Ascending, stable order of a float64 CUDA vector (two LSD passes of the u32 pair sort)."""
    ops._need_cuda(x)
    if x.dtype != torch.float64 or x.dim() != 1:
        raise ValueError('argsort_f64: expected a 1-D float64 tensor')
    x = x.contiguous()
    n, dev = x.numel(), x.device
    L = lib()
    u32 = lambda: torch.empty(max(n, 1), dtype=torch.int32, device=dev)[:n]
    hi, lo, idx = u32(), u32(), u32()
    check(L.gg_f64_sort_keys(_ptr(x), n, _ptr(hi), _ptr(lo), _ptr(idx), _stream()), 'gg_f64_sort_keys')
    _, order = ops.sort_pairs(lo, idx, 32)
    hi2 = u32()
    check(L.gg_gather_u32(_ptr(hi), _ptr(order), n, _ptr(hi2), _stream()), 'gg_gather_u32')
    _, order = ops.sort_pairs(hi2, order, 32)
    return order


def balanced_bin_edges(values, feature_dim):
    """ref: feature_augment.py:219-231 — sorted_arr[linspace(0, len, num=feature_dim, endpoint=False).astype(int)], unique."""
    n = values.numel()
    order = argsort_f64(values)
    pos = np.linspace(0, n, num=feature_dim, endpoint=False).astype(int)
    picks = order[torch.from_numpy(pos).to(order.device)].long()        # feature_dim indices
    edges = values[picks].cpu().numpy()                                  # the only values that cross to the host
    return np.unique(edges)


def digitize(values, bin_edges):
    """np.digitize(values, bin_edges) - 1 as int64 labels (ref: feature_augment.py:139-143)."""
    ops._need_cuda(values)
    values = values.contiguous()
    bins = torch.as_tensor(np.asarray(bin_edges, dtype=np.float64), device=values.device)
    out = torch.empty(max(values.numel(), 1), dtype=torch.int64, device=values.device)[:values.numel()]
    check(lib().gg_digitize_f64(_ptr(values), values.numel(), _ptr(bins), bins.numel(), _ptr(out), _stream()),
          'gg_digitize_f64')
    return out


def balanced_labels(values, feature_dim):
    """-> (labels int64 [n], bin_edges float64 numpy): the reference's 'balanced' label representation."""
    edges = balanced_bin_edges(values, feature_dim)
    return digitize(values, edges), edges
