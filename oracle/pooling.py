"""CPU restatement of the pooling path (TEST INFRASTRUCTURE ONLY).

``scatter`` restates ``torch_scatter.scatter(src, index, dim=0, dim_size, reduce)`` as the reference calls it
(graphgym/models/pooling.py:17,25,33; torch_scatter is not vendored and unpinned — semantics from its
documentation: 'add' sums, 'mean' divides by the group's count (0 for an empty group), 'max' takes the maximum and
leaves 0 in an empty group; the gradient of 'max' goes to the arg-max element).  ``global_pool`` follows
pooling.py:12-33 including the ``cfg.dataset.transform == 'ego'`` select of the centre rows.
"""
import torch


def scatter(src, index, dim=0, dim_size=None, reduce='add'):
    assert dim == 0
    size = int(index.max()) + 1 if dim_size is None else int(dim_size)
    f = src.size(1)
    idx = index.view(-1, 1).expand(-1, f)
    out = torch.zeros((size, f), dtype=src.dtype)
    if reduce in ('add', 'sum'):
        return out.scatter_add(0, idx, src)
    if reduce == 'mean':
        total = out.scatter_add(0, idx, src)
        count = torch.zeros(size, dtype=src.dtype).scatter_add(0, index, torch.ones_like(index, dtype=src.dtype))
        return total / count.clamp(min=1).view(-1, 1)
    if reduce == 'max':
        return out.scatter_reduce(0, idx, src, reduce='amax', include_self=False)
    raise ValueError(reduce)


def global_pool(x, batch, ids, size, mode, ego):
    if ego:
        x = torch.index_select(x, 0, ids)
        batch = torch.index_select(batch, 0, ids)
    return scatter(x, batch, dim=0, dim_size=size, reduce=mode)
