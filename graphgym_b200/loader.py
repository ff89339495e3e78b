"""On-device mini-batch collation.

The reference collates on the host every iteration — ``DataLoader(dataset, collate_fn=Batch.collate(), ...)``
(ref: graphgym/loader.py:245-250; DeepSNAP's ``Batch.collate`` concatenates the graphs block-diagonally, adding the running
node count to every ``*_index`` tensor) — and then copies the batch to the device (ref: graphgym/train.py:21).  Here the
dataset's tensors live in HBM and a batch is assembled by two segmented-copy kernels (csrc/collate.cu), one launch per
output array whatever the number of graphs; ego-net batches produced on the device (models/transform.py) feed the layers
without a host round trip.
"""
import ctypes

import torch

from . import ops
from ._lib import check, lib


class GraphData:
    """One graph: the attributes DeepSNAP's ``Graph`` carries on this path (node_feature [n, f], edge_index [2, E] int64,
    optional node_label / node_id_index / augmented feature blocks such as node_identity)."""

    def __init__(self, **tensors):
        self.__dict__.update(tensors)

    @property
    def num_nodes(self):
        return int(self.__dict__.get('_num_nodes', self.node_feature.size(0)))

    def keys(self):
        return [k for k, v in self.__dict__.items() if torch.is_tensor(v)]

    def __getitem__(self, k):
        return self.__dict__[k]

    def __setitem__(self, k, v):
        self.__dict__[k] = v


class CollatedBatch(GraphData):
    """Block-diagonal batch: same attribute names as the graphs plus ``batch`` (graph id per node), ``num_graphs`` and
    ``node_ptr`` ([G+1] node offsets)."""


_DT = {torch.float32: 0, torch.int64: 1, torch.uint8: 2, torch.bool: 2}


def _is_index(key, t):
    return 'index' in key and t.dtype == torch.int64      # DeepSNAP: keys containing "index" are offset by num_nodes


def collate(graphs):
    """list of GraphData (CUDA tensors) -> CollatedBatch.  Concatenation rules follow DeepSNAP's Batch:
    ``*_index`` int64 tensors are concatenated along their LAST dimension with the running node count added;
    everything else along dimension 0 unchanged; ``batch`` is the graph id of every node."""
    if not graphs:
        raise ValueError('collate: empty batch')
    keys = graphs[0].keys()
    dev = graphs[0][keys[0]].device
    for g in graphs:
        for k in keys:
            ops._need_cuda(g[k])
    G = len(graphs)
    n_off = [0]
    for g in graphs:
        n_off.append(n_off[-1] + g.num_nodes)
    # ---- one table for every output array: 40-byte records (gg_index_segment / gg_rows_segment), uploaded once ---------
    table, jobs, keep = [], [], []    # jobs: (kind, first record, records, destination, total, src_dtype)
    out = CollatedBatch(num_graphs=G, _num_nodes=n_off[-1])

    def index_job(views, adds, dst):
        first, off = len(table), 0
        for v, add in zip(views, adds):
            table.append((v.data_ptr() if v.numel() else 0, v.numel(), off, add, 0))
            off += v.numel()
        jobs.append(('index', first, len(views), dst, off, 0))

    for k in keys:
        t0 = graphs[0][k]
        tens = [g[k].contiguous() for g in graphs]
        keep.extend(tens)
        if t0.dtype == torch.int64:
            adds = [n_off[i] if _is_index(k, t0) else 0 for i in range(G)]
            if t0.dim() == 2 and _is_index(k, t0):       # [r, E_i]: concatenate along dim 1, row by row
                total = sum(t.size(1) for t in tens)
                res = torch.empty((t0.size(0), max(total, 1)), dtype=torch.int64, device=dev)
                for r in range(t0.size(0)):
                    index_job([t[r] for t in tens], adds, res[r])
                out[k] = res[:, :total] if total == res.size(1) else res[:, :total].contiguous()
            elif t0.dim() == 1:
                total = sum(t.numel() for t in tens)
                res = torch.empty(max(total, 1), dtype=torch.int64, device=dev)[:total]
                index_job(tens, adds, res)
                out[k] = res
            else:
                raise NotImplementedError(f'collate: int64 attribute {k!r} with shape {tuple(t0.shape)}')
        elif t0.dtype in _DT and t0.dim() in (1, 2):
            f = t0.size(1) if t0.dim() == 2 else 1
            rows = sum(t.size(0) for t in tens)
            res = torch.empty((max(rows, 1), f), dtype=torch.float32, device=dev)[:rows]
            first, r_off = len(table), 0
            for t in tens:
                table.append((t.data_ptr() if t.numel() else 0, f, t.size(0), f, r_off))
                r_off += t.size(0)
            jobs.append(('rows', first, G, res, rows * f, _DT[t0.dtype]))
            out[k] = res      # 1-D float attributes become [n, 1] columns (ref: feature_augment.py:168-170)
        else:
            raise NotImplementedError(f'collate: attribute {k!r} of dtype {t0.dtype}')
    # the `batch` vector: graph id per node (src null, add = graph number)
    batch_vec = torch.empty(max(n_off[-1], 1), dtype=torch.int64, device=dev)[:n_off[-1]]
    first = len(table)
    for i, g in enumerate(graphs):
        table.append((0, g.num_nodes, n_off[i], i, 0))
    jobs.append(('index', first, G, batch_vec, n_off[-1], 0))
    out['batch'] = batch_vec
    out['node_ptr'] = torch.tensor(n_off, dtype=torch.int64, device=dev)

    tab = torch.tensor(table, dtype=torch.int64).to(dev, non_blocking=True)
    scratch = torch.empty(max(len(table), 1), dtype=torch.int64, device=dev)
    L, st = lib(), ops._stream()
    for kind, first, records, dst, total, src_dtype in jobs:
        seg_ptr = ctypes.c_void_p(tab.data_ptr() + first * 40)
        ends = ctypes.c_void_p(scratch.data_ptr() + first * 8)
        if kind == 'index':
            check(L.gg_collate_index_i64(seg_ptr, records, total, ops._ptr(dst), ends, st), 'gg_collate_index_i64')
        else:
            check(L.gg_collate_rows_f32(seg_ptr, records, total, src_dtype, ops._ptr(dst), dst.size(1), 0, ends, st),
                  'gg_collate_rows_f32')
    out.__dict__['_keep'] = keep + [tab, scratch]   # sources and tables stay alive until the kernels have run
    return out


def concat_columns(blocks, out=None):
    """Preprocess's ``torch.cat([batch[name].float() ...], dim=1)`` (ref: feature_augment.py:329-333) as one launch per
    block: every block [n, d_k] (float32 / int64 / uint8 / bool; 1-D = one column) lands at its column offset."""
    dev = blocks[0].device
    n = blocks[0].size(0)
    widths = [b.size(1) if b.dim() == 2 else 1 for b in blocks]
    if out is None:
        out = torch.empty((max(n, 1), sum(widths)), dtype=torch.float32, device=dev)[:n]
    table = []
    keep = []
    for b, w in zip(blocks, widths):
        ops._need_cuda(b)
        if b.size(0) != n or b.dtype not in _DT:
            raise ValueError(f'concat_columns: block {tuple(b.shape)} {b.dtype}')
        b = b.contiguous()
        keep.append(b)
        table.append((b.data_ptr() if b.numel() else 0, w, n, w, 0))
    tab = torch.tensor(table, dtype=torch.int64).to(dev, non_blocking=True)
    scratch = torch.empty(len(table), dtype=torch.int64, device=dev)
    L, st, col = lib(), ops._stream(), 0
    for i, (b, w) in enumerate(zip(keep, widths)):
        check(L.gg_collate_rows_f32(ctypes.c_void_p(tab.data_ptr() + i * 40), 1, n * w, _DT[b.dtype], ops._ptr(out),
                                    out.size(1), col, ctypes.c_void_p(scratch.data_ptr() + i * 8), st), 'gg_collate_rows_f32')
        col += w
    return out
