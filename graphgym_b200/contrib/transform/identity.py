"""ID-GNN Fast cycle features on the GPU (ref: graphgym/contrib/transform/identity.py:7-35).

``compute_identity(edge_index, n, k)`` keeps the reference's signature and meaning —
``stack([diag(A_hat^1), ..., diag(A_hat^k)], dim=1)`` with A_hat = D^-1/2 (A + I) D^-1/2 (remaining
self loops, degree over edge_index[0]) — but never densifies: blocks of source nodes are propagated
through the sparse layout by ``gg_cycle_diag_*`` (csrc/cycle.cu).

``closed_walk_counts`` is the exact int64 variant diag(A^p) the north star names; the reference has no
integer mode, so its parity target is the dense int64 oracle.
"""
import os

import torch

from graphgym_b200 import ops
from graphgym_b200.ops import _ptr, _stream, check, lib


CYCLE_STEP = os.environ.get("GG_CYCLE_STEP", "sell")   # sell (sliced-ELL aggregation) | mp (merge-path) | row (one warp per row)
MP_STEP = CYCLE_STEP in ("mp", "sell")


def _ranges(n, graph_ptr, block):
    """Yield (row_begin, row_end, src_begin, src_count): source blocks never straddle more graphs than
    needed — the row range is the run of graphs that holds the block."""
    if graph_ptr is None:
        for s in range(0, n, block):
            yield 0, n, s, min(block, n - s)
        return
    gp = graph_ptr.tolist()
    g = 0
    for s in range(0, n, block):
        e = min(n, s + block)
        while gp[g + 1] <= s:
            g += 1
        h = g
        while gp[h + 1] < e:
            h += 1
        yield gp[g], gp[h + 1], s, e - s


def _run(layout_csr, w_slot, n, k, symmetric, graph_ptr, integer):
    dev = layout_csr.rowptr.device
    L = lib()
    block = 64 if integer else 128
    out = torch.empty((n, k), dtype=torch.int64 if integer else torch.float32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    gp = graph_ptr.cpu() if graph_ptr is not None else None
    # one big graph in float mode: every hop is a whole-graph aggregation -> the load-balanced merge-path kernel
    use_mp = (not integer) and gp is None and n + layout_csr.num_slots >= (1 << 14) and MP_STEP
    use_sell = use_mp and CYCLE_STEP == "sell"
    if use_sell:
        sl = ops.sell_layout(layout_csr)
        w_sell = sl.weights(w_slot)
    elif use_mp:
        item_row, item_slot, items = layout_csr.plan
    ws = None
    for rb, re, sb, sc in _ranges(n, gp, block):
        need = int(L.gg_cycle_diag_sell_workspace_bytes(n, sl.partial_rows)) if use_sell else \
            int(L.gg_cycle_diag_mp_workspace_bytes(n, items)) if use_mp else int(L.gg_cycle_diag_workspace_bytes(re - rb))
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
        dst = out[sb:sb + sc]
        if integer:
            check(L.gg_cycle_diag_i64(_ptr(layout_csr.rowptr), _ptr(layout_csr.nbr), rb, re, k, int(symmetric),
                                      sb, sc, _ptr(dst), k, _ptr(overflow), _ptr(ws), ws.numel(), _stream()),
                  "gg_cycle_diag_i64")
        elif use_sell:
            check(L.gg_cycle_diag_sell_f32(_ptr(layout_csr.rowptr), _ptr(sl.chunk_ptr), sl.chunks, _ptr(sl.idx), _ptr(w_sell),
                                           _ptr(sl.vdst), _ptr(sl.hub_rows), _ptr(sl.hub_pptr), sl.hubs, sl.partial_rows, n, k,
                                           int(symmetric), sb, sc, _ptr(dst), k, _ptr(ws), ws.numel(), _stream()),
                  "gg_cycle_diag_sell_f32")
        elif use_mp:
            check(L.gg_cycle_diag_mp_f32(_ptr(layout_csr.rowptr), _ptr(layout_csr.nbr), _ptr(w_slot), _ptr(item_row),
                                         _ptr(item_slot), items, n, k, int(symmetric), sb, sc, _ptr(dst), k, _ptr(ws),
                                         ws.numel(), _stream()), "gg_cycle_diag_mp_f32")
        else:
            check(L.gg_cycle_diag_f32(_ptr(layout_csr.rowptr), _ptr(layout_csr.nbr), _ptr(w_slot), rb, re, k,
                                      int(symmetric), sb, sc, _ptr(dst), k, _ptr(ws), ws.numel(), _stream()),
                  "gg_cycle_diag_f32")
    return out, overflow


def is_symmetric(edge_index, n):
    """True iff the multiset of (src, tgt) pairs equals the multiset of (tgt, src) pairs.  A one-off check at
    feature-augmentation time (two int64 sorts, off the per-step path): the half-power trick diag(A^2t) = <V_t, V_t> is
    only valid on a symmetric edge list, and the reference's ``compute_identity`` is correct for any."""
    a = torch.sort(edge_index[0] * int(n) + edge_index[1]).values
    b = torch.sort(edge_index[1] * int(n) + edge_index[0]).values
    return bool(torch.equal(a, b))


def compute_identity(edge_index, n, k, symmetric=None, graph_ptr=None):
    """[n, k] fp32, column p-1 = diag(A_hat^p).  ``symmetric=None`` (default) checks the edge list once and enables the
    half-power trick only for undirected graphs (as every reference dataset is); True / False skip the check.
    ``graph_ptr`` ([G+1] node offsets of a block-diagonal batch) restricts every source block to its
    own graphs."""
    ops._need_cuda(edge_index)
    if symmetric is None:
        symmetric = is_symmetric(edge_index, n)
    csr = ops.layout_build(edge_index, n, ops.LOOPS_ADD_REMAINING, ops.BY_TARGET)
    csc = ops.layout_build(edge_index, n, ops.LOOPS_ADD_REMAINING, ops.BY_SOURCE)
    deg = ops.segment_degree(csc)                 # degree over edge_index[0] (identity.py:17-18)
    w = ops.gcn_norm(csr, deg)
    out, _ = _run(csr, w, n, k, symmetric, graph_ptr, integer=False)
    return out


def closed_walk_counts(edge_index, n, k, symmetric=None, graph_ptr=None):
    """([n, k] int64 exact closed-walk counts diag(A^p), number of overflowed entries).  ``symmetric`` as above."""
    ops._need_cuda(edge_index)
    if symmetric is None:
        symmetric = is_symmetric(edge_index, n)
    csr = ops.layout_build(edge_index, n, ops.LOOPS_KEEP, ops.BY_TARGET)
    out, overflow = _run(csr, None, n, k, symmetric, graph_ptr, integer=True)
    return out, int(overflow.item())
