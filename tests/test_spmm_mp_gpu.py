"""Merge-path (load-balanced, TMA-staged) aggregation kernel: parity against the oracle and against the
one-warp-per-row kernel, on the shapes that stress the plan — hub rows split over many items, long runs
of empty rows, items that contain only markers, tiny graphs with a single item."""
import numpy as np
import pytest
import torch

from graphgym_b200 import ops
from oracle import pyg_utils as U
from util import FP32_TOL, powerlaw_graph, random_graph, rel_err

pytestmark = pytest.mark.gpu


def oracle(ei, x, w_edge, reduce, self_scale, bias):
    xj = x.double().index_select(0, ei[0])
    if w_edge is not None:
        xj = xj * w_edge.double().view(-1, 1)
    out = U.propagate(ei, xj, x.size(0), 'mean' if reduce == ops.MEAN else 'add')
    if self_scale:
        out = out + self_scale * x.double()
    if bias is not None:
        out = out + bias.double()
    return out


def run(csr, x, w, reduce, self_scale, bias, algo, stage, monkeypatch):
    monkeypatch.setattr(ops, 'SPMM_ALGO', algo)
    monkeypatch.setattr(ops, 'SPMM_STAGE', stage)
    return ops.spmm(csr, x, w, reduce, x if self_scale else None, self_scale, bias)


def check_plan(csr):
    item_row, item_slot, items = csr.plan
    L = ops.lib()
    units = int(L.gg_spmm_plan_units(csr.num_nodes, csr.num_slots))
    rp = csr.rowptr.cpu().numpy().astype(np.int64)
    ir, isl = item_row.cpu().numpy().astype(np.int64), item_slot.cpu().numpy().astype(np.int64)
    n, e = csr.num_nodes, csr.num_slots
    d = np.minimum(np.arange(items + 1, dtype=np.int64) * units, n + e)
    assert (ir + isl == d).all()                       # on the diagonal
    assert ir[-1] == n and isl[-1] == e and ir[0] == 0 and isl[0] == 0
    assert (isl >= rp[ir]).all()                       # the slot belongs to the row or is its end
    inside = ir < n
    assert (isl[inside] <= rp[ir[inside] + 1]).all()
    # minimality: the marker of row-1 lies before the diagonal
    prev = ir > 0
    assert (rp[ir[prev]] + ir[prev] - 1 < d[prev]).all()


@pytest.mark.parametrize('stage', ['tma', 'ldg'])
@pytest.mark.parametrize('f', [68, 100, 128, 136, 256, 300, 512, 1024])
def test_parity_random(cuda, monkeypatch, stage, f):
    n = 997
    ei = random_graph(f, n, 9000, loops=30, dups=40)
    g = torch.Generator().manual_seed(f)
    x = torch.randn(n, f, generator=g)
    w_edge = torch.rand(ei.size(1), generator=g)
    bias = torch.randn(f, generator=g)
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    check_plan(csr)
    w_slot = w_edge[csr.perm.cpu().long()].to(cuda)
    for weighted in (False, True):
        for reduce in (ops.SUM, ops.MEAN):
            for self_scale, b in ((0.0, None), (1.25, bias)):
                got = run(csr, x.to(cuda), w_slot if weighted else None, reduce, self_scale,
                          b.to(cuda) if b is not None else None, 'mp', stage, monkeypatch)
                want = oracle(ei, x, w_edge if weighted else None, reduce, self_scale, b)
                assert rel_err(got, want) < FP32_TOL, (f, weighted, reduce, self_scale)


@pytest.mark.parametrize('n,avg', [(20000, 16), (200000, 25)])
def test_powerlaw_hubs_split_over_items(cuda, monkeypatch, n, avg):
    f = 128
    ei = powerlaw_graph(5, n, avg)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    check_plan(csr)
    deg = ops.segment_degree(csr)
    w = ops.gcn_norm(csr, deg)
    assert int(torch.diff(csr.rowptr).max()) > 2000       # rows spanning >= 4 items
    mp = run(csr, x.to(cuda), w, ops.SUM, 0.0, None, 'mp', 'tma', monkeypatch)
    row = run(csr, x.to(cuda), w, ops.SUM, 0.0, None, 'row', 'tma', monkeypatch)
    # fp64 arbiter: the same weighted adjacency as a sparse matrix on the CPU
    a = torch.sparse_coo_tensor(torch.stack([csr.rowid.cpu().long(), csr.nbr.cpu().long()]), w.cpu().double(),
                                (n, n)).coalesce()
    want = torch.sparse.mm(a, x.double())
    assert rel_err(mp, want) < FP32_TOL
    # one warp summing a 10^4-slot hub row strictly in sequence accumulates more fp32 rounding than the
    # merge-path kernel's <= 480-slot partials (the reference's atomic scatter-add is no better)
    assert rel_err(row, want) < 5 * FP32_TOL
    # rows that live inside one item are summed in the same order by both kernels: bitwise equal
    same = (mp == row).all(dim=1).float().mean().item()
    assert same > (0.8 if n >= 200000 else 0.05)   # small graphs use 64-unit items: most rows are split
    assert torch.equal(mp, run(csr, x.to(cuda), w, ops.SUM, 0.0, None, 'mp', 'tma', monkeypatch))


def test_empty_rows_and_marker_only_items(cuda, monkeypatch):
    """20000 isolated nodes between two small cliques: most items hold nothing but row markers."""
    n, f = 20000, 128
    a = torch.combinations(torch.arange(0, 40), 2).t()
    b = torch.combinations(torch.arange(n - 30, n), 2).t()
    ei = torch.cat([a, a.flip(0), b, b.flip(0)], dim=1)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(2))
    bias = torch.randn(f, generator=torch.Generator().manual_seed(3))
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    check_plan(csr)
    for reduce in (ops.SUM, ops.MEAN):
        got = run(csr, x.to(cuda), None, reduce, 2.0, bias.to(cuda), 'mp', 'tma', monkeypatch)
        assert rel_err(got, oracle(ei, x, None, reduce, 2.0, bias)) < FP32_TOL


def test_no_edges_and_single_item(cuda, monkeypatch):
    f = 128
    for n, e in [(5, 0), (1, 0), (3, 4), (40, 10)]:
        ei = random_graph(n, n, e) if e else torch.zeros((2, 0), dtype=torch.int64)
        x = torch.randn(n, f, generator=torch.Generator().manual_seed(n))
        csr = ops.layout_build(ei.to(cuda), n, 0, 0)
        check_plan(csr)
        got = run(csr, x.to(cuda), None, ops.SUM, 1.0, None, 'mp', 'tma', monkeypatch)
        assert rel_err(got, oracle(ei, x, None, ops.SUM, 1.0, None)) < FP32_TOL


def test_one_giant_row(cuda, monkeypatch):
    """every edge points at node 7: one row of 60000 slots = 125 items chained through the fix-up."""
    n, e, f = 300, 60000, 256
    g = torch.Generator().manual_seed(0)
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.full((e,), 7)])
    x = torch.randn(n, f, generator=g)
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    check_plan(csr)
    for reduce in (ops.SUM, ops.MEAN):
        got = run(csr, x.to(cuda), None, reduce, 0.0, None, 'mp', 'tma', monkeypatch)
        assert rel_err(got, oracle(ei, x, None, reduce, 0.0, None)) < FP32_TOL


# ---- sub-warp-group kernel (narrow rows) and peer-memory output -------------------------------------
def check_group_plan(csr, f):
    lanes = int(ops.lib().gg_spmm_group_lanes(f))
    assert lanes in (4, 8, 16, 32) and lanes * 4 >= f
    check_plan(csr)


@pytest.mark.parametrize('f', [4, 8, 12, 16, 24, 32, 48, 64, 100, 128])
def test_group_kernel_parity_random(cuda, monkeypatch, f):
    n = 997
    ei = random_graph(f, n, 9000, loops=30, dups=40)
    g = torch.Generator().manual_seed(f)
    x = torch.randn(n, f, generator=g)
    w_edge = torch.rand(ei.size(1), generator=g)
    bias = torch.randn(f, generator=g)
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    check_group_plan(csr, f)
    w_slot = w_edge[csr.perm.cpu().long()].to(cuda)
    for weighted in (False, True):
        for reduce in (ops.SUM, ops.MEAN):
            for self_scale, b in ((0.0, None), (1.25, bias)):
                got = run(csr, x.to(cuda), w_slot if weighted else None, reduce, self_scale,
                          b.to(cuda) if b is not None else None, 'mpg', 'tma', monkeypatch)
                want = oracle(ei, x, w_edge if weighted else None, reduce, self_scale, b)
                assert rel_err(got, want) < FP32_TOL, (f, weighted, reduce, self_scale)


@pytest.mark.parametrize('f', [16, 32, 64, 128])
def test_group_kernel_powerlaw_and_edge_shapes(cuda, monkeypatch, f):
    n = 60000
    ei = powerlaw_graph(6, n, 20)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    check_group_plan(csr, f)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    got = run(csr, x.to(cuda), w, ops.SUM, 0.0, None, 'mpg', 'tma', monkeypatch)
    a = torch.sparse_coo_tensor(torch.stack([csr.rowid.cpu().long(), csr.nbr.cpu().long()]), w.cpu().double(),
                                (n, n)).coalesce()
    assert rel_err(got, torch.sparse.mm(a, x.double())) < FP32_TOL
    assert torch.equal(got, run(csr, x.to(cuda), w, ops.SUM, 0.0, None, 'mpg', 'tma', monkeypatch))   # deterministic
    # one giant row, marker-only items, no edges
    e = 30000
    g = torch.Generator().manual_seed(0)
    ei2 = torch.stack([torch.randint(0, 300, (e,), generator=g), torch.full((e,), 7)])
    x2 = torch.randn(300, f, generator=g)
    csr2 = ops.layout_build(ei2.to(cuda), 300, 0, 0)
    for reduce in (ops.SUM, ops.MEAN):
        got = run(csr2, x2.to(cuda), None, reduce, 0.0, None, 'mpg', 'tma', monkeypatch)
        assert rel_err(got, oracle(ei2, x2, None, reduce, 0.0, None)) < FP32_TOL
    cl = torch.combinations(torch.arange(0, 40), 2).t()
    ei3 = torch.cat([cl, cl.flip(0)], dim=1)
    x3 = torch.randn(20000, f, generator=g)
    csr3 = ops.layout_build(ei3.to(cuda), 20000, 0, 0)
    got = run(csr3, x3.to(cuda), None, ops.MEAN, 2.0, None, 'mpg', 'tma', monkeypatch)
    assert rel_err(got, oracle(ei3, x3, None, ops.MEAN, 2.0, None)) < FP32_TOL
    for nn in (1, 5):
        csr4 = ops.layout_build(torch.zeros((2, 0), dtype=torch.int64, device=cuda), nn, 0, 0)
        x4 = torch.randn(nn, f, generator=g)
        got = run(csr4, x4.to(cuda), None, ops.SUM, 1.0, None, 'mpg', 'tma', monkeypatch)
        assert rel_err(got, x4) < FP32_TOL


@pytest.mark.parametrize('world,f_total', [(2, 128), (4, 128), (8, 128), (8, 256), (3, 96)])
def test_peer_output_and_column_scatter_on_one_gpu(cuda, world, f_total):
    """The two kernels of the feature-sliced exchange with every 'rank' living on this GPU: the column
    scatter builds each rank's [N, F/P] slice, the peer-output SpMM of slice c stores rows into the
    owners' [rows_per_rank, F] blocks; stitched together they must equal the plain aggregation."""
    import ctypes
    n = 5003
    fs = f_total // world
    ei = powerlaw_graph(9, n, 12)
    g = torch.Generator().manual_seed(world)
    x = torch.randn(n, f_total, generator=g).to(cuda)
    bias = torch.randn(f_total, generator=g).to(cuda)
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    per = (n + world - 1) // world
    slices = [torch.full((n, fs), float('nan'), device=cuda) for _ in range(world)]
    L = ops.lib()
    for r in range(world):   # rank r scatters its rows
        lo, hi = min(n, r * per), min(n, (r + 1) * per)
        dst = (ctypes.c_void_p * world)(*[s.data_ptr() for s in slices])
        ops.check(L.gg_peer_scatter_cols_f32(ctypes.c_void_p(x[lo:hi].data_ptr()), f_total, hi - lo, f_total, dst,
                                             world, r, lo, ops._stream()), 'gg_peer_scatter_cols_f32')
    for c in range(world):
        assert torch.equal(slices[c], x[:, c * fs:(c + 1) * fs])
    blocks = [torch.full((per, f_total), float('nan'), device=cuda) for _ in range(world)]
    for c in range(world):   # rank c aggregates its slice and stores into the owners' blocks
        peers = ops.PeerRows([b.data_ptr() + c * fs * 4 for b in blocks], per, f_total)
        assert ops.spmm(csr, slices[c], w, ops.SUM, None, 0.0, bias[c * fs:(c + 1) * fs].clone(),
                        out_peers=peers) is None
    got = torch.cat(blocks)[:n]
    want = ops.spmm(csr, x, w, ops.SUM, None, 0.0, bias)
    assert not torch.isnan(got).any()
    assert rel_err(got, want) < FP32_TOL
    # the bulk-push return leg: local slice -> gg_peer_push_rows_f32 (whole world, then owner group by owner group) ->
    # gg_peer_gather_slices_f32 on every owner; bitwise the fused-return result
    for groups in ([(0, world)], [(o, o + 1) for o in range(world)]):
        recv = [torch.full((world * per * fs,), float('nan'), device=cuda) for _ in range(world)]
        arr = (ctypes.c_void_p * world)(*[t.data_ptr() for t in recv])
        for c in range(world):
            out_slice = ops.spmm(csr, slices[c], w, ops.SUM, None, 0.0, bias[c * fs:(c + 1) * fs].clone())
            for o0, o1 in groups:
                ops.check(L.gg_peer_push_rows_f32(ctypes.c_void_p(out_slice.data_ptr()), n, fs, per, world, c, o0, o1, arr,
                                                  ops._stream()), 'gg_peer_push_rows_f32')
        rows_out = []
        for o in range(world):
            rows = min(n, (o + 1) * per) - min(n, o * per)
            res = torch.empty((rows, f_total), device=cuda)
            ops.check(L.gg_peer_gather_slices_f32(ctypes.c_void_p(recv[o].data_ptr()), per, rows, fs, world,
                                                  ctypes.c_void_p(res.data_ptr()), f_total, ops._stream()),
                      'gg_peer_gather_slices_f32')
            rows_out.append(res)
        assert torch.equal(torch.cat(rows_out), got)


@pytest.mark.parametrize('f', [16, 32, 64, 128])
def test_narrow_sddmm_and_rank1_epilogue(cuda, monkeypatch, f):
    """The two extra pieces of the row-partitioned GAT: this rank's share of dalpha over a narrow column slice,
    and the rank-1 epilogue terms of the narrow-row aggregation."""
    n = 40000
    ei = powerlaw_graph(11, n, 14)
    g = torch.Generator().manual_seed(f)
    h = torch.randn(n, f, generator=g).to(cuda)
    gr = torch.randn(n, f, generator=g).to(cuda)
    csr = ops.layout_build(ei.to(cuda), n, 2, 0)
    got = ops.gat_sddmm_slice(csr, h, gr)
    rows, nbr = csr.rowid.long(), csr.nbr.long()
    want = (gr.double()[rows] * h.double()[nbr]).sum(1)
    assert rel_err(got, want) < FP32_TOL
    # odd tail: f/4 not a power of two
    if f == 32:
        got = ops.gat_sddmm_slice(csr, h[:, :24].contiguous(), gr[:, :24].contiguous())
        want = (gr.double()[rows, :24] * h.double()[nbr, :24]).sum(1)
        assert rel_err(got, want) < FP32_TOL
    w = torch.rand(csr.num_slots, generator=g).to(cuda)
    s1, s2 = torch.randn(n, generator=g).to(cuda), torch.randn(n, generator=g).to(cuda)
    v1, v2 = torch.randn(f, generator=g).to(cuda), torch.randn(f, generator=g).to(cuda)
    monkeypatch.setattr(ops, 'SPMM_ALGO', 'mpg')
    got = ops.spmm(csr, h, w, ops.SUM, rank1=(s1, v1, s2, v2))
    base = ops.spmm(csr, h, w, ops.SUM)
    want = base.double() + s1.double().view(-1, 1) * v1.double() + s2.double().view(-1, 1) * v2.double()
    assert rel_err(got, want) < FP32_TOL


@pytest.mark.parametrize('f', [32, 64, 104, 128, 256])
def test_bf16_gather_aggregation(cuda, f):
    """1e-2 mode: the gathered operand in bf16, fp32 products and sums.  Against an fp64 aggregation of the SAME
    bf16-rounded rows the kernel must be fp32-accurate; against the unrounded rows it is inside 1e-2."""
    n = 30000
    ei = powerlaw_graph(3, n, 16)
    g = torch.Generator().manual_seed(f)
    x = torch.randn(n, f, generator=g)
    bias = torch.randn(f, generator=g)
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    xb = ops.cast_bf16(x.to(cuda))
    assert xb.dtype == torch.bfloat16 and torch.equal(xb.cpu(), x.to(torch.bfloat16))      # round to nearest even
    a = torch.sparse_coo_tensor(torch.stack([csr.rowid.cpu().long(), csr.nbr.cpu().long()]), w.cpu().double(),
                                (n, n)).coalesce()
    for reduce, wts, self_scale, b in ((ops.SUM, w, 0.0, bias), (ops.SUM, None, 1.5, None), (ops.MEAN, None, 0.0, bias)):
        got = ops.spmm(csr, xb, wts, reduce, x.to(cuda) if self_scale else None, self_scale,
                       b.to(cuda) if b is not None else None)
        assert got.dtype == torch.float32
        if wts is not None:
            want_r = torch.sparse.mm(a, xb.cpu().double())
            want_x = torch.sparse.mm(a, x.double())
        else:
            # unweighted: every slot counts once (the generator emits duplicate edges, so not the coalesced pattern)
            ones = torch.sparse_coo_tensor(torch.stack([csr.rowid.cpu().long(), csr.nbr.cpu().long()]),
                                           torch.ones(csr.num_slots, dtype=torch.float64), (n, n)).coalesce()
            want_r, want_x = torch.sparse.mm(ones, xb.cpu().double()), torch.sparse.mm(ones, x.double())
            if reduce == ops.MEAN:
                deg = torch.diff(csr.rowptr.cpu()).clamp(min=1).double().view(-1, 1)
                want_r, want_x = want_r / deg, want_x / deg
        if self_scale:
            want_r, want_x = want_r + self_scale * x.double(), want_x + self_scale * x.double()
        if b is not None:
            want_r, want_x = want_r + b.double(), want_x + b.double()
        assert rel_err(got, want_r) < FP32_TOL, (f, reduce)
        assert rel_err(got, want_x) < 1e-2, (f, reduce)
    assert torch.equal(ops.spmm(csr, xb, w), ops.spmm(csr, xb, w))      # deterministic
