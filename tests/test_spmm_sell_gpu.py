"""Degree-sorted sliced-ELL aggregation (csrc/spmm_sell.cu): the integer build against its numpy restatement (bit-exact),
the kernel against the fp64 oracle of the reference's gather / scale / scatter (ref: idconv.py:89-92,177-180) on the shapes
that stress the layout — rows split into virtual rows, empty rows, fewer rows than a chunk, every group width."""
import numpy as np
import pytest
import torch

from graphgym_b200 import ops
from oracle import pyg_utils as U
from oracle import sell as osell
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


def check_build(csr, seg):
    sl = ops.SellLayout(csr, seg)
    want = osell.build(csr.rowptr.cpu().numpy(), csr.nbr.cpu().numpy(), seg)
    for k in ('vrows', 'units', 'hubs', 'partial_rows'):
        assert getattr(sl, k) == want[k], k
    ch = want['chunks']
    assert sl.chunks >= ch                                   # capacity chunks past the last virtual row are empty
    cp = sl.chunk_ptr.cpu().numpy().astype(np.int64)
    assert (cp[:ch + 1] == want['chunk_ptr']).all() and (cp[ch:sl.chunks + 1] == want['units']).all()
    assert sl.total == 4 * want['units']
    assert (sl.idx.cpu().numpy()[:sl.total] == want['idx']).all()
    assert (sl.slot_of.cpu().numpy()[:sl.total] == want['slot_of']).all()
    vd = sl.vdst.cpu().numpy().astype(np.int64)
    assert (vd[:ch * 8] == want['vdst']).all() and (vd[ch * 8:sl.chunks * 8] == osell.NO_ROW).all()
    assert (sl.hub_rows.cpu().numpy()[:sl.hubs] == want['hub_rows']).all()
    assert (sl.hub_pptr.cpu().numpy()[:sl.hubs + 1] == want['hub_pptr']).all()
    return sl


@pytest.mark.parametrize('seg', [4, 8, 64, 256])
@pytest.mark.parametrize('n,e', [(1, 0), (5, 3), (8, 40), (9, 200), (300, 2500), (1000, 300)])
def test_build_bit_exact(cuda, n, e, seg):
    ei = random_graph(n * 7 + e, n, e, loops=min(n, 3), dups=min(e, 5)) if e else torch.zeros((2, 0), dtype=torch.int64)
    for policy in (ops.LOOPS_KEEP, ops.LOOPS_ADD_REMAINING):
        csr = ops.layout_build(ei.to(cuda), n, policy, ops.BY_TARGET)
        check_build(csr, seg)


def test_build_bit_exact_powerlaw_hubs(cuda):
    ei = powerlaw_graph(3, 3000, 12)
    csr = ops.layout_build(ei.to(cuda), 3000, ops.LOOPS_ADD_REMAINING, ops.BY_TARGET)
    sl = check_build(csr, 32)
    assert sl.hubs > 0 and sl.partial_rows > 2 * sl.hubs - 1


@pytest.mark.parametrize('f', [4, 12, 16, 24, 32, 60, 64, 100, 128])
def test_parity_random(cuda, monkeypatch, f):
    n = 997
    ei = random_graph(f, n, 9000, loops=30, dups=40)
    g = torch.Generator().manual_seed(f)
    x = torch.randn(n, f, generator=g)
    w_edge = torch.rand(ei.size(1), generator=g)
    bias = torch.randn(f, generator=g)
    monkeypatch.setattr(ops, 'SELL_SEG', 8)                  # most rows are split: exercises partials + fix-up
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    w_slot = w_edge[csr.perm.cpu().long()].to(cuda)
    for weighted in (False, True):
        for reduce in (ops.SUM, ops.MEAN):
            for self_scale, b in ((0.0, None), (1.25, bias)):
                got = ops.spmm(csr, x.to(cuda), w_slot if weighted else None, reduce, x.to(cuda) if self_scale else None,
                               self_scale, b.to(cuda) if b is not None else None, algo='sell')
                want = oracle(ei, x, w_edge if weighted else None, reduce, self_scale, b)
                assert rel_err(got, want) < FP32_TOL, (f, weighted, reduce, self_scale)


@pytest.mark.parametrize('f', [16, 32, 128])
def test_powerlaw_matches_oracle_and_is_deterministic(cuda, f):
    n = 50000
    ei = powerlaw_graph(5, n, 20)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    got = ops.spmm(csr, x.to(cuda), w, algo='sell')
    again = ops.spmm(csr, x.to(cuda), w, algo='sell')
    assert torch.equal(got, again)
    ei2, w2 = U.add_remaining_self_loops(ei, torch.ones(ei.size(1), dtype=torch.float64), 1, n)
    deg = U.scatter_add(w2, ei2[1], 0, n)
    norm = deg.pow(-0.5)[ei2[0]] * w2 * deg.pow(-0.5)[ei2[1]]
    want = U.propagate(ei2, x.double().index_select(0, ei2[0]) * norm.view(-1, 1), n, 'add')
    assert rel_err(got, want) < FP32_TOL
    # element-wise as well (atol + rtol |ref|)
    assert torch.allclose(got.double().cpu(), want, rtol=1e-5, atol=1e-6)


def test_rank1_terms_nonfinite_row0_and_empty(cuda):
    n, f = 500, 16
    ei = random_graph(11, n, 3000)
    ei = ei[:, (ei[0] != 0)]                                 # node 0 is nobody's source
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, f, generator=g)
    x[0] = float('inf')                                      # padding must not gather it
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    s1, s2 = torch.randn(n, generator=g), torch.randn(n, generator=g)
    v1, v2 = torch.randn(f, generator=g), torch.randn(f, generator=g)
    got = ops.spmm(csr, x.to(cuda), None, ops.SUM, rank1=(s1.to(cuda), v1.to(cuda), s2.to(cuda), v2.to(cuda)), algo='sell')
    want = oracle(ei, x, None, ops.SUM, 0.0, None) + s1.double().view(-1, 1) * v1.double() + s2.double().view(-1, 1) * v2.double()
    assert torch.isfinite(got).all() and rel_err(got, want) < FP32_TOL
    empty = ops.layout_build(torch.zeros((2, 0), dtype=torch.int64, device=cuda), 7, 0, 0)
    out = ops.spmm(empty, torch.randn(7, 8, device=cuda), None, ops.SUM, bias=torch.ones(8, device=cuda), algo='sell')
    assert torch.equal(out, torch.ones(7, 8, device=cuda))


def test_hub_hint_marks_the_most_referenced_sources_and_changes_nothing(cuda, monkeypatch):
    """gg_sell_hub_hint: bit 30 set exactly on the entries whose source is referenced at least T times, T the smallest
    count with #{count >= T} <= hubs; the aggregation with the hint is bitwise the one without."""
    n, f = 30000, 64
    ei = powerlaw_graph(5, n, 16)
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    sl = ops.sell_layout(csr)
    hubs = 1024
    hint = sl.idx_hint(hubs).cpu().numpy()[:sl.total]
    idx = sl.idx.cpu().numpy()[:sl.total]
    freq = np.bincount(csr.nbr.cpu().numpy(), minlength=n)
    counts = np.bincount(np.minimum(freq, 4095), minlength=4096)
    above, t = 0, 4096
    for b in range(4095, 0, -1):
        if above + counts[b] > hubs:
            break
        above += counts[b]
        t = b
    valid = idx >= 0
    assert (hint[~valid] == -1).all()
    assert ((hint[valid] & ~(1 << 30)) == idx[valid]).all()
    assert (((hint[valid] >> 30) & 1) == (freq[idx[valid]] >= t)).all()
    assert 0 < ((hint[valid] >> 30) & 1).sum() < valid.sum()
    x = torch.randn(n, f, device=cuda, generator=torch.Generator(device=cuda).manual_seed(0))
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    monkeypatch.setattr(ops, 'SELL_HUB_BYTES', 0)
    ref = ops.spmm(csr, x, w, algo='sell')
    monkeypatch.setattr(ops, 'SELL_HUB_BYTES', 1 << 18)       # 256 KB: far below the 7.7 MB matrix -> the hint engages
    assert torch.equal(ops.spmm(csr, x, w, algo='sell'), ref)
