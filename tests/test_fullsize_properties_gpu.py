"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run there
in seconds): layout = a stable permutation, aggregation = a linear operator whose CSC call is its adjoint,
softmax rows sum to one, closed walks of length 2 = degree, ego-nets of radius 0 = the centres."""
import numpy as np
import pytest
import torch

import bench
from graphgym_b200 import ops
from graphgym_b200.graph import GraphLayout

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', params=['products_gcn', 'ba1m_sage'])
def big(request):
    dev = torch.device('cuda:0')
    spec = bench.WORKLOADS[request.param]
    n, ei = bench.gen_graph(spec, dev, seed=0)
    return spec, n, ei


def test_layout_is_a_stable_counting_sort(big):
    spec, n, ei = big
    for policy in (ops.LOOPS_KEEP, ops.LOOPS_ADD_REMAINING):
        for group_by in (ops.BY_TARGET, ops.BY_SOURCE):
            csr = ops.layout_build(ei, n, policy, group_by)
            E = ei.size(1)
            key_row = 1 if group_by == ops.BY_TARGET else 0
            perm, rowid, nbr = csr.perm.long(), csr.rowid.long(), csr.nbr.long()
            # rowptr is the histogram prefix of the kept group keys
            loops = (ei[0] == ei[1])
            kept = ~loops if policy == ops.LOOPS_ADD_REMAINING else torch.ones_like(loops)
            counts = torch.bincount(ei[key_row][kept], minlength=n) + (1 if policy == ops.LOOPS_ADD_REMAINING else 0)
            assert torch.equal(torch.diff(csr.rowptr.long()), counts)
            assert csr.num_slots == int(counts.sum())
            # sorted by key, and inside a key by edge id (= stability), every kept edge exactly once
            assert bool((rowid[1:] >= rowid[:-1]).all())
            same = rowid[1:] == rowid[:-1]
            assert bool((perm[1:][same] > perm[:-1][same]).all())
            orig = perm < E
            assert int(orig.sum()) == int(kept.sum())
            assert torch.equal(ei[key_row][perm[orig]], rowid[orig])
            assert torch.equal(ei[1 - key_row][perm[orig]], nbr[orig])
            assert torch.equal(perm[~orig] - E, rowid[~orig]) and torch.equal(nbr[~orig], rowid[~orig])


def test_aggregation_equals_fp64_scatter_at_full_size(big):
    """Direct comparison at BASELINE's full size: the aggregation kernels (sliced-ELL default, merge-path) against the
    reference's gather / scale / index_add_ evaluated in float64 on the device over slot chunks (a checker: library ops
    only) — norm-wise 1e-5 and element-wise."""
    spec, n, ei = big
    f = 128
    lay = GraphLayout(ei, n, ops.LOOPS_ADD_REMAINING)
    w_csr, _ = lay.weights('gcn_tgt')
    x = torch.randn(n, f, device=ei.device, generator=torch.Generator(device=ei.device).manual_seed(3))
    csr = lay.csr
    want = torch.zeros((n, f), dtype=torch.float64, device=ei.device)
    chunk = 1 << 21
    for s0 in range(0, csr.num_slots, chunk):
        sl = slice(s0, min(csr.num_slots, s0 + chunk))
        want.index_add_(0, csr.rowid[sl].long(), x.index_select(0, csr.nbr[sl].long()).double() * w_csr[sl].double().view(-1, 1))
    for algo in ('auto', 'mpg'):
        got = ops.spmm(csr, x, w_csr, algo=algo).double()
        assert float((got - want).abs().max() / want.abs().max()) < 1e-5, algo
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5 * float(want.abs().max())), algo
    del want


def test_aggregation_is_linear_and_csc_is_its_adjoint(big):
    spec, n, ei = big
    f = 128
    lay = GraphLayout(ei, n, ops.LOOPS_ADD_REMAINING)
    w_csr, w_csc = lay.weights('gcn_tgt')
    g = torch.Generator(device=ei.device).manual_seed(1)
    x = torch.randn(n, f, device=ei.device, generator=g)
    y = torch.randn(n, f, device=ei.device, generator=g)
    ax, ay = ops.spmm(lay.csr, x, w_csr), ops.spmm(lay.csr, y, w_csr)
    axy = ops.spmm(lay.csr, 2.0 * x - 3.0 * y, w_csr)
    scale = float(axy.abs().max())
    assert float((axy - (2.0 * ax - 3.0 * ay)).abs().max()) / scale < 1e-5                     # linearity
    aty = ops.spmm(lay.csc, y, w_csc)
    lhs, rhs = (ax.double() * y.double()).sum(), (x.double() * aty.double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-6                                      # <Ax,y> = <x,A^T y>
    ones = torch.ones(n, 4 * 32, device=ei.device)
    rowsum = ops.spmm(lay.csr, ones, w_csr)[:, 0]
    ref = torch.zeros(n, device=ei.device, dtype=torch.float64).index_add_(0, lay.csr.rowid.long(), w_csr.double())
    assert float((rowsum.double() - ref).abs().max()) / float(ref.abs().max()) < 1e-5          # checksum of rows
    mean1 = ops.spmm(lay.csr, ones, None, ops.MEAN)
    assert float((mean1 - 1.0).abs().max()) < 1e-6                                             # mean of a constant
    assert torch.equal(ax, ops.spmm(lay.csr, x, w_csr))                                        # bitwise repeatable


def test_edge_softmax_rows_sum_to_one(big):
    spec, n, ei = big
    lay = GraphLayout(ei, n, ops.LOOPS_REMOVE_ADD)
    g = torch.Generator(device=ei.device).manual_seed(2)
    h = torch.randn(n, 128, device=ei.device, generator=g)
    att = torch.randn(1, 1, 256, device=ei.device, generator=g) * 0.1
    import pytest
    mp = pytest.MonkeyPatch()
    try:
        mp.setattr(ops, 'GAT_ALGO', 'mp')                       # split passes: alpha is materialised
        out, alpha, _, _, _ = ops.gat_forward(lay.csr, h, att, 1, 0.2, None)
    finally:
        mp.undo()
    sums = torch.zeros(n, device=ei.device, dtype=torch.float64).index_add_(0, lay.csr.rowid.long(),
                                                                             alpha.view(-1).double())
    assert float((sums - 1.0).abs().max()) < 1e-5
    assert bool((alpha >= 0).all()) and bool(torch.isfinite(out).all())
    # a convex combination of rows stays inside the per-column range of h
    assert float(out.max()) <= float(h.max()) + 1e-4 and float(out.min()) >= float(h.min()) - 1e-4
    # fused path (online softmax inside the sliced-ELL aggregation, alpha never stored): the same output, and on h = 1
    # every row is sum_e alpha_e = 1
    fused, rowstat, a_tgt, a_src, _ = ops.gat_forward(lay.csr, h, att, 1, 0.2, None)
    assert rowstat.shape == (n, 2)
    assert float((fused - out).abs().max()) / float(out.abs().max()) < 1e-5
    ones = torch.ones(n, 128, device=ei.device)
    one_out, _, _ = ops.gat_sell_forward(lay.csr, ones, 128, a_tgt.view(-1), a_src.view(-1), 0.2, None)
    assert float((one_out - 1.0).abs().max()) < 1e-5
    # training forward: the same output bit for bit, plus the positive-logit share (a_pos in [0, 1], out_pos inside h's
    # range); the one-pass backward built on it agrees with the two-pass backward that moves dz through memory
    out_t, rs_t, _, _, pos = ops.gat_forward(lay.csr, h, att, 1, 0.2, None, need_grad=True)
    assert torch.equal(out_t, fused) and torch.equal(rs_t, rowstat) and pos is not None
    assert float(pos[1].min()) >= 0.0 and float(pos[1].max()) <= 1.0 + 1e-5
    gy = torch.randn(n, 128, device=ei.device, generator=g)
    dh1, datt1 = ops.gat_backward(lay.csr, lay.csc, lay.csc2csr, h, att, 1, 0.2, None, rowstat, a_tgt, a_src, fused, gy, pos)
    dh2, datt2 = ops.gat_backward(lay.csr, lay.csc, lay.csc2csr, h, att, 1, 0.2, None, rowstat, a_tgt, a_src, fused, gy)
    assert float((dh1 - dh2).norm() / dh2.norm()) < 1e-5
    assert float((datt1 - datt2).norm() / datt2.norm()) < 1e-4          # sums of 2.4 M cancelling terms


def test_cycle_counts_and_ego_invariants_on_a_large_batch():
    """4096 graphs of 64 nodes (the reference's dataset shape x16): closed walks of length 2 = degree,
    ego-nets of radius 0 = the centres alone, radius 5 = n copies of every graph."""
    from graphgym_b200.contrib.transform.identity import closed_walk_counts
    from graphgym_b200.models.transform import ego_nets_batch
    dev = torch.device('cuda:0')
    G, m = 1024, 64
    gen = torch.Generator(device=dev).manual_seed(0)
    a = torch.randint(0, m, (G, 160), device=dev, generator=gen)
    b = torch.randint(0, m, (G, 160), device=dev, generator=gen)
    off = (torch.arange(G, device=dev) * m).view(-1, 1)
    keep = a != b
    code = torch.unique(torch.minimum(a, b)[keep] * 0 + ((off + torch.minimum(a, b)) * (G * m) + off + torch.maximum(a, b))[keep])
    u, v = code // (G * m), code % (G * m)
    ei = torch.stack([torch.cat([u, v]), torch.cat([v, u])])
    n = G * m
    ptr = torch.arange(0, n + 1, m, device=dev)
    counts, overflow = closed_walk_counts(ei, n, 4, graph_ptr=ptr)
    deg = torch.bincount(ei[0], minlength=n)
    assert overflow == 0 and bool((counts[:, 0] == 0).all()) and torch.equal(counts[:, 1], deg)
    r0 = ego_nets_batch(ei, n, 0, ptr)
    assert r0['num_nodes'] == n and r0['edge_index'].size(1) == 0
    r5 = ego_nets_batch(ei, n, 5, ptr)
    assert r5['num_nodes'] == n * m and r5['edge_index'].size(1) == ei.size(1) * m
    r2 = ego_nets_batch(ei, n, 2, ptr)
    src, tgt = r2['edge_index']
    # induced edges are symmetric and stay inside one ego block; relabelled endpoints map back to real edges
    o = r2['orig_id']
    key = o[src] * n + o[tgt]
    assert bool(torch.isin(key, ei[0] * n + ei[1]).all())
