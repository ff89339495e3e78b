"""CSR aggregation kernel vs the oracle's gather/scale/scatter restatement (fp32, 1e-5 rel)."""
import numpy as np
import pytest
import torch

from graphgym_b200 import ops
from oracle import layers as olayers
from oracle import pyg_utils as U
from util import FP32_TOL, powerlaw_graph, random_graph, rel_err

pytestmark = pytest.mark.gpu


def oracle_spmm(ei, x, w_edge, reduce, self_scale, bias):
    xj = x.double().index_select(0, ei[0])
    if w_edge is not None:
        xj = xj * w_edge.double().view(-1, 1)
    out = U.propagate(ei, xj, x.size(0), 'mean' if reduce == ops.MEAN else 'add')
    if self_scale:
        out = out + self_scale * x.double()
    if bias is not None:
        out = out + bias.double()
    return out


@pytest.mark.parametrize('f', [1, 3, 4, 8, 16, 20, 32, 64, 100, 128, 136, 256, 300, 512, 1024, 1433])
def test_feature_widths(cuda, f):
    n = 211
    ei = random_graph(f, n, 1500, loops=10, dups=20)
    g = torch.Generator().manual_seed(f)
    x = torch.randn(n, f, generator=g)
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    w_edge = torch.rand(ei.size(1), generator=g)
    w_slot = w_edge[csr.perm.cpu().long()].to(cuda)
    bias = torch.randn(f, generator=g)
    for weighted in (False, True):
        for reduce in (ops.SUM, ops.MEAN):
            for self_scale, b in ((0.0, None), (1.5, bias)):
                got = ops.spmm(csr, x.to(cuda), w_slot if weighted else None, reduce,
                               x.to(cuda) if self_scale else None, self_scale,
                               b.to(cuda) if b is not None else None)
                want = oracle_spmm(ei, x, w_edge if weighted else None, reduce, self_scale, b)
                assert rel_err(got, want) < FP32_TOL, (f, weighted, reduce, self_scale)


def test_empty_rows_and_empty_graph(cuda):
    n, f = 64, 128
    x = torch.randn(n, f)
    ei = torch.zeros((2, 0), dtype=torch.int64)
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    for reduce in (ops.SUM, ops.MEAN):
        assert torch.count_nonzero(ops.spmm(csr, x.to(cuda), None, reduce)) == 0


@pytest.mark.parametrize('f', [64, 128, 256])
def test_powerlaw_hubs(cuda, f):
    """Hub rows with thousands of slots next to degree-1 rows (the load-balance case)."""
    n = 20000
    ei = powerlaw_graph(3, n, 16)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(0))
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    deg = ops.segment_degree(csr)
    w = ops.gcn_norm(csr, deg)
    got = ops.spmm(csr, x.to(cuda), w)
    ei2, norm = olayers.gcn_norm_tgt(ei, n, torch.float64)
    want = oracle_spmm(ei2, x, norm, ops.SUM, 0.0, None)
    assert rel_err(got, want) < FP32_TOL
    assert int(torch.diff(csr.rowptr).max()) > 500


def test_transpose_identity(cuda):
    """<A x, y> == <x, A^T y>: the CSC call really is the adjoint of the CSR call."""
    n, f = 3000, 128
    ei = random_graph(9, n, 40000, loops=50)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(n, f, generator=g).to(cuda), torch.randn(n, f, generator=g).to(cuda)
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    csc = ops.layout_build(ei.to(cuda), n, 1, 1)
    deg = ops.segment_degree(csr)
    ax = ops.spmm(csr, x, ops.gcn_norm(csr, deg))
    aty = ops.spmm(csc, y, ops.gcn_norm(csc, deg))
    lhs, rhs = (ax.double() * y.double()).sum(), (x.double() * aty.double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-6


def test_strided_output_and_determinism(cuda):
    n, f = 1000, 64
    ei = random_graph(4, n, 9000)
    x = torch.randn(n, f).to(cuda)
    csr = ops.layout_build(ei.to(cuda), n, 0, 0)
    wide = torch.zeros(n, 2 * f, device=cuda)
    ops.spmm(csr, x, None, ops.MEAN, out=wide[:, f:])
    ref = ops.spmm(csr, x, None, ops.MEAN)
    assert torch.equal(wide[:, f:], ref) and torch.count_nonzero(wide[:, :f]) == 0
    assert torch.equal(ref, ops.spmm(csr, x, None, ops.MEAN))  # bitwise run-to-run


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.layout_build(torch.zeros((2, 3), dtype=torch.int64), 4)
