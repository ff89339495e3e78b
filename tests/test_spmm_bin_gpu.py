"""EXPERIMENTAL degree-binned aggregation (csrc/spmm_bin.cu, GG_SPMM_ALGO=bin): written at the end of round 1; the parity
cases below passed on a B200 once, the empty-layout case was fixed after the GPU budget ran out, nothing is timed yet —
so these tests stay opt-in (GG_TEST_EXPERIMENTAL=1) until the kernel has been measured."""
import os

import pytest
import torch

from graphgym_b200 import ops
from util import FP32_TOL, powerlaw_graph, random_graph, rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('GG_TEST_EXPERIMENTAL') != '1',
                                 reason='experimental kernel: set GG_TEST_EXPERIMENTAL=1')]


@pytest.mark.parametrize('f', [4, 16, 24, 32, 64, 128])
def test_binned_matches_merge_path(cuda, monkeypatch, f):
    n = 50000
    ei = powerlaw_graph(13, n, 18)                      # hubs above BIN_HUB_DEGREE and isolated rows
    g = torch.Generator().manual_seed(f)
    x = torch.randn(n, f, generator=g).to(cuda)
    bias = torch.randn(f, generator=g).to(cuda)
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    assert int(torch.diff(csr.rowptr).max()) > ops.BIN_HUB_DEGREE
    for wts, reduce, self_scale, b in ((w, ops.SUM, 0.0, bias), (None, ops.MEAN, 0.0, None), (None, ops.SUM, 1.25, bias)):
        monkeypatch.setattr(ops, 'SPMM_ALGO', 'auto')
        want = ops.spmm(csr, x, wts, reduce, x if self_scale else None, self_scale, b)
        monkeypatch.setattr(ops, 'SPMM_ALGO', 'bin')
        got = ops.spmm(csr, x, wts, reduce, x if self_scale else None, self_scale, b)
        assert rel_err(got, want) < FP32_TOL, (f, reduce)
        assert torch.equal(got, ops.spmm(csr, x, wts, reduce, x if self_scale else None, self_scale, b))


def test_binned_small_and_empty(cuda, monkeypatch):
    for n, e in ((1, 0), (5, 0), (40, 10), (997, 9000)):
        ei = random_graph(n, n, e, loops=2 if e else 0) if e else torch.zeros((2, 0), dtype=torch.int64)
        x = torch.randn(n, 32, generator=torch.Generator().manual_seed(n)).to(cuda)
        csr = ops.layout_build(ei.to(cuda), n, 0, 0)
        monkeypatch.setattr(ops, 'SPMM_ALGO', 'row')
        want = ops.spmm(csr, x, None, ops.SUM, x, 1.0)
        monkeypatch.setattr(ops, 'SPMM_ALGO', 'bin')
        assert rel_err(ops.spmm(csr, x, None, ops.SUM, x, 1.0), want) < FP32_TOL
