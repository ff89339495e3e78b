"""Next-step rows of SURVEY §8f: graph-level pooling (pooling.py) and the clustering-coefficient label
(feature_augment.py:81-82) on the GPU, against torch / networkx."""
import networkx as nx
import numpy as np
import pytest
import torch

from graphgym_b200.config import cfg
from graphgym_b200.contrib.transform.clustering import clustering_coefficient
from graphgym_b200.models import pooling
from util import FP32_TOL, rel_err

pytestmark = pytest.mark.gpu


def _ref_pool(x, batch, size, mode):
    out = torch.zeros(size, x.size(1), dtype=torch.float64)
    for g in range(size):
        rows = x[batch == g].double()
        if rows.numel():
            out[g] = {'add': rows.sum(0), 'mean': rows.mean(0), 'max': rows.max(0).values}[mode]
    return out


@pytest.mark.parametrize('mode', ['add', 'mean', 'max'])
@pytest.mark.parametrize('ego', [False, True])
def test_global_pooling_matches_scatter(cuda, mode, ego):
    g = torch.Generator().manual_seed(3)
    sizes = [5, 1, 0, 64, 300, 17]          # one empty graph
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    n, f = batch.numel(), 37
    x = torch.randn(n, f, generator=g)
    gy = torch.randn(len(sizes), f, generator=g)
    ids = torch.sort(torch.randperm(n, generator=g)[:n // 3]).values if ego else None
    cfg.dataset.transform = 'ego' if ego else 'none'
    try:
        xd = x.to(cuda).requires_grad_(True)
        out = pooling.pooling_dict[mode](xd, batch.to(cuda), ids.to(cuda) if ego else None, size=len(sizes))
        out.backward(gy.to(cuda))
    finally:
        cfg.dataset.transform = 'none'
    xs, bs = (x[ids], batch[ids]) if ego else (x, batch)
    xr = xs.double().requires_grad_(True)
    ref = torch.zeros(len(sizes), f, dtype=torch.float64)
    parts = []
    for k in range(len(sizes)):
        rows = xr[bs == k]
        parts.append({'add': rows.sum(0), 'mean': rows.mean(0), 'max': rows.max(0).values}[mode]
                     if rows.size(0) else torch.zeros(f, dtype=torch.float64))
    ref = torch.stack(parts)
    ref.backward(gy.double())
    assert rel_err(out.detach(), ref.detach()) < FP32_TOL
    want_gx = torch.zeros(n, f, dtype=torch.float64)
    if ego:
        want_gx[ids] = xr.grad
    else:
        want_gx = xr.grad
    assert rel_err(xd.grad, want_gx) < FP32_TOL
    assert torch.equal(out, pooling.pooling_dict[mode](xd.detach(), batch.to(cuda), None, size=len(sizes))) or ego


def test_pooling_rejects_unsorted_batch(cuda):
    x = torch.randn(6, 4, device=cuda)
    with pytest.raises(ValueError):
        pooling.global_add_pool(x, torch.tensor([0, 1, 0, 1, 2, 2], device=cuda), size=3)


def test_clustering_coefficient_matches_networkx(cuda):
    rng = np.random.default_rng(0)
    graphs = [nx.barabasi_albert_graph(64, 4, seed=1), nx.watts_strogatz_graph(64, 6, 0.2, seed=2),
              nx.complete_graph(5), nx.path_graph(7), nx.empty_graph(3)]
    off, edges, want, ptr = 0, [], [], [0]
    for gph in graphs:
        e = np.array(list(gph.edges()), dtype=np.int64).reshape(-1, 2) + off
        edges.append(np.concatenate([e, e[:, ::-1]]))
        want += list(nx.clustering(gph).values())
        off += gph.number_of_nodes()
        ptr.append(off)
    ei = torch.from_numpy(np.concatenate(edges).T.copy())
    ei = ei[:, torch.from_numpy(rng.permutation(ei.size(1)))]
    got = clustering_coefficient(ei.to(cuda), off)
    got_batched = clustering_coefficient(ei.to(cuda), off, graph_ptr=torch.tensor(ptr, dtype=torch.int32))
    want = torch.tensor(want, dtype=torch.float64)
    assert torch.allclose(got.cpu(), want, atol=1e-12)
    assert torch.allclose(got_batched.cpu(), want, atol=1e-12)


def test_pooling_matches_reference_golden(cuda, golden_dir):
    """against the outputs and gradients of the reference's own pooling.py (tests/golden/make_golden.py)"""
    import os
    z = np.load(os.path.join(golden_dir, 'pooling.npz'))
    x0, batch, ids, gy = (torch.from_numpy(z[k]).to(cuda) for k in ('x', 'batch', 'ids', 'gy'))
    size = int(z['size'])
    for transform in ('none', 'ego'):
        cfg.dataset.transform = transform
        try:
            for mode in ('add', 'mean', 'max'):
                x = x0.clone().requires_grad_(True)
                y = pooling.pooling_dict[mode](x, batch, ids, size=size)
                y.backward(gy)
                assert rel_err(y.detach(), torch.from_numpy(z['%s_%s/y' % (transform, mode)])) < FP32_TOL
                assert rel_err(x.grad, torch.from_numpy(z['%s_%s/gx' % (transform, mode)])) < FP32_TOL
        finally:
            cfg.dataset.transform = 'none'
