"""Row-partitioned layer on 2 ranks over gloo (CPU): partition bounds, padded all-gather of the halo,
degree exchange, backward all-gather and gradient all-reduce must reproduce the single-process oracle.
The kernels are replaced by oracle-backed stand-ins (tests/fake_ops.py) — this tier has no GPU; the same
code path runs on NCCL in tests/test_parallel_gpu.py and bench.py --gpus N."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, exchange, ret, fout=8, name='gcnconv'):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import fake_ops
        from graphgym_b200 import ops, parallel
        from oracle import layers as olayers
        from util import random_graph
        fake_ops.install(lambda mod, name, fn: setattr(mod, name, fn))
        torch.manual_seed(0)
        fin = 12
        ei = random_graph(3, n, 6 * n, loops=5, dups=7)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n, fin, generator=g)
        gy = torch.randn(n, fout, generator=g)
        part = parallel.RowPartition(n, world, rank)
        cls, policy, _ = parallel.ROW_PARTITIONED[name]
        if name not in ('gcnconv', 'gcnidconv', 'gatconv'):
            fin = fout if exchange.startswith('sliced') else fin   # SAGE / GIN aggregate the INPUT features
            x = torch.randn(n, fin, generator=g)
        layer = cls(fin, fout, bias=True)   # same seed on every rank
        with torch.no_grad():
            for prm in layer.parameters():
                if prm.dim() == 1:
                    prm.uniform_(-0.5, 0.5)
        playout = parallel.PartitionedLayout(ei, n, policy, part, exchange=exchange)
        xl = x[part.lo:part.hi].clone().requires_grad_(True)
        ids = torch.arange(0, n, 3)                        # ID-GNN centres on both ranks
        y = layer(xl, playout, ids) if name.endswith('idconv') else layer(xl, playout)
        y.backward(gy[part.lo:part.hi])
        parallel.allreduce_grads(layer)
        # single-process oracle
        P = {k: v.detach().clone().requires_grad_(True) for k, v in layer.model.named_parameters()}
        xo = x.clone().requires_grad_(True)
        if name == 'gcnconv':
            yo = olayers.gcnconv(xo, ei, P['weight'], P['bias'])
        elif name == 'gcnidconv':
            yo = olayers.gcn_idconv(xo, ei, ids, P['weight'], P['weight_id'], P['bias'])
        elif name == 'sageconv':
            yo = olayers.sageconv(xo, ei, P['lin_l.weight'], P['lin_l.bias'], P['lin_r.weight'])
        elif name == 'gatconv':
            yo = olayers.gatconv(xo, ei, P['weight'], P['att'], P['bias'])
        elif name == 'sageidconv':
            yo = olayers.sage_idconv(xo, ei, ids, P['weight'], P['weight_id'], P['bias'], concat=True)
        elif name == 'ginidconv':
            pn = (P['nn.0.weight'], P['nn.0.bias'], P['nn.2.weight'], P['nn.2.bias'])
            pi = (P['nn_id.0.weight'], P['nn_id.0.bias'], P['nn_id.2.weight'], P['nn_id.2.bias'])
            yo = olayers.gin_idconv(xo, ei, ids, pn, pi)
        else:
            yo = olayers.ginconv(xo, ei, P['nn.0.weight'], P['nn.0.bias'], P['nn.2.weight'], P['nn.2.bias'])
        yo.backward(gy)
        ok = torch.allclose(y.detach(), yo.detach()[part.lo:part.hi], atol=1e-5)
        ok &= torch.allclose(xl.grad, xo.grad[part.lo:part.hi], atol=1e-5)
        for k, v in layer.model.named_parameters():
            ok &= torch.allclose(v.grad, P[k].grad, atol=1e-4)
        ret[rank] = bool(ok) and part.rows == (part.hi - part.lo)
    finally:
        dist.destroy_process_group()


# one all-gather | per-peer send/recv rounds | feature-sliced transposition (all-to-all form of the peer-memory path)
@pytest.mark.parametrize('exchange', ['allgather', 'pipelined', 'sliced_nccl'])
@pytest.mark.parametrize('n', [40, 41])   # 41: the last rank's block is shorter -> padded exchange
def test_row_partitioned_gcn_world2(n, exchange):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n, exchange, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_partition_bounds_cover_all_rows():
    sys.path.insert(0, os.path.dirname(HERE))
    from graphgym_b200.parallel import RowPartition
    for n in (0, 1, 7, 8, 9, 2449029):
        for world in (1, 2, 4, 8):
            spans = [RowPartition(n, world, r) for r in range(world)]
            assert spans[0].lo == 0 and spans[-1].hi == n
            assert all(a.hi == b.lo for a, b in zip(spans, spans[1:]))
            assert all(s.rows <= s.per for s in spans)


def test_row_partitioned_gcn_world3_pipelined():
    """three ranks: two exchange rounds with different peers per round"""
    world = 3
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), 50, 'pipelined', ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_row_partitioned_gcn_world3_sliced():
    """three ranks, 12 output columns -> slices of 4; 50 rows -> the last block is short"""
    world = 3
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), 50, 'sliced_nccl', ret, 12), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


@pytest.mark.parametrize('exchange', ['allgather', 'sliced_nccl'])
@pytest.mark.parametrize('name', ['sageconv', 'ginconv', 'gcnidconv', 'sageidconv', 'ginidconv'])
def test_row_partitioned_sage_gin_world2(name, exchange):
    """mean aggregation with global degrees (SAGE), self term through the exchange (GIN)"""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), 41, exchange, ret, 8, name), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


@pytest.mark.parametrize('world,fout', [(2, 8), (3, 12)])
def test_row_partitioned_gat_sliced(world, fout):
    """edge-softmax layer on the partition: logits all-gathered, sliced aggregations with rank-1 terms, sliced SDDMM
    shares all-reduced — host logic over gloo with per-slot stand-ins for the GAT passes"""
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), 41, 'sliced_nccl', ret, fout, 'gatconv'), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_sliced_width():
    sys.path.insert(0, os.path.dirname(HERE))
    from graphgym_b200.parallel import sliced_width
    assert sliced_width(128, 8) == 16 and sliced_width(128, 2) == 64 and sliced_width(256, 2) == 128
    assert sliced_width(100, 2) == 0      # 50-column slices are not 16-byte aligned
    assert sliced_width(512, 2) == 0      # 256-column slices exceed the sub-warp-group kernel
    assert sliced_width(128, 1) == 128
