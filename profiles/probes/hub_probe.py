"""Hub hint of the sliced-ELL aggregation: time and parity on the products-shaped graph for several L2 budgets."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gcn'], dev)
csr = ops.layout_build(ei, n, 1, 0)
w = ops.gcn_norm(csr, ops.segment_degree(csr))


def timeit(fn, it=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


for f in (128, 64, 16):
    x = torch.randn(n, f, device=dev)
    ops.SELL_HUB_BYTES, ops.SELL_HUB_MIN_F = 0, 0
    ref = ops.spmm(csr, x, w, algo='sell')
    line = [f'f={f}: off {timeit(lambda: ops.spmm(csr, x, w, algo="sell")):.3f} ms']
    for mb in (16, 32, 48, 64, 96):
        ops.SELL_HUB_BYTES = mb << 20
        got = ops.spmm(csr, x, w, algo='sell')
        assert torch.equal(got, ref), 'the hint must not change the result'
        line.append(f'{mb} MB {timeit(lambda: ops.spmm(csr, x, w, algo="sell")):.3f}')
    print(', '.join(line), flush=True)
