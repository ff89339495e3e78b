"""Times the sliced-ELL aggregation against the grouped-slot merge-path kernel on the products-shaped graph, all rows."""
import sys, time, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gcn'], dev)
csr = ops.layout_build(ei, n, 1, 0)
w = ops.gcn_norm(csr, ops.segment_degree(csr))


def timeit(fn, it=5):
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


for seg in (128, 256, 1024):
    ops.SELL_SEG = seg
    csr._sell = None
    torch.cuda.synchronize(); t0 = time.time()
    sl = ops.sell_layout(csr)
    torch.cuda.synchronize()
    print(f'seg={seg}: build {1e3 * (time.time() - t0):.1f} ms, vrows {sl.vrows}, chunks {sl.chunks}, padded slots '
          f'{4 * sl.units} ({4 * sl.units / csr.num_slots:.3f}x), split rows {sl.hubs}, partial rows {sl.partial_rows}', flush=True)
    for f in (16, 32, 64, 128):
        x = torch.randn(n, f, device=dev)
        a = timeit(lambda: ops.spmm(csr, x, w, algo='mpg'))
        b = timeit(lambda: ops.spmm(csr, x, w, algo='sell'))
        c = timeit(lambda: ops.spmm(csr, x, None, algo='sell'))
        ref = ops.spmm(csr, x, w, algo='mpg')
        err = float((ops.spmm(csr, x, w, algo='sell') - ref).abs().max() / ref.abs().max())
        print(f'  f={f}: merge-path {a:.3f} ms, sell {b:.3f} ms ({csr.num_slots / b / 1e6:.1f} G slots/s), unweighted {c:.3f} ms, '
              f'rel diff {err:.1e}', flush=True)
