"""Round-2 starting point: time the experimental degree-binned aggregation (GG_SPMM_ALGO=bin, csrc/spmm_bin.cu) against
the default grouped-slot merge-path kernel on the products-shaped graph, all rows, F = 16 / 32 / 64 / 128.

    gpurun -- 'python scratch/bin_probe.py'
"""
import sys, torch
sys.path.insert(0, '/root/repo')
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


for thr in (256, 1024, 4096):
    ops.BIN_HUB_DEGREE = thr
    csr._binned = None
    for f in (16, 32, 64, 128):
        x = torch.randn(n, f, device=dev)
        a = timeit(lambda: ops.spmm(csr, x, w, algo='auto'))
        b = timeit(lambda: ops.spmm(csr, x, w, algo='bin'))
        err = float((ops.spmm(csr, x, w, algo='bin') - ops.spmm(csr, x, w, algo='auto')).abs().max())
        print(f'hub>{thr} f={f}: merge-path {a:.3f} ms, binned {b:.3f} ms (hubs={csr._binned.hubs}), max|diff| {err:.2e}', flush=True)
