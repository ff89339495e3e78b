"""Does a degree-descending relabelling of the SOURCE rows speed up narrow-slice gathers (hot rows share L2 lines / DRAM pages)?"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
spec = bench.WORKLOADS['products_gcn']
n, ei = bench.gen_graph(spec, dev)
def timeit(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
deg = torch.bincount(ei[0], minlength=n)
order = torch.argsort(deg, descending=True, stable=True)
rank = torch.empty_like(order); rank[order] = torch.arange(n, device=dev)
for name, e2 in (('natural', ei), ('sources by degree', torch.stack([rank[ei[0]], ei[1]])), ('both by degree', rank[ei])):
    csr = ops.layout_build(e2.contiguous(), n, 1, 0)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    for fs in (16, 32, 64, 128):
        x = torch.randn(n, fs, device=dev)
        ops.SPMM_ALGO = 'mpg' if fs < 128 else 'mp'
        ms = timeit(lambda: ops.spmm(csr, x, w))
        print(f'{name:18s} fs={fs}: {ms:.3f} ms  {csr.num_slots/ms/1e6:.1f} G slots/s', flush=True)
    del csr, w
