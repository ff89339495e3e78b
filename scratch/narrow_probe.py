"""Why is the fs=16 / fs=32 slice aggregation stuck at ~33 G slots/s?  (a) same edge count on a graph whose
slice matrix fits L2, (b) the products graph for ncu."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
def timeit(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
ops.SPMM_ALGO = 'mpg'
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for nn in ((2_449_029,) if which == 'ncu' else (16_000, 32_000, 64_000, 128_000, 300_000, 2_449_029)):
    spec = dict(bench.WORKLOADS['products_gcn'], n=nn, max_deg=10**9)
    n, ei = bench.gen_graph(spec, dev)
    csr = ops.layout_build(ei, n, 1, 0)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    for fs in ((16,) if which == 'ncu' else (16, 32)):
        x = torch.randn(n, fs, device=dev)
        ms = timeit(lambda: ops.spmm(csr, x, w), it=1 if which == 'ncu' else 5)
        print(f'n={n} slots={csr.num_slots} fs={fs} matrix={n*fs*4/1e6:.0f} MB: {ms:.3f} ms = {csr.num_slots/ms/1e6:.1f} G slots/s', flush=True)
    del csr, w, ei
