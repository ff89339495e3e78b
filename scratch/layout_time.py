import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
for wl in ('products_gcn', 'ba1m_sage'):
    n, ei = bench.gen_graph(bench.WORKLOADS[wl], dev)
    for _ in range(2): ops.layout_build(ei, n, 1, 0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): ops.layout_build(ei, n, 1, 0)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    E = ei.size(1)
    print(wl, 'layout_build (one grouping):', round(ms, 3), 'ms', round((E * 16 + (E + n) * 12) / ms / 1e6, 1), 'GB/s algorithmic')
