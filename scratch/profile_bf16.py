import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gcn'], dev)
csr = ops.layout_build(ei, n, 1, 0)
w = ops.gcn_norm(csr, ops.segment_degree(csr))
x = torch.randn(n, 128, device=dev)
xb = ops.cast_bf16(x)
for _ in range(3):
    ops.spmm(csr, xb, w)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ops.spmm(csr, xb, w)
e.record(); torch.cuda.synchronize()
print('bf16 spmm F=128', s.elapsed_time(e) / 5, 'ms')
