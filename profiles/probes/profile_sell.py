"""ncu target: the sliced-ELL aggregation at F = 16 and 32 on the products-shaped graph (what a rank of the 8- / 4-GPU
feature-sliced exchange runs), plus the fused post-ops at F = 128."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import bench
from graphgym_b200 import ops, functional as F_
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gcn'], dev)
csr = ops.layout_build(ei, n, 1, 0)
w = ops.gcn_norm(csr, ops.segment_degree(csr))
for f in (16, 32):
    x = torch.randn(n, f, device=dev)
    for _ in range(2):
        ops.spmm(csr, x, w, algo='sell')
torch.cuda.synchronize()
bn = torch.nn.BatchNorm1d(128).to(dev)
y = torch.randn(n, 128, device=dev, requires_grad=True)
for _ in range(2):
    F_.post_ops(y, bn, True, ops.ACT_RELU, 0.0, True).backward(torch.ones_like(y))
torch.cuda.synchronize()
