"""ncu target: one launch each of the narrow-row aggregation (F/P = 16 and 32 columns of the products graph =
the per-rank work of the feature-sliced exchange at 8 and 4 GPUs) and of the narrow SDDMM."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gcn'], dev)
csr = ops.layout_build(ei, n, 1, 0)
w = ops.gcn_norm(csr, ops.segment_degree(csr))
ops.SPMM_ALGO = 'mpg'
for fs in (16, 32):
    x = torch.randn(n, fs, device=dev)
    for _ in range(3):
        ops.spmm(csr, x, w)
    g = torch.randn(n, fs, device=dev)
    for _ in range(2):
        ops.gat_sddmm_slice(csr, x, g)
torch.cuda.synchronize()
print('ok')
