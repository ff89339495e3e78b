"""One launch of every fused-GAT pass on the products-shaped graph (for ncu: -k regex:gat_sell)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from graphgym_b200 import ops
from graphgym_b200.graph import GraphLayout
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gat'], dev)
lay = GraphLayout(ei, n, ops.LOOPS_REMOVE_ADD)
f = 128
h = torch.randn(n, f, device=dev); g = torch.randn(n, f, device=dev)
att = torch.randn(1, 1, 2 * f, device=dev) * 0.1
bias = torch.zeros(f, device=dev)
for _ in range(2):
    out, al, a_tgt, a_src, pos = ops.gat_forward(lay.csr, h, att, 1, 0.2, bias, True)
    ops.gat_backward(lay.csr, lay.csc, lay.csc2csr, h, att, 1, 0.2, bias, al, a_tgt, a_src, out, g, pos)
torch.cuda.synchronize()
