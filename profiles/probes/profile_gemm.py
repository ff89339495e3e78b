"""A few launches of the layer-sized NN GEMM on the products shape (for ncu: -k regex:tc_gemm)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graphgym_b200 import ops
dev = torch.device('cuda')
n, k, f = 2449029, 100, 128
x = torch.randn(n, k, device=dev); w = torch.randn(k, f, device=dev); out = torch.empty(n, f, device=dev)
for _ in range(3):
    ops.id_gemm([(x, w, None)], n, f, out=out)
torch.cuda.synchronize()
