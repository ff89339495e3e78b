"""A few launches of the layer-sized weight-gradient GEMM on the products shape (for ncu: -k regex:tc_gemm_tn)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graphgym_b200 import ops
dev = torch.device('cuda')
n, k, f = 2449029, 100, 128
x = torch.randn(n, k, device=dev); g = torch.randn(n, f, device=dev)
for _ in range(3):
    ops.gemm_tn(x, g)
torch.cuda.synchronize()
