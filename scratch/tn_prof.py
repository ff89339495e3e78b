import sys, torch
sys.path.insert(0, '/root/repo')
from graphgym_b200 import ops
dev = torch.device('cuda')
n, k, f = 2449029, 100, 128
a, g = torch.randn(n, k, device=dev), torch.randn(n, f, device=dev)
ops.GEMM_MODE = 'tc'
for _ in range(2):
    ops.gemm_tn(a, g)
torch.cuda.synchronize()
print('ok')
