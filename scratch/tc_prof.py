import sys, torch
sys.path.insert(0, '/root/repo')
from graphgym_b200 import ops
dev = torch.device('cuda')
n, k, f = 2449029, 100, 128
x, w = torch.randn(n, k, device=dev), torch.randn(k, f, device=dev)
ops.GEMM_MODE = 'tc'
for _ in range(3):
    ops.id_gemm([(x, w, None)], n, f)
torch.cuda.synchronize()
print('ok')
