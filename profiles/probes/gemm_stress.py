"""Determinism stress of the persistent GEMMs: many launches on fixed inputs must give the same bits, interleaved with
other work on the device (the per-tile kernel, an aggregation) so that launches overlap at kernel boundaries."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graphgym_b200 import ops
dev = torch.device('cuda')
torch.manual_seed(0)
bad = 0
for n, k, f in ((30001, 128, 128), (30001, 100, 128), (148 * 256 + 5, 100, 128), (2449029, 100, 128)):
    x = torch.randn(n, k, device=dev); w = torch.randn(k, f, device=dev); b = torch.randn(f, device=dev)
    m = torch.randn(n, f, device=dev); g = torch.randn(n, f, device=dev)
    small = torch.randn(3000, k, device=dev)
    ref1 = ops.id_gemm([(x, w, None)], n, f, bias=b, act=ops.ACT_RELU).clone()
    ref2 = ops.id_gemm([(g, w, None)], n, k, b_trans=True, relu_mask=x).clone()
    ref3 = ops.gemm_tn(x, g).clone()
    want = (x[:2048].double() @ w.double() + b.double()).relu()
    print(n, k, f, 'err', float((ref1[:2048] - want).norm() / want.norm()), flush=True)
    reps = 40 if n > 1000000 else 300
    for it in range(reps):
        ops.id_gemm([(small, w, None)], 3000, f)          # per-tile kernel in between
        o1 = ops.id_gemm([(x, w, None)], n, f, bias=b, act=ops.ACT_RELU)
        o2 = ops.id_gemm([(g, w, None)], n, k, b_trans=True, relu_mask=x)
        o3 = ops.gemm_tn(x, g)
        if not (torch.equal(o1, ref1) and torch.equal(o2, ref2) and torch.equal(o3, ref3)):
            bad += 1
            print('MISMATCH', n, k, f, it, torch.equal(o1, ref1), torch.equal(o2, ref2), torch.equal(o3, ref3),
                  int((o1 != ref1).sum()), int((o2 != ref2).sum()), int((o3 != ref3).sum()), flush=True)
print('mismatching launches:', bad)
