"""Layer-sized GEMMs on the products shape: persistent tcgen05 kernel vs the per-tile kernel (GG_GEMM_PERSIST=0) vs cuBLAS."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graphgym_b200 import ops
dev = torch.device('cuda')
n = 2449029


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


for k, f, bt in ((100, 128, False), (128, 100, True), (128, 128, False), (64, 64, False)):
    x = torch.randn(n, k, device=dev)
    w = torch.randn((f, k) if bt else (k, f), device=dev)
    out = torch.empty(n, f, device=dev)
    t = timeit(lambda: ops.id_gemm([(x, w, None)], n, f, b_trans=bt, out=out))
    gb = (n * k + n * f) * 4 / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    wt = w.t() if bt else w
    tc = timeit(lambda: torch.matmul(x, wt, out=out))
    ref = (x[:4096].double() @ wt.double())
    got = ops.id_gemm([(x, w, None)], n, f, b_trans=bt)[:4096]
    err = float((got - ref).norm() / ref.norm())
    print(f'K={k} F={f} b_trans={bt}: ours {t:.3f} ms = {gb / t:.2f} TB/s ({gb / t / 6.5443:.2f} of copy peak); '
          f'cuBLAS fp32 {tc:.3f} ms; rel err {err:.2e}', flush=True)

for k, f in ((100, 128), (128, 128), (64, 64)):
    x = torch.randn(n, k, device=dev)
    gr = torch.randn(n, f, device=dev)
    t = timeit(lambda: ops.gemm_tn(x, gr))
    gb = (n * k + n * f) * 4 / 1e9
    tc = timeit(lambda: torch.matmul(x.t(), gr))
    ref = x[:200000].double().t() @ gr[:200000].double()
    got = ops.gemm_tn(x[:200000], gr[:200000])
    print(f'TN K={k} F={f}: ours {t:.3f} ms = {gb / t:.2f} TB/s ({gb / t / 6.5443:.2f} of copy peak); cuBLAS fp32 {tc:.3f} ms; '
          f'rel err {float((got - ref).norm() / ref.norm()):.2e}', flush=True)
