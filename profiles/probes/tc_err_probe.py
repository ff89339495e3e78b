import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))); sys.path.insert(0, '/root/repo/tests')
from graphgym_b200 import ops
from util import rel_err
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
for k in (16, 64, 128, 256, 512, 1024, 1433, 2048, 4096):
    n, f = 2048, 128
    x, w = torch.randn(n, k, generator=g), torch.randn(k, f, generator=g)
    want = x.double() @ w.double()
    res = {}
    for mode in ('tc', 'simt'):
        ops.GEMM_MODE = mode
        got = ops.id_gemm([(x.to(dev), w.to(dev), None)], n, f)
        res[mode] = rel_err(got, want)
        bias_sign = float(((got.cpu().double() - want) * want.sign()).mean() / want.abs().mean())
        res[mode + '_bias'] = bias_sign
    print(k, {a: f'{b:.2e}' for a, b in res.items()})
# timing
import time
n, k, f = 2449029, 100, 128
x, w = torch.randn(n, k, device=dev), torch.randn(k, f, device=dev)
for mode in ('tc', 'simt'):
    ops.GEMM_MODE = mode
    for _ in range(3): ops.id_gemm([(x, w, None)], n, f)
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): ops.id_gemm([(x, w, None)], n, f)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(mode, 'products 100->128:', round(ms, 3), 'ms', round(2 * n * k * f / ms / 1e9, 1), 'TFLOP/s(fp32-equivalent)', round((n*k*4 + n*f*4) / ms / 1e6, 1), 'GB/s')
n, k, f = 1000000, 256, 256
x, w = torch.randn(n, k, device=dev), torch.randn(k, f, device=dev)
for mode in ('tc', 'simt'):
    ops.GEMM_MODE = mode
    for _ in range(3): ops.id_gemm([(x, w, None)], n, f)
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): ops.id_gemm([(x, w, None)], n, f)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(mode, 'ba1m 256->256:', round(ms, 3), 'ms', round(2 * n * k * f / ms / 1e9, 1), 'TFLOP/s(fp32-equivalent)')
# weight-gradient kernel
for (n, k, f) in ((2449029, 100, 128), (1000000, 256, 256)):
    a, gg_ = torch.randn(n, k, device=dev), torch.randn(n, f, device=dev)
    want = (a[:200000].double().t() @ gg_[:200000].double()).cpu()
    for mode in ('tc', 'simt'):
        ops.GEMM_MODE = mode
        err = rel_err(ops.gemm_tn(a[:200000], gg_[:200000]), want)
        for _ in range(3): ops.gemm_tn(a, gg_)
        torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): ops.gemm_tn(a, gg_)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(mode, f'gemm_tn n={n} {k}x{f}: {ms:.3f} ms, {(n*(k+f)*4)/ms/1e6:.0f} GB/s, err(200K rows) {err:.2e}')
