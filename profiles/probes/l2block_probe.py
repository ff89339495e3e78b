"""Round-2 probe: source-blocked (L2-resident) narrow aggregation.  Each rank of the feature-sliced exchange walks ALL
slots at F/P columns and is slot-rate bound (43 G slots/s, DRAM row activations).  Here the sources are cut into K blocks
whose column slice fits the L2; block b's sub-layout is aggregated into the same output (accumulating epilogue).

    gpurun -- 'python profiles/probes/l2block_probe.py'
"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gcn'], dev)
csr = ops.layout_build(ei, n, 1, 0)
deg = ops.segment_degree(csr)
w = ops.gcn_norm(csr, deg)


def timeit(fn, it=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


for K in (1, 2, 3, 4, 6, 8):
    per = (n + K - 1) // K
    subs = [ops.layout_build(ei, n, 1, 0, nbr_range=(b * per, min(n, (b + 1) * per))) for b in range(K)]
    ws = [ops.gcn_norm(s, deg) for s in subs]
    for s in subs:
        _ = s.plan
    for f in (16, 32, 64, 128):
        if K > 4 and f > 32:
            continue
        x = torch.randn(n, f, device=dev)
        ref = ops.spmm(csr, x, w)

        def run():
            out = ops.spmm(subs[0], x, ws[0])
            for b in range(1, K):
                ops.spmm(subs[b], x, ws[b], ops.SUM, out, 1.0, None, out=out)
            return out
        t = timeit(run)
        err = float((run() - ref).abs().max() / ref.abs().max())
        print(f'K={K} f={f}: {t:.3f} ms  block={per * f * 4 / 1e6:.0f} MB  rel err {err:.1e}', flush=True)
    del subs, ws
