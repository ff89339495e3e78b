"""1-GPU probe of the per-rank work of the feature-sliced exchange: full products-shaped graph, F/P-wide slices."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from graphgym_b200 import ops
dev = torch.device('cuda')
spec = bench.WORKLOADS['products_gcn']
n, ei = bench.gen_graph(spec, dev)
csr = ops.layout_build(ei, n, 1, 0)
w = ops.gcn_norm(csr, ops.segment_degree(csr))
def timeit(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
import os
for hint in (False, True):
  ops.SPMM_L2HINT = hint
  print('L2 hints', hint)
  for fs in (16, 32, 64, 128):
      x = torch.randn(n, fs, device=dev)
      ops.SPMM_ALGO = 'mpg'
      ms = timeit(lambda: ops.spmm(csr, x, w))
      by = bench.spmm_bytes(n, csr.num_slots, fs, True)
      print(f'mpg fs={fs}: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s algorithmic', flush=True)
      if fs > 64:
          ops.SPMM_ALGO = 'mp'
          ms = timeit(lambda: ops.spmm(csr, x, w))
          print(f'mp  fs={fs}: {ms:.3f} ms  {by/ms/1e6:.0f} GB/s algorithmic', flush=True)
      P = 128 // fs
      if P > 1:
          per = (n + P - 1) // P
          blocks = [torch.empty(per, 128, device=dev) for _ in range(P)]
          peers = ops.PeerRows([b.data_ptr() for b in blocks], per, 128)
          ms = timeit(lambda: ops.spmm(csr, x, w, out_peers=peers))
          print(f'mpg fs={fs} peer-out(local): {ms:.3f} ms', flush=True)
          src = torch.randn(per, 128, device=dev)
          slices = [torch.empty(n, fs, device=dev) for _ in range(P)]
          import ctypes
          from graphgym_b200 import parallel
          ms = timeit(lambda: parallel._peer_scatter_cols(src, [s.data_ptr() for s in slices], 0))
          print(f'scatter_cols P={P}: {ms:.3f} ms', flush=True)
