"""Per-step memory and time of the e2e pipeline (bench.py's public-API leg): reserved / allocated HBM, device allocations."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from graphgym_b200 import ops
from graphgym_b200.graph import get_layout, clear_cache, _CACHE
from graphgym_b200.models.layer import Batch, layer_dict
dev = torch.device('cuda')
spec = bench.WORKLOADS['products_gcn']
n, ei = bench.gen_graph(spec, dev)
x = bench.gen_features(n, 100, dev); gy = bench.gen_features(n, 128, dev, seed=7)
layer = layer_dict['gcnconv'](100, 128, bias=True).to(dev)
x_host, ei_host = x.cpu().pin_memory(), ei.cpu().pin_memory()
side = torch.cuda.Stream()
res_host = torch.empty(128).pin_memory()

def issue():
    with torch.cuda.stream(side):
        eid = ei_host.to(dev, non_blocking=True); e1 = torch.cuda.Event(); e1.record(side)
        xd = x_host.to(dev, non_blocking=True); e2 = torch.cuda.Event(); e2.record(side)
    return eid, xd, e1, e2

def compute(eid, xd, e1, e2):
    cur = torch.cuda.current_stream()
    cur.wait_event(e1); eid.record_stream(cur)
    lay = get_layout(eid, n, 1); _ = lay.csr, lay.csc
    cur.wait_event(e2); xd.record_stream(cur)
    layer.zero_grad(set_to_none=True)
    xg = xd.detach().requires_grad_(True)
    b = layer(Batch(xg, eid, None)); b.node_feature.backward(gy)
    res_host.copy_(layer.model.bias.grad, non_blocking=True)

nxt = issue()
for i in range(14):
    t0 = time.perf_counter()
    cur_in = nxt
    nxt = issue()
    compute(*cur_in)
    del cur_in
    torch.cuda.synchronize()
    st = torch.cuda.memory_stats()
    print(f'step {i}: {1e3 * (time.perf_counter() - t0):7.1f} ms  reserved {torch.cuda.memory_reserved() / 2**30:6.1f} GiB  allocated '
          f'{torch.cuda.memory_allocated() / 2**30:6.1f} GiB  device_allocs {st["num_device_alloc"]}  cache entries {len(_CACHE)}', flush=True)
