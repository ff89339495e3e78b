"""Per-pass times of the GAT layer on the products-shaped graph: fused sliced-ELL passes vs the split merge-path passes."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import bench
from graphgym_b200 import ops
from graphgym_b200.graph import GraphLayout
dev = torch.device('cuda')
n, ei = bench.gen_graph(bench.WORKLOADS['products_gat'], dev)
lay = GraphLayout(ei, n, ops.LOOPS_REMOVE_ADD)
csr, csc, m = lay.csr, lay.csc, lay.csc2csr
f = 128
h = torch.randn(n, f, device=dev)
g = torch.randn(n, f, device=dev)
att = torch.randn(1, 1, 2 * f, device=dev) * 0.1
bias = torch.zeros(f, device=dev)


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


for algo in ('sell', 'sell_two', 'mp'):
    ops.GAT_ALGO = algo.split('_')[0]
    ops.GAT_BWD = 'two' if algo.endswith('two') else 'one'
    out, al, a_tgt, a_src, pos = ops.gat_forward(csr, h, att, 1, 0.2, bias, True)
    t_f = timeit(lambda: ops.gat_forward(csr, h, att, 1, 0.2, bias, True))
    t_b = timeit(lambda: ops.gat_backward(csr, csc, m, h, att, 1, 0.2, bias, al, a_tgt, a_src, out, g, pos))
    print(f'{algo}: forward {t_f:.3f} ms, backward {t_b:.3f} ms', flush=True)
ops.GAT_ALGO = 'sell'
ops.GAT_BWD = 'one'
out, rs, a_tgt, a_src, pos = ops.gat_forward(csr, h, att, 1, 0.2, bias, True)
at, asr = a_tgt.view(-1), a_src.view(-1)
print('fwd kernel only: eval', timeit(lambda: ops.gat_sell_forward(csr, h, f, at, asr, 0.2, bias)), 'train', timeit(lambda: ops.gat_sell_forward(csr, h, f, at, asr, 0.2, bias, True)))
print('spmm sell weighted', timeit(lambda: ops.spmm(csr, h, torch.ones(csr.num_slots, device=dev), algo='sell')))
L = ops.lib()
import ctypes
from graphgym_b200.ops import _ptr, _stream, check
s1, s2 = ops.sell_layout(csr), ops.sell_layout(csc)
dz = torch.empty(csr.num_slots, device=dev); da_t = torch.empty(n, device=dev); da_s = torch.empty(n, device=dev)
ws_bytes = int(L.gg_gat_sell_workspace_bytes(max(s1.partial_rows, s2.partial_rows), f)); ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
def edge():
    check(L.gg_gat_sell_bwd_edge_f32(_ptr(s1.chunk_ptr), s1.chunks, _ptr(s1.idx), _ptr(s1.slot_of), _ptr(s1.vdst), _ptr(s1.hub_rows), _ptr(s1.hub_pptr), s1.hubs, s1.partial_rows, _ptr(h), f, _ptr(g), f, _ptr(out), f, _ptr(bias), _ptr(at), _ptr(asr), _ptr(rs), n, f, 0.2, _ptr(dz), _ptr(da_t), _ptr(ws), ws_bytes, _stream()), 'edge')
emap = ops._sell_edge_map(csc, m)
dh = torch.empty(n, f, device=dev); tstat = torch.empty(n, 4, device=dev)
a_t, a_s = att[0, 0, :f].contiguous(), att[0, 0, f:].contiguous()
def src():
    check(L.gg_gat_sell_bwd_src_f32(_ptr(s2.chunk_ptr), s2.chunks, _ptr(s2.idx), _ptr(emap), _ptr(s2.vdst), _ptr(s2.hub_rows), _ptr(s2.hub_pptr), s2.hubs, s2.partial_rows, _ptr(g), f, _ptr(at), _ptr(asr), _ptr(rs), _ptr(dz), _ptr(da_t), _ptr(a_s), _ptr(a_t), n, f, 0.2, _ptr(dh), f, _ptr(da_s), _ptr(tstat), _ptr(ws), ws_bytes, _stream()), 'src')
print('bwd edge kernel', timeit(edge)); print('bwd src kernel', timeit(src))

da_t2 = torch.empty(n, device=dev); dh2 = torch.empty(n, f, device=dev); da_s2 = torch.empty(n, device=dev)
def one():
    check(L.gg_gat_sell_bwd_one_f32(_ptr(s2.chunk_ptr), s2.chunks, _ptr(s2.idx), _ptr(s2.vdst), _ptr(s2.hub_rows), _ptr(s2.hub_pptr), s2.hubs, s2.partial_rows, _ptr(h), f, _ptr(g), f, _ptr(out), f, _ptr(bias), _ptr(pos[0]), f, _ptr(pos[1]), _ptr(at), _ptr(asr), _ptr(rs), _ptr(a_s), _ptr(a_t), n, f, 0.2, _ptr(dh2), f, _ptr(da_t2), _ptr(da_s2), _ptr(tstat), _ptr(ws), ws_bytes, _stream()), 'one')
edge(); src(); one(); torch.cuda.synchronize()
rel = lambda a, b: float((a - b).norm() / b.norm())
print('bwd one (three launches)', timeit(one), 'vs two-pass: dh', rel(dh2, dh), 'da_tgt', rel(da_t2, da_t), 'da_src', rel(da_s2, da_s))
