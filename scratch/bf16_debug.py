import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from graphgym_b200 import ops
from util import powerlaw_graph
cuda = torch.device('cuda')
for f in (32, 128):
    n = 30000
    ei = powerlaw_graph(3, n, 16)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(f))
    csr = ops.layout_build(ei.to(cuda), n, 1, 0)
    w = ops.gcn_norm(csr, ops.segment_degree(csr))
    xb = ops.cast_bf16(x.to(cuda))
    a = torch.sparse_coo_tensor(torch.stack([csr.rowid.cpu().long(), csr.nbr.cpu().long()]), w.cpu().double(), (n, n)).coalesce()
    want = torch.sparse.mm(a, xb.cpu().double())
    got = ops.spmm(csr, xb, w).cpu().double()
    err = (got - want).abs().max(dim=1).values
    deg = torch.diff(csr.rowptr.cpu())
    bad = torch.nonzero(err > 1e-3).flatten()
    print('f', f, 'bad rows', bad.numel(), 'of', n, 'max err', float(err.max()))
    item_row, item_slot, items = csr.plan
    ir = item_row.cpu()
    for r in bad[:8].tolist():
        k = int(torch.searchsorted(ir, r, right=True)) - 1
        print('  row', r, 'deg', int(deg[r]), 'rowptr', int(csr.rowptr[r]), 'item', k, 'item_row', ir[k:k+2].tolist(), 'item_slot', item_slot.cpu()[k:k+2].tolist(),
              'err cols', torch.nonzero((got[r]-want[r]).abs() > 1e-3).flatten()[:12].tolist(), 'ratio', float((got[r]/want[r]).median()))
    ref32 = ops.spmm(csr, xb.float(), w).cpu().double()
    print('  fp32 kernel on the same rounded rows: max err', float((ref32 - want).abs().max()))
print('--- with bias / mean / self')
f = 32
n = 30000
ei = powerlaw_graph(3, n, 16)
g = torch.Generator().manual_seed(f)
x = torch.randn(n, f, generator=g)
bias = torch.randn(f, generator=g)
csr = ops.layout_build(ei.to(cuda), n, 1, 0)
w = ops.gcn_norm(csr, ops.segment_degree(csr))
xb = ops.cast_bf16(x.to(cuda))
a = torch.sparse_coo_tensor(torch.stack([csr.rowid.cpu().long(), csr.nbr.cpu().long()]), w.cpu().double(), (n, n)).coalesce()
want = torch.sparse.mm(a, xb.cpu().double())
bc = bias.to(cuda)
got0 = ops.spmm(csr, xb, w).cpu().double()
got1 = ops.spmm(csr, xb, w, ops.SUM, None, 0.0, bc).cpu().double()
print('no bias err', float((got0 - want).abs().max()), 'bias err', float((got1 - want - bias.double()).abs().max()))
d = got1 - got0
print('got1-got0 row0', d[0, :8].tolist(), 'bias', bias[:8].tolist())
bad = torch.nonzero((d - bias.double()).abs().max(dim=1).values > 1e-3).flatten()
print('rows where added != bias:', bad.numel(), bad[:10].tolist())
if bad.numel():
    r = int(bad[0]); print('row', r, 'deg', int(torch.diff(csr.rowptr.cpu())[r]), (d[r] - bias.double())[:16].tolist())
