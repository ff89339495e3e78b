"""Characterise run-to-run differences of the persistent GEMMs (which rows / tiles / magnitudes), with other kernels
interleaved as in gemm_stress.py.  GG_DIAG_MODE: masked | act | plain."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graphgym_b200 import ops
dev = torch.device('cuda')
torch.manual_seed(0)
mode = os.environ.get('GG_DIAG_MODE', 'masked')
inter = os.environ.get('GG_DIAG_INTER', '1') == '1'
n, k, f = 2449029, 100, 128
x = torch.randn(n, k, device=dev); w = torch.randn(k, f, device=dev); b = torch.randn(f, device=dev)
g = torch.randn(n, f, device=dev); small = torch.randn(3000, k, device=dev)


def run():
    if mode == 'masked':
        return ops.id_gemm([(g, w, None)], n, k, b_trans=True, relu_mask=x)
    if mode == 'act':
        return ops.id_gemm([(x, w, None)], n, f, bias=b, act=ops.ACT_RELU)
    return ops.id_gemm([(g, w, None)], n, k, b_trans=True)


ref = run().clone()
fo = ref.size(1)
bad = 0
for it in range(16):
    if inter:
        ops.id_gemm([(small, w, None)], 3000, f)
        ops.gemm_tn(x, g)
    o = run()
    d = (o != ref)
    if bool(d.any()):
        bad += 1
        rows = d.any(1).nonzero().view(-1)
        if bad <= 2:
            tiles = torch.unique(rows // 128)
            mag = float((o - ref)[d].abs().max())
            print(f'it {it}: {int(d.sum())} elements in {rows.numel()} rows, {tiles.numel()} tiles; distinct rows%128 {len(set((rows % 128).tolist()))} head {sorted(set((rows % 128).tolist()))[:8]}; '
                  f'distinct tiles%148 {len(set((tiles % 148).tolist()))}; tile//148 {sorted(set((tiles // 148).tolist()))[:12]}; max|diff| {mag:.3e} vs max|ref| {float(ref.abs().max()):.2e}; '
                  f'cols with diffs {int((d.sum(0) > 0).sum())}; rows with all nonzero cols differing {int((d.sum(1) >= (ref != 0).sum(1).clamp(min=1)).sum())}', flush=True)
            r0 = int(rows[0]); print('   row', r0, 'o', o[r0, :4].tolist(), 'ref', ref[r0, :4].tolist(), flush=True)
            # does the bad row equal another row's result (stale / shifted data)?
            hit = (ref == o[r0]).all(1).nonzero().view(-1)
            print('   same as ref row(s):', hit[:4].tolist(), flush=True)
print(f'mode {mode} inter {inter} flags {os.environ.get("GG_PS_FLAGS")}: launches with differences {bad} of 16', flush=True)
