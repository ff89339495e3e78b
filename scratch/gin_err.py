import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from graphgym_b200 import ops
from graphgym_b200.config import reset_cfg
from graphgym_b200.models.layer import Batch, layer_dict
from util import powerlaw_graph, rel_err
from test_layers_gpu import _oracle, run_ours
dev = torch.device('cuda')
n, fin, fout = 2708, 1433, 128
for name in ('ginconv', 'gcnconv', 'sageconv', 'ginidconv'):
    for seed in range(4):
        for mode in ('tc', 'simt'):
            ops.GEMM_MODE = mode
            reset_cfg()
            torch.manual_seed(seed)
            ei = powerlaw_graph(n, n, 8)
            g = torch.Generator().manual_seed(n + fin)
            x = torch.randn(n, fin, generator=g)
            ids = torch.randperm(n, generator=g)[: n // 10].sort().values
            layer = layer_dict[name](fin, fout, bias=True)
            params = {k: v.detach().clone() for k, v in layer.named_parameters()}
            xd, yo, P = _oracle(name, x, ei, ids, params)
            gy = torch.randn(n, fout, generator=g)
            yo.backward(gy.double())
            y, gx, grads = run_ours(layer, x, ei, ids, gy, dev)
            errs = {'y': rel_err(y, yo.detach()), 'gx': rel_err(gx, xd.grad)}
            errs.update({k.split('.', 1)[1]: rel_err(v, P[k].grad) for k, v in grads.items()})
            print(name, seed, mode, {k: f'{v:.1e}' for k, v in errs.items()})
