"""Fused layer post-ops (csrc/postops.cu) against the reference's own modules in fp64:
nn.BatchNorm1d (train and eval, running statistics included) -> activation -> F.normalize
(ref: graphgym/models/layer.py:26-46, graphgym/models/gnn.py:79-80)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from graphgym_b200 import functional as F_
from graphgym_b200 import ops
from graphgym_b200.config import cfg, reset_cfg
from util import FP32_TOL, rel_err

pytestmark = pytest.mark.gpu
ACTS = {ops.ACT_NONE: lambda t: t, ops.ACT_RELU: F.relu, ops.ACT_LRELU: lambda t: F.leaky_relu(t, 0.25)}


def reference(y, gy, bn64, act, l2):
    y = y.double().requires_grad_(True)
    a = bn64(y) if bn64 is not None else y
    r = ACTS[act](a)
    o = F.normalize(r, p=2, dim=1) if l2 else r
    o.backward(gy.double())
    return o.detach(), y.grad


@pytest.mark.parametrize('n,f', [(1000, 128), (777, 100), (5000, 33), (64, 256), (3, 8), (20000, 16), (300, 1024)])
@pytest.mark.parametrize('act', [ops.ACT_NONE, ops.ACT_RELU, ops.ACT_LRELU])
@pytest.mark.parametrize('l2', [False, True])
def test_bn_act_l2_train(cuda, n, f, act, l2):
    g = torch.Generator().manual_seed(n + f)
    y = torch.randn(n, f, generator=g) * 3 + 50 * torch.randn(f, generator=g)   # |mean| >> std columns
    gy = torch.randn(n, f, generator=g)
    bn = nn.BatchNorm1d(f, eps=1e-5, momentum=0.1)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-1, 1, generator=g)
    bn64 = nn.BatchNorm1d(f, eps=1e-5, momentum=0.1).double()
    bn64.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    bn = bn.to(cuda)
    want, want_dy = reference(y, gy, bn64, act, l2)
    yc = y.to(cuda).requires_grad_(True)
    out = F_.post_ops(yc, bn, True, act, 0.25, l2)
    out.backward(gy.to(cuda))
    assert rel_err(out, want) < FP32_TOL
    assert torch.allclose(out.detach().cpu().double(), want, rtol=1e-4, atol=1e-5)
    # gradients: ReLU gates of pre-activations within rounding of zero may differ; the band is checked like the GIN tests
    assert rel_err(yc.grad, want_dy) < (FP32_TOL if act == ops.ACT_NONE else 2e-5)
    assert rel_err(bn.weight.grad, bn64.weight.grad) < FP32_TOL
    assert rel_err(bn.bias.grad, bn64.bias.grad) < FP32_TOL
    assert rel_err(bn.running_mean, bn64.running_mean) < FP32_TOL
    assert rel_err(bn.running_var, bn64.running_var) < FP32_TOL
    assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize('l2', [False, True])
def test_eval_mode_and_no_bn(cuda, l2):
    n, f = 900, 64
    g = torch.Generator().manual_seed(3)
    y, gy = torch.randn(n, f, generator=g), torch.randn(n, f, generator=g)
    bn64 = nn.BatchNorm1d(f).double()
    with torch.no_grad():
        bn64.running_mean.uniform_(-1, 1, generator=g)
        bn64.running_var.uniform_(0.5, 2, generator=g)
        bn64.weight.uniform_(0.5, 1.5, generator=g)
    bn = nn.BatchNorm1d(f)
    bn.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in bn64.state_dict().items()})
    bn, bn64 = bn.to(cuda).eval(), bn64.eval()
    want, want_dy = reference(y, gy, bn64, ops.ACT_RELU, l2)
    yc = y.to(cuda).requires_grad_(True)
    out = F_.post_ops(yc, bn, False, ops.ACT_RELU, 0.0, l2)
    out.backward(gy.to(cuda))
    assert rel_err(out, want) < FP32_TOL and rel_err(yc.grad, want_dy) < 2e-5
    assert rel_err(bn.weight.grad, bn64.weight.grad) < FP32_TOL and rel_err(bn.bias.grad, bn64.bias.grad) < FP32_TOL
    assert int(bn.num_batches_tracked) == 0
    want, want_dy = reference(y, gy, None, ops.ACT_RELU, l2)
    yc = y.to(cuda).requires_grad_(True)
    out = F_.post_ops(yc, None, True, ops.ACT_RELU, 0.0, l2)
    out.backward(gy.to(cuda))
    assert rel_err(out, want) < FP32_TOL and rel_err(yc.grad, want_dy) < FP32_TOL


def test_all_zero_rows_l2(cuda):
    """ReLU kills a whole row: F.normalize returns 0 / eps = 0, gradient 0."""
    y = torch.full((10, 16), -1.0)
    y[3] = 1.0
    yc = y.to(cuda).requires_grad_(True)
    out = F_.post_ops(yc, None, True, ops.ACT_RELU, 0.0, True)
    out.backward(torch.ones(10, 16, device=cuda))
    assert torch.isfinite(out).all() and torch.isfinite(yc.grad).all()
    assert float(out[0].abs().max()) == 0.0 and abs(float(out[3].norm()) - 1.0) < 1e-6


def test_general_layer_fused_equals_unfused_modules(cuda):
    """GeneralLayer (ref: layer.py:16-47) with the fused post-ops == the same layer running its nn modules one by one."""
    from graphgym_b200.models.layer import Batch, GeneralLayer
    from util import random_graph
    reset_cfg()
    n, fin, fout = 2000, 48, 64
    ei = random_graph(1, n, 12000).to(cuda)
    g = torch.Generator().manual_seed(0)
    x, gy = torch.randn(n, fin, generator=g).to(cuda), torch.randn(n, fout, generator=g).to(cuda)
    res = {}
    for fused in (True, False):
        cfg.b200.fused_postops = fused
        torch.manual_seed(0)
        layer = GeneralLayer('gcnconv', fin, fout, has_act=True, has_bn=True, has_l2norm=True).to(cuda)
        xg = x.clone().requires_grad_(True)
        out = layer(Batch(xg, ei)).node_feature
        out.backward(gy)
        res[fused] = (out.detach(), xg.grad, layer.layer.model.weight.grad, layer.post_layer[0].weight.grad,
                      layer.post_layer[0].running_var.clone())
        assert sorted(layer.state_dict()) == sorted(['layer.model.weight', 'post_layer.0.weight', 'post_layer.0.bias',
                                                     'post_layer.0.running_mean', 'post_layer.0.running_var',
                                                     'post_layer.0.num_batches_tracked'])
    reset_cfg()
    for a, b in zip(res[True], res[False]):
        assert rel_err(a, b) < 2e-5
