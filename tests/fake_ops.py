"""CPU stand-ins for the gg_* kernels, built on the oracle, used ONLY by the world_size-2 gloo test of the
row-partitioned host logic (there is no GPU in the CPU test tier).  They mirror the signatures of
graphgym_b200.ops; the real ops reject CPU tensors."""
import numpy as np
import torch

from graphgym_b200 import ops
from oracle import layout as olayout


def layout_build(edge_index, num_nodes, policy=0, group_by=0, row_range=None, nbr_range=None):
    n = int(num_nodes)
    lo, hi = (0, n) if row_range is None else row_range
    nlo, nhi = (0, n) if nbr_range is None else nbr_range
    src, tgt, eid = olayout.edited_edges(edge_index.numpy(), n, policy)
    key, other = (tgt, src) if group_by == 0 else (src, tgt)
    keep = (key >= lo) & (key < hi) & (other >= nlo) & (other < nhi)
    key, other, eid = key[keep], other[keep], eid[keep]
    order = np.argsort(key, kind='stable')
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(key, minlength=n), out=rowptr[1:])
    t = lambda a: torch.from_numpy(a.astype(np.int32))
    return ops.Csr(t(rowptr[lo:hi + 1]), t(other[order]), t(eid[order]), t(key[order]), int(rowptr[-1]), hi - lo,
                   edge_index.size(1), policy, group_by)


def segment_degree(csr, w_slot=None):
    assert w_slot is None
    return (csr.rowptr[1:] - csr.rowptr[:-1]).float()


def gcn_norm(csr, deg, w_slot=None):
    dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
    return dis[csr.rowid.long()] * dis[csr.nbr.long()]


def mean_weights(csr_t, deg):
    inv = torch.where(deg > 0, 1.0 / deg, torch.zeros_like(deg))
    return inv[csr_t.nbr.long()]


def spmm(csr, x, w_slot=None, reduce=0, x_self=None, self_scale=0.0, bias=None, out=None, rank1=None,
         x_row_base=0, out_peers=None):
    assert out_peers is None, 'peer memory needs CUDA: the gloo tier runs the all-to-all form'
    rows = csr.num_nodes
    seg = torch.repeat_interleave(torch.arange(rows), (csr.rowptr[1:] - csr.rowptr[:-1]).long())
    msg = x[csr.nbr.long() - int(x_row_base)]
    if w_slot is not None:
        msg = msg * w_slot.view(-1, 1)
    res = torch.zeros((rows, x.size(1)), dtype=x.dtype).index_add_(0, seg, msg)
    if reduce == 1:
        res = res / (csr.rowptr[1:] - csr.rowptr[:-1]).clamp(min=1).view(-1, 1)
    if x_self is not None:
        res = res + self_scale * x_self
    if bias is not None:
        res = res + bias
    if rank1 is not None:
        s1, v1, s2, v2 = rank1
        res = res + s1.view(-1, 1) * v1.view(1, -1) + s2.view(-1, 1) * v2.view(1, -1)
    if out is not None:
        out.copy_(res)
        return out
    return res


# ---- the GAT passes (heads = 1), restated per slot ----------------------------------------------------
def slot_map(a, b):
    inv = torch.empty(a.num_edges + a.num_nodes + 1, dtype=torch.long)
    inv[a.perm.long()] = torch.arange(a.num_slots)
    return inv[b.perm.long()].to(torch.int32)


def gat_scores(h, att_row):
    c = h.size(1)
    return h @ att_row[0, :c], h @ att_row[0, c:]


def _seg_sum(csr, v):
    return torch.zeros(csr.num_nodes, dtype=v.dtype).index_add_(0, csr.rowid.long(), v)


def gat_alpha(csr, a_tgt, a_src, slope):
    rows, nbr = csr.rowid.long(), csr.nbr.long()
    z = torch.nn.functional.leaky_relu(a_tgt[rows] + a_src[nbr], slope)
    zmax = torch.full((csr.num_nodes,), -float('inf')).scatter_reduce(0, rows, z, reduce='amax')
    e = torch.exp(z - zmax[rows])
    return e / (_seg_sum(csr, e)[rows] + 1e-16)


def gat_sddmm_slice(csr, h_slice, g_slice):
    return (g_slice[csr.rowid.long()] * h_slice[csr.nbr.long()]).sum(1)


def gat_dz(csr, a_tgt, a_src, alpha, dalpha, slope):
    rows, nbr = csr.rowid.long(), csr.nbr.long()
    z = a_tgt[rows] + a_src[nbr]
    d = _seg_sum(csr, alpha * dalpha)
    dz = alpha * (dalpha - d[rows]) * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, slope))
    return dz, _seg_sum(csr, dz)


def gat_csc_gather(csc, csc2csr, alpha, dz):
    m = csc2csr.long()
    return alpha[m], _seg_sum(csc, dz[m])


def gat_att_grad(h, da_tgt, da_src):
    return torch.cat([da_tgt @ h, da_src @ h]).view(1, -1)


def id_gemm(segments, n, f, b_trans=False, bias=None, act=0, relu_mask=None, out=None):
    res = torch.zeros((n, f))
    for a, b, scale in segments:
        y = a @ (b.t() if b_trans else b)
        res = res + (y if scale is None else y * scale.view(-1, 1))
    if bias is not None:
        res = res + bias
    if act == 1:
        res = res.relu()
    if relu_mask is not None:
        res = res * (relu_mask > 0)
    return res


def gemm_tn(a, g, row_index=None):
    if row_index is not None:
        a, g = a[row_index], g[row_index]
    return a.t() @ g


def colsum(g):
    return g.sum(0)


def relu_grad(g, y):
    return g * (y > 0)


def gather_rows(x, ids):
    return x[ids.long()]


def scatter_add_rows_(out, ids, x):
    return out.index_add_(0, ids.long(), x)


def id_count(ids, num_nodes):
    return torch.bincount(ids.long(), minlength=num_nodes).float()


def install(monkeypatch_setattr):
    for name in ('layout_build', 'segment_degree', 'gcn_norm', 'mean_weights', 'spmm', 'id_gemm', 'gemm_tn',
                 'colsum', 'relu_grad', 'id_count', 'gather_rows', 'scatter_add_rows_', 'slot_map', 'gat_scores',
                 'gat_alpha', 'gat_sddmm_slice', 'gat_dz', 'gat_csc_gather', 'gat_att_grad'):
        monkeypatch_setattr(ops, name, globals()[name])
