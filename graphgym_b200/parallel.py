"""Row-partitioned message passing over the GPUs of one box.

The reference is single-process (no torch.distributed / NCCL anywhere); this is the part of the path
that shards naturally: one process per GPU, rank r owns the target rows [lo_r, hi_r) of the adjacency,
their CSR rows (global source ids) and the matching output rows.

  forward   H_r = X_r W (local rows)  ->  all-gather of the owned H blocks over NVLink (the "halo":
            on a power-law graph nearly every remote row is referenced, so the halo is the full
            matrix)  ->  local CSR SpMM over H_full.
            Pipelined variant: P-1 send/recv rounds, the SpMM of block b runs while the later blocks are
            still in flight (per-peer sub-layouts, partial sums accumulated in a fixed block order).
  backward  all-gather of dOut  ->  local CSC SpMM (rows = owned SOURCE nodes, global target ids)
            gives dH_r  ->  dW_r = X_r^T dH_r, all-reduced with the other (tiny) parameter gradients.
Both directions use the same collective, and every reduction on the data path is rank-local in a
fixed order: the N-GPU result equals the 1-GPU result up to fp32 re-association (rows that the
merge-path plan splits at different places, and the all-reduced dW).

The collective is issued through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

from . import functional as F_
from . import ops


class RowPartition:
    """Contiguous, equal-sized row blocks: rank r owns [lo, hi) = [r*per, min(n, (r+1)*per))."""

    def __init__(self, num_nodes, world_size, rank):
        self.n, self.world, self.rank = int(num_nodes), int(world_size), int(rank)
        self.per = (self.n + self.world - 1) // self.world
        self.lo = min(self.n, self.rank * self.per)
        self.hi = min(self.n, self.lo + self.per)

    @property
    def rows(self):
        return self.hi - self.lo

    def bounds(self, r):
        lo = min(self.n, r * self.per)
        return lo, min(self.n, lo + self.per)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def all_gather_rows(local, part, group=None):
    """[rows_r, F] blocks -> [n, F]; blocks are padded to ``per`` rows for the collective."""
    f = local.size(1)
    if part.world == 1:
        return local
    padded = local
    if local.size(0) != part.per:
        padded = local.new_zeros((part.per, f))
        padded[:local.size(0)] = local
    full = local.new_empty((part.world * part.per, f))
    dist.all_gather_into_tensor(full, padded.contiguous(), group=group)
    return full[:part.n]


def exchange_blocks(local, part, group=None):
    """Start the P-1 rounds of the halo exchange: in round s every rank sends its block to rank+s and
    receives the block of rank-s (all pairs are distinct, NVSwitch serves them at full bandwidth).  Returns
    [(source rank, buffer, works)] in arrival order; ``work.wait()`` makes the current stream wait for
    that round only, so the aggregation of block rank-s overlaps the transfer of the later rounds."""
    per, f = part.per, local.size(1)
    padded = local
    if local.size(0) != per:
        padded = local.new_zeros((per, f))
        padded[:local.size(0)] = local
    padded = padded.contiguous()
    rounds = []
    for s in range(1, part.world):
        dst, src = (part.rank + s) % part.world, (part.rank - s) % part.world
        buf = local.new_empty((per, f))
        works = dist.batch_isend_irecv([dist.P2POp(dist.isend, padded, dst, group),
                                        dist.P2POp(dist.irecv, buf, src, group)])
        rounds.append((src, buf, works))
    return rounds


class PartitionedLayout:
    """Rank-local CSR (owned targets) and CSC (owned sources) of one edge_index under a self-loop policy,
    with the per-slot weights of the aggregation kinds (global degrees are all-gathered once).

    ``pipelined=True`` additionally splits both layouts by the rank that owns the neighbour
    (``sub_csr[b]`` / ``sub_csc[b]``: slots whose neighbour lives in block b), so that a peer's block can be
    aggregated as soon as it has arrived instead of after a full all-gather."""

    def __init__(self, edge_index, num_nodes, policy, part, group=None, pipelined=False):
        self.part, self.group, self.policy = part, group, policy
        self.num_nodes = int(num_nodes)
        self.pipelined = bool(pipelined) and part.world > 1
        rng = (part.lo, part.hi)
        self.csr = ops.layout_build(edge_index, num_nodes, policy, ops.BY_TARGET, row_range=rng)
        self.csc = ops.layout_build(edge_index, num_nodes, policy, ops.BY_SOURCE, row_range=rng)
        self._weights = {}
        self._sub_weights = {}
        self.sub_csr = self.sub_csc = None
        if self.pipelined:
            blocks = [part.bounds(b) for b in range(part.world)]
            self.sub_csr = [ops.layout_build(edge_index, num_nodes, policy, ops.BY_TARGET, row_range=rng,
                                             nbr_range=b) for b in blocks]
            self.sub_csc = [ops.layout_build(edge_index, num_nodes, policy, ops.BY_SOURCE, row_range=rng,
                                             nbr_range=b) for b in blocks]

    def sub_weights(self, kind):
        """Per-block (w_csr[b], w_csc[b]) lists; the normalisation uses the GLOBAL degrees."""
        if kind not in self._sub_weights:
            if kind == 'sum':
                w = ([None] * self.part.world, [None] * self.part.world)
            elif kind in ('gcn_src', 'gcn_tgt'):
                deg = self._global_degree(self.csc if kind == 'gcn_src' else self.csr)
                w = ([ops.gcn_norm(c, deg) for c in self.sub_csr], [ops.gcn_norm(c, deg) for c in self.sub_csc])
            else:
                raise NotImplementedError(f'pipelined halo exchange: aggregation kind {kind!r}')
            self._sub_weights[kind] = w
        return self._sub_weights[kind]

    def _global_degree(self, local_layout):
        deg_local = ops.segment_degree(local_layout).view(-1, 1)
        return all_gather_rows(deg_local, self.part, self.group).reshape(-1).contiguous()

    def weights(self, kind):
        if kind not in self._weights:
            if kind == 'sum':
                w = (None, None)
            elif kind == 'mean':
                indeg = self._global_degree(self.csr)
                w = (None, ops.mean_weights(self.csc, indeg))
            elif kind in ('gcn_src', 'gcn_tgt'):
                deg = self._global_degree(self.csc if kind == 'gcn_src' else self.csr)
                w = (ops.gcn_norm(self.csr, deg), ops.gcn_norm(self.csc, deg))
            else:
                raise KeyError(kind)
            self._weights[kind] = w
        return self._weights[kind]


class _DistAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_local, playout, kind, bias):
        part = playout.part
        h_full = all_gather_rows(h_local.contiguous(), part, playout.group)
        w_fwd, _ = playout.weights(kind)
        out = ops.spmm(playout.csr, h_full, w_fwd, ops.MEAN if kind == 'mean' else ops.SUM, None, 0.0, bias)
        ctx.playout, ctx.kind, ctx.has_bias = playout, kind, bias is not None
        return out

    @staticmethod
    def backward(ctx, g_local):
        playout = ctx.playout
        gh = gb = None
        if ctx.needs_input_grad[0]:
            g_full = all_gather_rows(g_local.contiguous(), playout.part, playout.group)
            _, w_bwd = playout.weights(ctx.kind)
            gh = ops.spmm(playout.csc, g_full, w_bwd, ops.SUM)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            gb = ops.colsum(g_local.contiguous())
        return gh, None, None, gb


def _pipelined_spmm(subs, weights, local, playout, bias):
    """out = sum_b A_b X_b: the local block first, then every peer block as its round completes."""
    part = playout.part
    rounds = exchange_blocks(local, part, playout.group)
    out = ops.spmm(subs[part.rank], local, weights[part.rank], ops.SUM, None, 0.0, bias, x_row_base=part.lo)
    for src, buf, works in rounds:
        for w in works:
            w.wait()
        ops.spmm(subs[src], buf, weights[src], ops.SUM, out, 1.0, None, out=out, x_row_base=part.bounds(src)[0])
    return out


class _DistAggregatePipelined(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_local, playout, kind, bias):
        w_fwd, _ = playout.sub_weights(kind)
        out = _pipelined_spmm(playout.sub_csr, w_fwd, h_local.contiguous(), playout, bias)
        ctx.playout, ctx.kind, ctx.has_bias = playout, kind, bias is not None
        return out

    @staticmethod
    def backward(ctx, g_local):
        playout = ctx.playout
        gh = gb = None
        g_local = g_local.contiguous()
        if ctx.needs_input_grad[0]:
            _, w_bwd = playout.sub_weights(ctx.kind)
            gh = _pipelined_spmm(playout.sub_csc, w_bwd, g_local, playout, None)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            gb = ops.colsum(g_local)
        return gh, None, None, gb


def dist_aggregate(h_local, playout, kind='sum', bias=None):
    """Row-partitioned ``functional.aggregate``: out_r = (A H)[rows of rank r] (+ bias)."""
    if playout.pipelined:
        return _DistAggregatePipelined.apply(h_local, playout, kind, bias)
    return _DistAggregate.apply(h_local, playout, kind, bias)


def allreduce_grads(module, group=None):
    """Sum the (replicated) parameters' gradients over the ranks — the only other exchange of a layer."""
    world, _ = _world(group)
    if world == 1:
        return
    for p in module.parameters():
        if p.grad is not None:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)


class RowPartitionedGCN(torch.nn.Module):
    """``gcnconv`` on a row partition: same parameters / init as models.layer._GCNConvLayer; forward takes
    the rank's rows of X and returns the rank's rows of the output."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .models.layer import _GCNConvLayer
        self.model = _GCNConvLayer(dim_in, dim_out, bias=bias)

    def forward(self, x_local, playout):
        h = F_.seg_linear([x_local], [self.model.weight], [(0, 0, False)])
        return dist_aggregate(h, playout, 'gcn_tgt', self.model.bias)
