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
  feature-sliced exchange (``exchange='sliced'``): the all-gather moves N*F*4 bytes INTO every rank however
            many ranks there are, so at 8 GPUs the two halo all-gathers are 60 % of the step.  Here the
            exchange is a transposition instead: rank c receives column slice c (F/P wide) of every rank's
            rows — (P-1)/P^2 * N*F*4 bytes per rank, 8x less at P = 8 — aggregates that slice over the
            WHOLE graph (every rank holds the full CSR/CSC; narrow rows run on the sub-warp-group kernel) and
            returns finished rows to their owners.  Inputs and outputs stay row-partitioned.  Both legs run
            over peer memory (csrc/peer.cu): a column-scatter kernel stores into the peers' slices, and the
            SpMM's epilogue stores each finished row into the owner's block, i.e. aggregation and exchange
            are ONE kernel; ranks are ordered by a flag barrier in peer memory — no NCCL on the data path.
            ``exchange='sliced_nccl'`` is the same algorithm over ``all_to_all_single`` (gloo in the CPU tests).
Both directions use the same collective, and every reduction on the data path is rank-local in a
fixed order: the N-GPU result equals the 1-GPU result up to fp32 re-association (rows that the
merge-path plan splits at different places, and the all-reduced dW).

The collective is issued through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import functional as F_
from . import ops
from ._lib import check, lib


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

    def __init__(self, edge_index, num_nodes, policy, part, group=None, pipelined=False, exchange=None,
                 pool=None):
        self.part, self.group, self.policy = part, group, policy
        self.num_nodes = int(num_nodes)
        self.exchange = exchange or ('pipelined' if pipelined else 'allgather')
        if part.world == 1 and self.exchange != 'allgather':
            self.exchange = 'allgather'
        if self.exchange not in ('allgather', 'pipelined', 'sliced', 'sliced_nccl'):
            raise ValueError(f'unknown halo exchange {self.exchange!r}')
        self.pipelined = self.exchange == 'pipelined'
        self.sliced = self.exchange in ('sliced', 'sliced_nccl')
        self.pool = pool
        self._weights = {}
        self._sub_weights = {}
        self._groups = {}
        self._csc2csr = None
        self.sub_csr = self.sub_csc = None
        if self.sliced:
            # every rank aggregates its column slice over the whole graph: full layouts, no row filter
            self.csr = ops.layout_build(edge_index, num_nodes, policy, ops.BY_TARGET)
            self.csc = ops.layout_build(edge_index, num_nodes, policy, ops.BY_SOURCE)
            if self.exchange == 'sliced' and self.pool is None:
                self.pool = PeerPool.shared(part, group)
            return
        rng = (part.lo, part.hi)
        self.csr = ops.layout_build(edge_index, num_nodes, policy, ops.BY_TARGET, row_range=rng)
        self.csc = ops.layout_build(edge_index, num_nodes, policy, ops.BY_SOURCE, row_range=rng)
        if self.pipelined:
            blocks = [part.bounds(b) for b in range(part.world)]
            self.sub_csr = [ops.layout_build(edge_index, num_nodes, policy, ops.BY_TARGET, row_range=rng,
                                             nbr_range=b) for b in blocks]
            self.sub_csc = [ops.layout_build(edge_index, num_nodes, policy, ops.BY_SOURCE, row_range=rng,
                                             nbr_range=b) for b in blocks]

    def owner_groups(self, layout, ngroups):
        """The full-graph layout cut into `ngroups` runs of owners (row ranges): per group a Csr VIEW of its rows (shifted
        row pointer, slices of the slot arrays) for the pipelined push.  Built once per layout (layout-build level)."""
        key = (id(layout), ngroups)
        hit = self._groups.get(key)
        if hit is not None:
            return hit
        part, P = self.part, self.part.world
        if ngroups <= 1:
            groups = [dict(owners=(0, P), rows=(0, part.n), slots=(0, layout.num_slots), csr=layout, whole=True)]
        else:
            groups = []
            rp = layout.rowptr
            for k in range(ngroups):
                o0, o1 = k * P // ngroups, (k + 1) * P // ngroups
                lo, hi = min(part.n, o0 * part.per), min(part.n, o1 * part.per)
                if o1 == o0:
                    continue
                s0, s1 = int(rp[lo].item()), int(rp[hi].item())
                sub = ops.Csr((rp[lo:hi + 1] - rp[lo]).contiguous(), layout.nbr[s0:s1], layout.perm[s0:s1],
                              layout.rowid[s0:s1], s1 - s0, hi - lo, layout.num_edges, layout.policy, layout.group_by)
                groups.append(dict(owners=(o0, o1), rows=(lo, hi), slots=(s0, s1), csr=sub, whole=False))
        self._groups[key] = groups
        return groups

    @staticmethod
    def group_weights(group, w):
        """This group's slice of a per-slot weight array; the view object is kept so that the sliced-ELL layout's
        permutation cache (keyed by tensor identity) hits on every step."""
        if w is None or group['whole']:
            return w
        cache = group.setdefault('_w', {})
        key = (w.data_ptr(), w._version)
        hit = cache.get(key)
        if hit is None or hit[0] is not w:
            cache.clear()
            hit = cache[key] = (w, w[group['slots'][0]:group['slots'][1]])
        return hit[1]

    @property
    def csc2csr(self):
        """CSC slot -> CSR slot of the same edge (full-graph layouts of the sliced exchange only)."""
        if not self.sliced:
            raise NotImplementedError('slot map across rank-local layouts')
        if self._csc2csr is None:
            self._csc2csr = ops.slot_map(self.csr, self.csc)
        return self._csc2csr

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
        if self.sliced:   # the layout is the whole graph already
            return ops.segment_degree(local_layout)
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
    def forward(ctx, h_local, playout, kind, bias, self_scale):
        part = playout.part
        h_local = h_local.contiguous()
        h_full = all_gather_rows(h_local, part, playout.group)
        w_fwd, _ = playout.weights(kind)
        out = ops.spmm(playout.csr, h_full, w_fwd, ops.MEAN if kind == 'mean' else ops.SUM,
                       h_local if self_scale != 0.0 else None, self_scale, bias)
        ctx.playout, ctx.kind, ctx.has_bias, ctx.self_scale = playout, kind, bias is not None, self_scale
        return out

    @staticmethod
    def backward(ctx, g_local):
        playout = ctx.playout
        gh = gb = None
        g_local = g_local.contiguous()
        if ctx.needs_input_grad[0]:
            g_full = all_gather_rows(g_local, playout.part, playout.group)
            _, w_bwd = playout.weights(ctx.kind)
            gh = ops.spmm(playout.csc, g_full, w_bwd, ops.SUM, g_local if ctx.self_scale != 0.0 else None,
                          ctx.self_scale)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            gb = ops.colsum(g_local)
        return gh, None, None, gb, None


def _pipelined_spmm(subs, weights, local, playout, bias, self_scale=0.0):
    """out = sum_b A_b X_b (+ self_scale * X_local): the local block first, then every peer block as its
    round completes."""
    part = playout.part
    rounds = exchange_blocks(local, part, playout.group)
    out = ops.spmm(subs[part.rank], local, weights[part.rank], ops.SUM, local if self_scale != 0.0 else None,
                   self_scale, bias, x_row_base=part.lo)
    for src, buf, works in rounds:
        for w in works:
            w.wait()
        ops.spmm(subs[src], buf, weights[src], ops.SUM, out, 1.0, None, out=out, x_row_base=part.bounds(src)[0])
    return out


class _DistAggregatePipelined(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_local, playout, kind, bias, self_scale):
        w_fwd, _ = playout.sub_weights(kind)
        out = _pipelined_spmm(playout.sub_csr, w_fwd, h_local.contiguous(), playout, bias, self_scale)
        ctx.playout, ctx.kind, ctx.has_bias, ctx.self_scale = playout, kind, bias is not None, self_scale
        return out

    @staticmethod
    def backward(ctx, g_local):
        playout = ctx.playout
        gh = gb = None
        g_local = g_local.contiguous()
        if ctx.needs_input_grad[0]:
            _, w_bwd = playout.sub_weights(ctx.kind)
            gh = _pipelined_spmm(playout.sub_csc, w_bwd, g_local, playout, None, ctx.self_scale)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            gb = ops.colsum(g_local)
        return gh, None, None, gb, None


# ------------------------------------------------------------------------------------------------
# peer memory (csrc/peer.cu) and the feature-sliced exchange
# ------------------------------------------------------------------------------------------------
class _RawCuda:
    """Lets torch view a device allocation it does not own (``__cuda_array_interface__``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                         'version': 2}


class PeerBuffer:
    """The same allocation on every rank: ``ptrs[r]`` is rank r's copy as seen from THIS process,
    ``local`` a uint8 torch view of this rank's copy."""

    def __init__(self, ptrs, nbytes, rank, device):
        self.ptrs, self.nbytes = ptrs, nbytes
        self.local = torch.as_tensor(_RawCuda(ptrs[rank], nbytes), device=device)

    def view(self, *shape):
        n = 1
        for d in shape:
            n *= d
        return self.local[:n * 4].view(torch.float32).view(*shape)


class PeerPool:
    """Peer-visible buffers of one process group (one process per GPU of a box) plus the flag barrier.
    Allocation is collective: every rank must request the same keys in the same order."""
    _shared = {}

    @classmethod
    def shared(cls, part, group=None):
        key = (id(group), part.world, part.rank)
        if key not in cls._shared:
            cls._shared[key] = cls(part.world, part.rank, group)
        return cls._shared[key]

    def __init__(self, world, rank, group=None):
        if world > 8:
            raise ValueError('peer memory exchange: at most 8 ranks (one NVSwitch box)')
        self.world, self.rank, self.group = world, rank, group
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._bufs = {}
        self._owned, self._opened = [], []
        self.flags = self._alloc(256)
        self._flag_arr = (ctypes.c_void_p * world)(*self.flags.ptrs)
        self.side_stream = torch.cuda.Stream(device=self.device)   # carries the pipelined pushes of the return leg

    def _alloc(self, nbytes):
        L = lib()
        hb = int(L.gg_peer_handle_bytes())
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(hb)
        check(L.gg_peer_alloc(int(nbytes), ctypes.byref(ptr), handle), 'gg_peer_alloc')
        self._owned.append(ptr.value)
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.device)
        every = torch.empty(self.world * hb, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine, group=self.group)   # also orders the allocation across ranks
        every = every.cpu().numpy().tobytes()
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(ptr.value)
                continue
            q = ctypes.c_void_p()
            check(L.gg_peer_open(every[r * hb:(r + 1) * hb], ctypes.byref(q)), 'gg_peer_open')
            self._opened.append(q.value)
            ptrs.append(q.value)
        return PeerBuffer(ptrs, int(nbytes), self.rank, self.device)

    def get(self, key, nbytes):
        buf = self._bufs.get(key)
        if buf is None or buf.nbytes < nbytes:
            buf = self._bufs[key] = self._alloc(max(int(nbytes), 256))
        return buf

    def barrier(self):
        """Stream-ordered: kernels enqueued after it see every store the peers enqueued before theirs."""
        check(lib().gg_peer_barrier(self._flag_arr, self.world, self.rank,
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'gg_peer_barrier')

    def close(self):
        torch.cuda.synchronize()
        L = lib()
        for q in self._opened:
            L.gg_peer_close(ctypes.c_void_p(q))
        if dist.is_initialized():
            dist.barrier(group=self.group)   # nobody frees while a peer still maps the buffer
        for q in self._owned:
            L.gg_peer_free(ctypes.c_void_p(q))
        self._opened, self._owned, self._bufs = [], [], {}
        PeerPool._shared = {k: v for k, v in PeerPool._shared.items() if v is not self}


# return leg of the sliced exchange: fused (aggregation epilogue stores rows to their owners) | push (bulk, contiguous) |
# auto (fused while a row piece fills an NVLink packet, fs * 4 >= 128 bytes, i.e. up to 4 ranks at 128 columns; push beyond)
PEER_RETURN = os.environ.get('GG_PEER_RETURN', 'auto')
PEER_PIPE = int(os.environ.get('GG_PEER_PIPE', '1'))   # owner groups of the pipelined push; 1 = one aggregation, one push (the default:
# measured at 8 GPUs, 2 / 4 groups cost more in extra launches, kernel tails and host time — 3.24 / 3.49 ms per step — than the
# hidden 0.2 ms leg saves against 2.93 ms)
_TRACE = os.environ.get('GG_PEER_TRACE', '0') == '1'   # CUDA-event timing of the exchange phases (diagnostics)
_trace_events = []


def _mark(name):
    if _TRACE:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        _trace_events.append((name, ev))


def trace_report():
    """Mean milliseconds between consecutive marks, keyed by 'from->to' (call after a synchronize)."""
    acc = {}
    for (a, ea), (b, eb) in zip(_trace_events, _trace_events[1:]):
        acc.setdefault(f'{a}->{b}', []).append(ea.elapsed_time(eb))
    _trace_events.clear()
    return {k: (round(sum(v) / len(v), 4), len(v)) for k, v in acc.items()}


def sliced_width(f, world):
    """Slice width of the feature-sliced exchange, or 0 when f does not split into 16-byte aligned slices
    of at most 128 columns (the sub-warp-group kernel's range)."""
    if world < 1 or f % (4 * world) != 0 or f // world > 128:
        return 0
    return f // world


def _peer_scatter_cols(local, ptrs, row_base, rank):
    local, ld = ops._rows(local, 'local')
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    check(lib().gg_peer_scatter_cols_f32(ctypes.c_void_p(local.data_ptr()), ld, local.size(0), local.size(1), arr,
                                         len(ptrs), int(rank), int(row_base),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
          'gg_peer_scatter_cols_f32')


def _to_slices(playout, local):
    """Forward leg of the exchange: this rank's rows [rows, F] -> its column slice of ALL rows [N, F/P].
    Peer form: a view of the pool's slice buffer.  It is valid only until ANY rank's next ``_to_slices`` of the same
    shape, i.e. until this rank's next barrier (the ``_from_slices`` that consumes it): whoever needs the slice later
    must clone it before that barrier."""
    part, group = playout.part, playout.group
    P, per, n = part.world, part.per, part.n
    rows, f = local.shape
    fs = sliced_width(f, P)
    if not fs:
        raise ValueError(f'feature-sliced exchange needs f % (4*world) == 0 and f/world <= 128 (f={f}, world={P})')
    if playout.exchange == 'sliced':
        pool = playout.pool
        xs = pool.get(('slice', n, fs), n * fs * 4)
        _mark('enter')
        _peer_scatter_cols(local, xs.ptrs, part.lo, part.rank)
        _mark('scattered')
        pool.barrier()                                       # my slice is complete
        _mark('barrier1')
        return xs.view(n, fs)
    send = local.new_zeros((P, per, fs))
    send[:, :rows] = local.reshape(rows, P, fs).permute(1, 0, 2)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    return recv.view(P * per, fs)[:n].contiguous()


def _from_slices(playout, layout, w, x_slice, rows, reduce, bias, self_scale=0.0, rank1=None):
    """Aggregate this rank's slice over the whole graph and return the finished rows of THIS rank [rows, F]
    (return leg of the exchange).  ``bias`` / the rank-1 vectors are full-width; their slice is cut here."""
    part, group = playout.part, playout.group
    P, per, n = part.world, part.per, part.n
    fs = x_slice.size(1)
    f = fs * P
    cut = lambda v: v[part.rank * fs:(part.rank + 1) * fs].contiguous() if v is not None else None
    b = cut(bias)
    if rank1 is not None:
        rank1 = (rank1[0], cut(rank1[1]), rank1[2], cut(rank1[3]))
    x_self = x_slice if self_scale != 0.0 else None
    if playout.exchange == 'sliced':
        pool = playout.pool
        ob = pool.get(('rows', per, f), per * f * 4)
        if PEER_RETURN == 'fused' or (PEER_RETURN == 'auto' and fs >= 32):
            # rows stored into their owners' blocks by the aggregation kernel's epilogue (one kernel, fs*4-byte stores)
            ops.spmm(layout, x_slice, w, reduce, x_self, self_scale, b, rank1=rank1,
                     out_peers=ops.PeerRows([q + part.rank * fs * 4 for q in ob.ptrs], per, f))
            _mark('aggregated')
            pool.barrier()                                   # every slice of my rows has landed
            _mark('barrier2')
            res = ob.view(per, f)[:rows].clone()             # the block is reused by the next exchange
            _mark('cloned')
            return res
        # default: aggregate into a local slice, push it to the owners in contiguous blocks (full NVLink packets), then
        # assemble the owner's rows from the P received slices — the assembly doubles as the copy out of the reused block.
        # The rows are aggregated in PEER_PIPE groups of owners; a group's push runs on a second stream under the next
        # group's aggregation (a 0.22 ms leg against a 0.63 ms kernel at 8 GPUs).
        arr = (ctypes.c_void_p * P)(*ob.ptrs)
        out_slice = torch.empty((n, fs), dtype=torch.float32, device=x_slice.device)
        groups = playout.owner_groups(layout, min(PEER_PIPE, P))
        main = torch.cuda.current_stream()
        side = pool.side_stream if len(groups) > 1 else main
        for g in groups:
            lo, hi = g['rows']
            ops.spmm(g['csr'], x_slice, playout.group_weights(g, w), reduce,
                     x_self[lo:hi] if x_self is not None else None, self_scale, b,
                     rank1=(rank1[0][lo:hi], rank1[1], rank1[2][lo:hi], rank1[3]) if rank1 is not None else None,
                     out=out_slice[lo:hi])
            if side is not main:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
            check(lib().gg_peer_push_rows_f32(ctypes.c_void_p(out_slice.data_ptr()), n, fs, per, P, part.rank, g['owners'][0],
                                              g['owners'][1], arr, ctypes.c_void_p(side.cuda_stream)), 'gg_peer_push_rows_f32')
        _mark('aggregated')
        if side is not main:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)
            out_slice.record_stream(side)
        _mark('pushed')
        pool.barrier()                                       # every slice of my rows has landed
        _mark('barrier2')
        res = torch.empty((rows, f), dtype=torch.float32, device=x_slice.device)
        check(lib().gg_peer_gather_slices_f32(ctypes.c_void_p(ob.ptrs[part.rank]), per, rows, fs, P,
                                              ctypes.c_void_p(res.data_ptr()), f,
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              'gg_peer_gather_slices_f32')
        _mark('cloned')
        return res
    out_slice = ops.spmm(layout, x_slice, w, reduce, x_self, self_scale, b, rank1=rank1)
    send = out_slice.new_zeros((P * per, fs))
    send[:n] = out_slice
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    return recv.view(P, per, fs).permute(1, 0, 2).reshape(per, f)[:rows].contiguous()


def _sliced_spmm(playout, layout, w, local, reduce, bias, self_scale=0.0):
    """out_r = (A X)[rows of rank r]: column slices out (transposition), full-graph aggregation of this
    rank's slice, finished rows back to their owners."""
    return _from_slices(playout, layout, w, _to_slices(playout, local), local.size(0), reduce, bias, self_scale)


class _DistAggregateSliced(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_local, playout, kind, bias, self_scale):
        w_fwd, _ = playout.weights(kind)
        out = _sliced_spmm(playout, playout.csr, w_fwd, h_local.contiguous(),
                           ops.MEAN if kind == 'mean' else ops.SUM, bias, self_scale)
        ctx.playout, ctx.kind, ctx.has_bias, ctx.self_scale = playout, kind, bias is not None, self_scale
        return out

    @staticmethod
    def backward(ctx, g_local):
        playout = ctx.playout
        gh = gb = None
        g_local = g_local.contiguous()
        if ctx.needs_input_grad[0]:
            _, w_bwd = playout.weights(ctx.kind)
            gh = _sliced_spmm(playout, playout.csc, w_bwd, g_local, ops.SUM, None, ctx.self_scale)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            gb = ops.colsum(g_local)
        return gh, None, None, gb, None


def dist_aggregate(h_local, playout, kind='sum', bias=None, self_scale=0.0):
    """Row-partitioned ``functional.aggregate``: out_r = (A H + self_scale * H)[rows of rank r] (+ bias).
    A feature width the sliced exchange cannot cut into aligned slices falls back to the all-to-all form."""
    self_scale = float(self_scale)
    if playout.sliced:
        if playout.exchange == 'sliced' and not sliced_width(h_local.size(1), playout.part.world):
            raise ValueError(f'feature-sliced exchange: {h_local.size(1)} columns do not split into 16-byte aligned '
                             f'slices over {playout.part.world} ranks; build the layout with exchange="allgather"')
        return _DistAggregateSliced.apply(h_local, playout, kind, bias, self_scale)
    if playout.pipelined:
        if kind == 'mean':
            raise NotImplementedError('pipelined halo exchange: mean aggregation (use allgather or sliced)')
        return _DistAggregatePipelined.apply(h_local, playout, kind, bias, self_scale)
    return _DistAggregate.apply(h_local, playout, kind, bias, self_scale)


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


class RowPartitionedSAGE(torch.nn.Module):
    """``sageconv`` on a row partition (parameters / init of models.layer._SAGEConvLayer): the mean over the
    neighbours is the exchanged aggregation (F_in wide), both linear maps are row-local."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .models.layer import _SAGEConvLayer
        self.model = _SAGEConvLayer(dim_in, dim_out, bias=bias)

    def forward(self, x_local, playout):
        m = self.model
        mean = dist_aggregate(x_local, playout, 'mean')
        return F_.seg_linear([mean, x_local], [m.lin_l.weight, m.lin_r.weight], [(0, 0, False), (1, 1, False)],
                             None, m.lin_l.bias, w_trans=True)


class RowPartitionedGIN(torch.nn.Module):
    """``ginconv`` on a row partition: z = (1 + eps) x + sum_j x_j exchanged, the MLP row-local."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .models.layer import _GINConvLayer
        self.model = _GINConvLayer(torch.nn.Sequential(torch.nn.Linear(dim_in, dim_out), torch.nn.ReLU(),
                                                       torch.nn.Linear(dim_out, dim_out)))

    def forward(self, x_local, playout):
        from .contrib.layer.idconv import _mlp
        z = dist_aggregate(x_local, playout, 'sum', None, 1.0 + float(self.model.initial_eps))
        return _mlp(self.model.nn, z)

class _DistGatSliced(torch.autograd.Function):
    """GAT aggregation (heads = 1, ref: idconv.py:299-342) on a row partition with the feature-sliced exchange.

    The per-node logit halves are computed where the rows live and all-gathered (2 floats per node); the
    softmax weights alpha and their backward are light per-slot passes that every rank runs on the full
    layout; the two heavy passes (sum_j alpha_ij H_j and its transpose) are sliced aggregations; the edge
    gradient dalpha_e = <g_i, H_j> is a sliced SDDMM whose per-rank shares are all-reduced."""

    @staticmethod
    def forward(ctx, h_local, att, bias, playout, slope):
        part, group = playout.part, playout.group
        h_local = h_local.contiguous()
        c = h_local.size(1)
        att_row = att.contiguous().view(1, 2 * c)
        a_tgt_l, a_src_l = ops.gat_scores(h_local, att_row)
        a_tgt = all_gather_rows(a_tgt_l.view(-1, 1), part, group).reshape(-1).contiguous()
        a_src = all_gather_rows(a_src_l.view(-1, 1), part, group).reshape(-1).contiguous()
        alpha = ops.gat_alpha(playout.csr, a_tgt, a_src, slope)
        h_slice = _to_slices(playout, h_local)
        # the slice lives in the pool's (reused) peer buffer: keep a private copy for the backward NOW, between the two
        # barriers of this exchange — a peer that has left barrier 2 may start its next _to_slices (its backward's
        # g slice has the same shape) and overwrite the buffer before this rank's host has enqueued a later clone
        h_keep = h_slice.clone() if playout.exchange == 'sliced' else h_slice
        out = _from_slices(playout, playout.csr, alpha, h_slice, h_local.size(0), ops.SUM, bias)
        ctx.playout, ctx.slope, ctx.has_bias = playout, slope, bias is not None
        ctx.save_for_backward(h_local, att_row, alpha, a_tgt, a_src, h_keep)
        return out

    @staticmethod
    def backward(ctx, g_local):
        playout = ctx.playout
        part, group = playout.part, playout.group
        h_local, att_row, alpha, a_tgt, a_src, h_slice = ctx.saved_tensors
        c = h_local.size(1)
        g_local = g_local.contiguous()
        g_slice = _to_slices(playout, g_local)
        dalpha = ops.gat_sddmm_slice(playout.csr, h_slice, g_slice)
        dist.all_reduce(dalpha, op=dist.ReduceOp.SUM, group=group)
        dz, da_tgt = ops.gat_dz(playout.csr, a_tgt, a_src, alpha, dalpha, ctx.slope)
        alpha_t, da_src = ops.gat_csc_gather(playout.csc, playout.csc2csr, alpha, dz)
        # dH = A_alpha^T g + da_src att_src + da_tgt att_tgt   (att = [tgt half | src half])
        dh = _from_slices(playout, playout.csc, alpha_t, g_slice, h_local.size(0), ops.SUM, None,
                          rank1=(da_src, att_row[0, c:], da_tgt, att_row[0, :c]))
        datt = ops.gat_att_grad(h_local, da_tgt[part.lo:part.hi], da_src[part.lo:part.hi])   # this rank's rows
        gb = ops.colsum(g_local) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dh, datt.view(1, 1, 2 * c), gb, None, None


class RowPartitionedGAT(torch.nn.Module):
    """``gatconv`` (heads = 1) on a row partition with the feature-sliced exchange: parameters / init of
    models.layer._GATConvLayer; datt / dW / dbias are per-rank partial sums (``allreduce_grads``)."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .models.layer import _GATConvLayer
        self.model = _GATConvLayer(dim_in, dim_out, heads=1, bias=bias)

    def forward(self, x_local, playout):
        if not playout.sliced:
            raise NotImplementedError('row-partitioned gatconv runs on the feature-sliced exchange '
                                      '(exchange="sliced" or "sliced_nccl")')
        m = self.model
        h = F_.seg_linear([x_local], [m.weight], [(0, 0, False)])
        return _DistGatSliced.apply(h, m.att, m.bias, playout, float(m.negative_slope))


class RowPartitionedGATID(torch.nn.Module):
    """``gatidconv`` (ref: idconv.py:266-347, heads = 1) on a row partition: the heterogeneous transform X W (+ X W_id on the
    centre rows, ``node_id_index`` cut to this rank's rows) is row-local; attention + aggregation run on the feature-sliced
    exchange exactly as for ``gatconv``."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .contrib.layer.idconv import GATIDConvLayer
        self.model = GATIDConvLayer(dim_in, dim_out, heads=1, bias=bias)

    def forward(self, x_local, playout, node_id_index):
        from .graph import IdIndex
        if not playout.sliced:
            raise NotImplementedError('row-partitioned gatidconv runs on the feature-sliced exchange '
                                      '(exchange="sliced" or "sliced_nccl")')
        m = self.model
        h = F_.id_linear(x_local, m.weight, m.weight_id, IdIndex(_local_ids(node_id_index, playout.part), x_local.size(0)))
        return _DistGatSliced.apply(h, m.att, m.bias, playout, float(m.negative_slope))


class RowPartitionedIDConv(torch.nn.Module):
    """``idconv`` (GeneralIDConvLayer, ref: idconv.py:16-101) on a row partition: ID transform row-local, aggregation kind
    from ``cfg.gnn.agg`` / ``cfg.gnn.normalize_adj`` as the single-GPU layer reads them at construction."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .contrib.layer.idconv import GeneralIDConvLayer
        self.model = GeneralIDConvLayer(dim_in, dim_out, bias=bias)

    @staticmethod
    def policy_and_kind():
        from .config import cfg
        if cfg.gnn.normalize_adj:
            if cfg.gnn.agg == 'mean':
                raise NotImplementedError('row-partitioned idconv: normalize_adj with agg="mean" (single-GPU only)')
            return ops.LOOPS_ADD_REMAINING, 'gcn_src'
        return ops.LOOPS_KEEP, ('mean' if cfg.gnn.agg == 'mean' else 'sum')

    def forward(self, x_local, playout, node_id_index):
        from .graph import IdIndex
        m = self.model
        h = F_.id_linear(x_local, m.weight, m.weight_id, IdIndex(_local_ids(node_id_index, playout.part), x_local.size(0)))
        return dist_aggregate(h, playout, self.policy_and_kind()[1], m.bias)


class RowPartitionedGCNID(torch.nn.Module):
    """``gcnidconv`` (ID-GNN's GCN layer, ref: idconv.py:104-189) on a row partition: the heterogeneous transform
    X W (+ X W_id on the centre rows) is row-local — ``node_id_index`` is cut to this rank's rows — and the
    normalised aggregation (degree over edge_index[0]) is the exchanged step."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .contrib.layer.idconv import GCNIDConvLayer
        self.model = GCNIDConvLayer(dim_in, dim_out, bias=bias)

    def forward(self, x_local, playout, node_id_index):
        from .graph import IdIndex
        ids = _local_ids(node_id_index, playout.part)
        h = F_.id_linear(x_local, self.model.weight, self.model.weight_id, IdIndex(ids, x_local.size(0)))
        return dist_aggregate(h, playout, 'gcn_src', self.model.bias)


def _local_ids(node_id_index, part):
    return node_id_index[(node_id_index >= part.lo) & (node_id_index < part.hi)] - part.lo


class RowPartitionedSAGEID(torch.nn.Module):
    """``sageidconv`` (ref: idconv.py:192-263, ``concat=True`` as GraphGym builds it, idconv.py:410) on a row
    partition: the neighbour mean is the exchanged aggregation, [x | mean] (W, and W_id on the centres) is row-local."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .contrib.layer.idconv import SAGEIDConvLayer
        self.model = SAGEIDConvLayer(dim_in, dim_out, bias=bias, concat=True)

    def forward(self, x_local, playout, node_id_index):
        from .graph import IdIndex
        m, k = self.model, self.model.in_channels
        info = IdIndex(_local_ids(node_id_index, playout.part), x_local.size(0))
        mean = dist_aggregate(x_local, playout, 'mean')
        return F_.seg_linear([x_local, mean], [m.weight, m.weight_id],
                             [(0, 0, False), (1, 0, False), (0, 1, True), (1, 1, True)], info, m.bias,
                             w_rows=((0, k), (k, 2 * k), (0, k), (k, 2 * k)))


class RowPartitionedGINID(torch.nn.Module):
    """``ginidconv`` (ref: idconv.py:350-382, MLPs of idconv.py:432-436) on a row partition: z = (1 + eps) x + sum_j x_j
    on the loop-free graph is the exchanged step; nn(z) everywhere and nn_id(z[id]) on this rank's centres are row-local."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        from .contrib.layer.idconv import GINIDConvLayer

        def mlp():
            return torch.nn.Sequential(torch.nn.Linear(dim_in, dim_out), torch.nn.ReLU(), torch.nn.Linear(dim_out, dim_out))

        self.model = GINIDConvLayer(mlp(), mlp())

    def forward(self, x_local, playout, node_id_index):
        from .contrib.layer.idconv import _mlp
        m = self.model
        ids = _local_ids(node_id_index, playout.part).contiguous().long()
        z = dist_aggregate(x_local, playout, 'sum', None, 1.0 + float(m.initial_eps))
        out = _mlp(m.nn, z)
        return F_.scatter_add_rows(out, ids, _mlp(m.nn_id, F_.gather_rows(z, ids)))


ROW_PARTITIONED = {'gcnconv': (RowPartitionedGCN, ops.LOOPS_ADD_REMAINING, 'gcn_tgt'),
                   'sageidconv': (RowPartitionedSAGEID, ops.LOOPS_KEEP, 'mean'),
                   'ginidconv': (RowPartitionedGINID, ops.LOOPS_REMOVE, 'sum'),
                   'gcnidconv': (RowPartitionedGCNID, ops.LOOPS_ADD_REMAINING, 'gcn_src'),
                   'sageconv': (RowPartitionedSAGE, ops.LOOPS_KEEP, 'mean'),
                   'ginconv': (RowPartitionedGIN, ops.LOOPS_KEEP, 'sum'),
                   'gatconv': (RowPartitionedGAT, ops.LOOPS_REMOVE_ADD, 'sum'),
                   'gatidconv': (RowPartitionedGATID, ops.LOOPS_REMOVE_ADD, 'sum'),
                   'idconv': (RowPartitionedIDConv, ops.LOOPS_KEEP, 'sum')}   # idconv: cfg.gnn defaults (add, no normalisation)
