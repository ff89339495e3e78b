"""CPU restatement of the degree-sorted sliced-ELL re-layout (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

The reference has no such structure (its aggregation is a COO gather + scatter, ref: graphgym/contrib/layer/idconv.py:89-92,
177-180); this restates OUR layout definition (include/gg_b200.h "Degree-sorted sliced-ELL") in numpy so that the integer
build kernels (csrc/spmm_sell.cu) can be checked bit for bit, and offers the aggregation over that layout in fp64.
"""
import numpy as np

ROWS = 8
NO_ROW = np.iinfo(np.int32).min


def build(rowptr, nbr, seg):
    """-> dict(chunk_ptr, idx, slot_of, vdst, hub_rows, hub_pptr, vrows, chunks, units, hubs, partial_rows); arrays are
    trimmed to the entries in use (capacity padding of the device arrays is not part of the definition)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    nbr = np.asarray(nbr, dtype=np.int64)
    n = rowptr.size - 1
    deg = rowptr[1:] - rowptr[:-1]
    cnt = np.where(deg <= seg, 1, (deg + seg - 1) // seg)
    vstart = np.concatenate([[0], np.cumsum(cnt)])
    split = cnt > 1
    hstart = np.concatenate([[0], np.cumsum(split)])
    V = int(vstart[-1])
    vlen = np.zeros(V, dtype=np.int64)
    voff = np.zeros(V, dtype=np.int64)
    vdst_tmp = np.zeros(V, dtype=np.int64)
    hub_rows, hub_pptr = [], []
    for r in range(n):
        p0 = int(vstart[r] - (r - hstart[r]))
        if split[r]:
            hub_rows.append(r)
            hub_pptr.append(p0)
        for s in range(int(cnt[r])):
            v = int(vstart[r]) + s
            vlen[v] = min(seg, deg[r] - s * seg)
            voff[v] = rowptr[r] + s * seg
            vdst_tmp[v] = r if cnt[r] == 1 else -(p0 + s) - 1
    partial_rows = int(vstart[n] - (n - hstart[n]))
    hub_pptr.append(partial_rows)
    order = np.argsort(-vlen, kind='stable')                 # descending length, ties in virtual-row order
    chunks = (V + ROWS - 1) // ROWS
    chunk_ptr = np.zeros(chunks + 1, dtype=np.int64)
    for c in range(chunks):
        first = vlen[order[c * ROWS]]
        chunk_ptr[c + 1] = chunk_ptr[c] + ((first + 3) // 4) * ROWS
    units = int(chunk_ptr[-1])
    idx = np.full(4 * units, -1, dtype=np.int64)
    slot_of = np.full(4 * units, -1, dtype=np.int64)
    vdst = np.full(chunks * ROWS, NO_ROW, dtype=np.int64)
    for c in range(chunks):
        nk = (chunk_ptr[c + 1] - chunk_ptr[c]) // ROWS
        for q in range(ROWS):
            i = c * ROWS + q
            if i >= V:
                continue
            v = order[i]
            vdst[i] = vdst_tmp[v]
            for k in range(int(vlen[v])):
                d = (chunk_ptr[c] + (k // 4) * ROWS + q) * 4 + k % 4
                idx[d] = nbr[voff[v] + k]
                slot_of[d] = voff[v] + k
            assert vlen[v] <= nk * 4
    return dict(chunk_ptr=chunk_ptr, idx=idx, slot_of=slot_of, vdst=vdst, hub_rows=np.asarray(hub_rows, dtype=np.int64),
                hub_pptr=np.asarray(hub_pptr, dtype=np.int64), vrows=V, chunks=chunks, units=units, hubs=len(hub_rows),
                partial_rows=partial_rows)
