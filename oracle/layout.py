"""CPU restatement of the graph layout contract (TEST INFRASTRUCTURE ONLY).

The reference has no CSR/CSC build; "bit-exact" is defined as: edit the COO list the way
the reference layer's PyG helper does, then take the STABLE counting sort by the group key
(``numpy.argsort(kind='stable')``), keeping duplicates.
"""
import numpy as np

KEEP, ADD_REMAINING, REMOVE_ADD, REMOVE, ADD = range(5)
BY_TARGET, BY_SOURCE = 0, 1


def edited_edges(edge_index, num_nodes, policy):
    """(src, tgt, eid) of the edited list: kept edges in order, then the appended loops (eid = E+i)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    E = ei.shape[1]
    src, tgt = ei[0], ei[1]
    eid = np.arange(E, dtype=np.int64)
    if policy in (ADD_REMAINING, REMOVE_ADD, REMOVE):
        keep = src != tgt
        src, tgt, eid = src[keep], tgt[keep], eid[keep]
    if policy in (ADD_REMAINING, REMOVE_ADD, ADD):
        loops = np.arange(num_nodes, dtype=np.int64)
        src = np.concatenate([src, loops])
        tgt = np.concatenate([tgt, loops])
        eid = np.concatenate([eid, E + loops])
    return src, tgt, eid


def layout_build(edge_index, num_nodes, policy=KEEP, group_by=BY_TARGET):
    """-> rowptr[N+1] int32, nbr[E'] int32, perm[E'] int32, rowid[E'] int32."""
    src, tgt, eid = edited_edges(edge_index, num_nodes, policy)
    key, other = (tgt, src) if group_by == BY_TARGET else (src, tgt)
    order = np.argsort(key, kind='stable')
    counts = np.bincount(key, minlength=num_nodes)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return (rowptr.astype(np.int32), other[order].astype(np.int32), eid[order].astype(np.int32),
            key[order].astype(np.int32))


def sort_pairs(keys, vals):
    order = np.argsort(np.asarray(keys, dtype=np.uint32), kind='stable')
    return np.asarray(keys)[order], np.asarray(vals)[order]
