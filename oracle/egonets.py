"""CPU restatement of ID-GNN Full's ego-net extraction (TEST INFRASTRUCTURE ONLY).

Follows graphgym/models/transform.py:11-38: for every centre i, the induced subgraph on
{v : dist(i, v) <= radius} (radius > 4 => the whole graph); centre copies keep ids 0..n-1, every other
member gets a fresh id from a running counter that starts at n, egos in centre order;
node_id_index = arange(n).  The order of the non-centre members inside one ego is networkx-internal
in the reference; the canonical form used for bit-exact parity is ASCENDING original id.

Input is an undirected simple graph given as a symmetric directed edge list (both directions
present, the DeepSNAP convention).  Output edges are directed, both directions, ordered by (member
ascending, adjacency slot = COO order) within each ego, egos concatenated in centre order.
"""
import numpy as np


def _adj_lists(edge_index, n):
    ei = np.asarray(edge_index, dtype=np.int64)
    order = np.argsort(ei[0], kind='stable')  # adjacency slots keep the COO order (the layout's order)
    src, tgt = ei[0][order], ei[1][order]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=ptr[1:])
    return ptr, tgt


def ego_members(edge_index, n, radius):
    """list over centres of the sorted member arrays (centre included)."""
    ptr, nbr = _adj_lists(edge_index, n)
    out = []
    for c in range(n):
        if radius > 4:
            out.append(np.arange(n, dtype=np.int64))
            continue
        dist = np.full(n, -1, dtype=np.int64)
        dist[c] = 0
        frontier = [c]
        for d in range(radius):
            nxt = []
            for u in frontier:
                for v in nbr[ptr[u]:ptr[u + 1]]:
                    if dist[v] < 0:
                        dist[v] = d + 1
                        nxt.append(int(v))
            frontier = nxt
        out.append(np.nonzero(dist >= 0)[0].astype(np.int64))
    return out


def ego_nets(edge_index, n, radius):
    """-> dict(num_nodes, edge_index[2,E_out], node_id_index[n], orig_id[num_nodes], ego_ptr[n+1],
    edge_ptr[n+1]).  ego_ptr[c]..ego_ptr[c+1] indexes the NON-centre members of ego c in the id space
    shifted by n (new id = n + position)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    members = ego_members(ei, n, radius)
    ptr, nbr = _adj_lists(ei, n)
    sizes = np.array([len(m) - 1 for m in members], dtype=np.int64)
    ego_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes, out=ego_ptr[1:])
    total = n + int(ego_ptr[-1])
    orig = np.empty(total, dtype=np.int64)
    orig[:n] = np.arange(n)
    srcs, tgts = [], []
    edge_ptr = np.zeros(n + 1, dtype=np.int64)
    for c, mem in enumerate(members):
        others = mem[mem != c]
        base = n + ego_ptr[c]
        orig[base:base + len(others)] = others
        new_id = np.full(n, -1, dtype=np.int64)
        new_id[others] = base + np.arange(len(others))
        new_id[c] = c
        cnt = 0
        for u in mem:  # members ascending, each one's adjacency in COO order
            for v in nbr[ptr[u]:ptr[u + 1]]:
                if new_id[v] >= 0:
                    srcs.append(new_id[u])
                    tgts.append(new_id[v])
                    cnt += 1
        edge_ptr[c + 1] = edge_ptr[c] + cnt
    edge_index_out = np.array([srcs, tgts], dtype=np.int64).reshape(2, -1)
    return dict(num_nodes=total, edge_index=edge_index_out, node_id_index=np.arange(n, dtype=np.int64),
                orig_id=orig, ego_ptr=ego_ptr, edge_ptr=edge_ptr)


def canonical(num_centres, edge_index_out, orig_id, ego_ptr):
    """Order-free form for comparing against the reference's networkx output: per centre the sorted
    member set and the sorted induced undirected edge set, in ORIGINAL ids."""
    ei = np.asarray(edge_index_out)
    n = num_centres
    ego_of = np.empty(len(orig_id), dtype=np.int64)
    ego_of[:n] = np.arange(n)
    for c in range(n):
        ego_of[n + ego_ptr[c]: n + ego_ptr[c + 1]] = c
    members = [set([c]) for c in range(n)]
    for v in range(n, len(orig_id)):
        members[ego_of[v]].add(int(orig_id[v]))
    edges = [set() for _ in range(n)]
    for s, t in zip(ei[0], ei[1]):
        c = ego_of[s]
        assert ego_of[t] == c, 'edge crosses ego blocks'
        a, b = int(orig_id[s]), int(orig_id[t])
        edges[c].add((min(a, b), max(a, b)))
    return [sorted(m) for m in members], [sorted(e) for e in edges]


def ego_nets_batch(edge_index, graph_ptr, radius):
    """The reference pipeline on a block-diagonal batch: ego_nets per graph (transform.py is applied per
    graph, loader.py:175-180), then concatenation with node offsets (DeepSNAP Batch.collate)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    gp = np.asarray(graph_ptr, dtype=np.int64)
    edges, origs, ids = [], [], []
    off = 0
    out_ptr = [0]
    for g in range(len(gp) - 1):
        lo, hi = gp[g], gp[g + 1]
        sel = (ei[0] >= lo) & (ei[0] < hi)
        res = ego_nets(ei[:, sel] - lo, int(hi - lo), radius)
        edges.append(res['edge_index'] + off)
        origs.append(res['orig_id'] + lo)
        ids.append(res['node_id_index'] + off)
        off += res['num_nodes']
        out_ptr.append(off)
    return dict(num_nodes=off, edge_index=np.concatenate(edges, axis=1) if edges else np.zeros((2, 0), np.int64),
                orig_id=np.concatenate(origs) if origs else np.zeros(0, np.int64),
                node_id_index=np.concatenate(ids) if ids else np.zeros(0, np.int64),
                out_node_ptr=np.array(out_ptr, dtype=np.int64))
