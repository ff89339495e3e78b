"""ID-GNN Full ego-net expansion on the GPU (ref: graphgym/models/transform.py:11-38).

``ego_nets(graph, radius)`` mirrors the reference transform: it mutates a graph-like object, replacing
its topology by the union of the radius-hop ego-nets of all its nodes (centre copies first, keeping ids
0..n-1, then every ego's other members with fresh consecutive ids) and setting
``graph.node_id_index = arange(n)``.  Node-level tensors are carried over by original id, as
``nx.relabel_nodes(copy=True)`` carries node attributes.  ``ego_nets_batch`` does the same for a
block-diagonal batch of graphs in two kernel launches (csrc/egonet.cu).
"""
import torch

from graphgym_b200 import ops
from graphgym_b200.ops import _ptr, _stream, check, lib


def ego_nets_batch(edge_index, num_nodes, radius, graph_ptr=None):
    """-> dict(edge_index [2,E_out] int64, orig_id [N_out] int64, node_id_index [num_nodes] int64,
    out_node_ptr [G+1] int64, ego_ptr [n+1], edge_ptr [n+1], num_nodes N_out).

    ``edge_index``: symmetric directed edge list of the batch (CUDA, int64); ``graph_ptr``: [G+1] node
    offsets (None = one graph)."""
    ops._need_cuda(edge_index)
    dev = edge_index.device
    n = int(num_nodes)
    if graph_ptr is None:
        graph_ptr = torch.tensor([0, n], dtype=torch.int32, device=dev)
    graph_ptr = graph_ptr.to(device=dev, dtype=torch.int32).contiguous()
    G = graph_ptr.numel() - 1
    sizes = (graph_ptr[1:] - graph_ptr[:-1])
    max_nodes = int(sizes.max().item()) if G > 0 else 1
    graph_of = torch.repeat_interleave(torch.arange(G, dtype=torch.int32, device=dev), sizes.long())
    adj = ops.layout_build(edge_index, n, ops.LOOPS_KEEP, ops.BY_SOURCE)
    L = lib()
    ego_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    edge_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    out_node_ptr = torch.empty(G + 1, dtype=torch.int64, device=dev)
    ws_bytes = int(L.gg_egonet_workspace_bytes(n, G))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(L.gg_egonet_sizes(_ptr(adj.rowptr), _ptr(adj.nbr), n, int(radius), _ptr(graph_ptr), _ptr(graph_of), G,
                            max_nodes, _ptr(ego_ptr), _ptr(edge_ptr), _ptr(out_node_ptr), _ptr(ws), ws_bytes,
                            _stream()), "gg_egonet_sizes")
    total_nodes = int(out_node_ptr[G].item())
    total_edges = int(edge_ptr[n].item()) & 0xffffffff
    orig_id = torch.empty(max(total_nodes, 1), dtype=torch.int64, device=dev)[:total_nodes]
    edge_out = torch.empty((2, max(total_edges, 1)), dtype=torch.int64, device=dev)[:, :total_edges].contiguous()
    check(L.gg_egonet_fill(_ptr(adj.rowptr), _ptr(adj.nbr), n, int(radius), _ptr(graph_ptr), _ptr(graph_of), G,
                           max_nodes, _ptr(ego_ptr), _ptr(edge_ptr), _ptr(out_node_ptr), total_edges,
                           _ptr(orig_id), _ptr(edge_out), _stream()), "gg_egonet_fill")
    # centres of graph g sit at out_node_ptr[g] + arange(n_g)
    node_id_index = (torch.arange(n, device=dev) - graph_ptr[graph_of.long()].long()
                     + out_node_ptr[graph_of.long()])
    return dict(edge_index=edge_out, orig_id=orig_id, node_id_index=node_id_index, out_node_ptr=out_node_ptr,
                ego_ptr=ego_ptr, edge_ptr=edge_ptr, num_nodes=total_nodes)


_NODE_ATTRS = ('node_feature', 'node_label', 'node_identity')


def ego_nets(graph, radius=2):
    """Reference signature (transform.py:11): expand ``graph`` in place.

    ``graph`` needs ``edge_index`` (CUDA int64 [2,E], both directions) and ``num_nodes``; node-level
    tensors named in ``_NODE_ATTRS`` are re-indexed by original id.  Sets ``graph.edge_index``,
    ``graph.num_nodes`` and ``graph.node_id_index = arange(n)``."""
    n = int(graph.num_nodes)
    res = ego_nets_batch(graph.edge_index, n, radius)
    for name in _NODE_ATTRS:
        t = getattr(graph, name, None)
        if torch.is_tensor(t) and t.size(0) == n:
            setattr(graph, name, t.index_select(0, res['orig_id'].to(t.device)))
    graph.edge_index = res['edge_index']
    graph.num_nodes = res['num_nodes']
    graph.node_id_index = torch.arange(n, device=graph.edge_index.device)
    graph.ego_orig_id = res['orig_id']
    return graph
