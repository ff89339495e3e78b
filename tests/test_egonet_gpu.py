"""ID-GNN Full ego-net extraction: GPU batched frontier BFS vs (a) the golden output of the reference's
own ego_nets (canonical form: member sets + induced edge sets per centre), (b) the oracle, bit-exact."""
import os

import numpy as np
import pytest
import torch

from graphgym_b200.models.transform import ego_nets, ego_nets_batch
from oracle import egonets as oego
from util import random_graph

pytestmark = pytest.mark.gpu


def test_golden_reference_ego_nets(cuda, golden_dir):
    d = np.load(os.path.join(golden_dir, 'egonets.npz'))
    for tag in d['names']:
        n, r = int(d[f'{tag}/n']), int(d[f'{tag}/radius'])
        ei = torch.from_numpy(d[f'{tag}/edge_index']).to(cuda)
        res = ego_nets_batch(ei, n, r)
        assert res['num_nodes'] == int(d[f'{tag}/num_nodes_out']), tag
        assert res['edge_index'].size(1) == 2 * int(d[f'{tag}/num_edges_out']), tag
        assert np.array_equal(res['node_id_index'].cpu().numpy(), d[f'{tag}/node_id_index'])
        ego_ptr = res['ego_ptr'].cpu().numpy().astype(np.int64)
        mem, edges = oego.canonical(n, res['edge_index'].cpu().numpy(), res['orig_id'].cpu().numpy(), ego_ptr)
        mp, ep = d[f'{tag}/member_ptr'], d[f'{tag}/edge_ptr']
        for c in range(n):
            assert mem[c] == d[f'{tag}/members'][mp[c]:mp[c + 1]].tolist(), (tag, c)
            assert edges[c] == [tuple(e) for e in d[f'{tag}/edges'][ep[c]:ep[c + 1]].tolist()], (tag, c)


def exact(res, want):
    assert res['num_nodes'] == want['num_nodes']
    assert np.array_equal(res['orig_id'].cpu().numpy(), want['orig_id'])
    assert np.array_equal(res['edge_index'].cpu().numpy(), want['edge_index'])
    assert np.array_equal(res['node_id_index'].cpu().numpy(), want['node_id_index'])


@pytest.mark.parametrize('radius', [0, 1, 2, 3, 4, 5])
def test_bit_exact_vs_oracle_single_graph(cuda, radius):
    n = 200
    ei = random_graph(radius, n, 300, symmetric=True)
    ei = ei[:, ei[0] != ei[1]]
    ei = torch.unique(ei, dim=1)
    ei = ei[:, torch.randperm(ei.size(1), generator=torch.Generator().manual_seed(1))]
    exact(ego_nets_batch(ei.to(cuda), n, radius), oego.ego_nets(ei.numpy(), n, radius))


def test_bit_exact_batch_of_graphs(cuda):
    sizes = [64, 1, 64, 33, 100, 2]
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    parts = []
    for g, s in enumerate(sizes):
        if s > 1:
            e = random_graph(g + 5, s, 2 * s, symmetric=True)
            e = torch.unique(e[:, e[0] != e[1]], dim=1)
            parts.append(e + int(ptr[g]))
    ei = torch.cat(parts, dim=1)
    for radius in (1, 2, 3):
        res = ego_nets_batch(ei.to(cuda), int(ptr[-1]), radius, torch.from_numpy(ptr))
        want = oego.ego_nets_batch(ei.numpy(), ptr, radius)
        exact(res, want)
        assert np.array_equal(res['out_node_ptr'].cpu().numpy(), want['out_node_ptr'])


def test_medium_graph_multiword_bitmaps(cuda):
    """2708 nodes (Cora-sized): 85 bitmap words per warp, ego-nets of thousands of members."""
    n = 2708
    ei = random_graph(0, n, 5278, symmetric=True)
    ei = torch.unique(ei[:, ei[0] != ei[1]], dim=1)
    exact(ego_nets_batch(ei.to(cuda), n, 2), oego.ego_nets(ei.numpy(), n, 2))


def test_reference_style_transform_mutates_graph(cuda):
    class G:
        pass
    n = 50
    ei = random_graph(2, n, 80, symmetric=True)
    ei = torch.unique(ei[:, ei[0] != ei[1]], dim=1)
    g = G()
    g.edge_index, g.num_nodes = ei.to(cuda), n
    g.node_feature = torch.arange(n, dtype=torch.float32, device=cuda).view(-1, 1)
    out = ego_nets(g, radius=2)
    want = oego.ego_nets(ei.numpy(), n, 2)
    assert out is g and g.num_nodes == want['num_nodes']
    assert g.node_id_index.cpu().tolist() == list(range(n))
    assert np.array_equal(g.node_feature.flatten().cpu().numpy(), want['orig_id'].astype(np.float32))
