"""On-device batch collation (csrc/collate.cu, graphgym_b200/loader.py) against the host-side
concatenation DeepSNAP's Batch.collate performs (ref: graphgym/loader.py:245-250): bit-exact for the index arrays,
exact for the float blocks; Preprocess's column concat (ref: graphgym/models/feature_augment.py:329-333)."""
import pytest
import torch

from graphgym_b200 import loader
from util import random_graph

pytestmark = pytest.mark.gpu


def host_collate(graphs):
    """torch.cat with the running node count added to every *_index tensor (DeepSNAP Batch semantics)."""
    out, off = {}, 0
    offs = []
    for g in graphs:
        offs.append(off)
        off += g['node_feature'].size(0)
    for k in graphs[0]:
        ts = [g[k] for g in graphs]
        if 'index' in k:
            out[k] = torch.cat([t + o for t, o in zip(ts, offs)], dim=-1)
        elif ts[0].dtype == torch.int64:
            out[k] = torch.cat(ts, 0)
        else:
            out[k] = torch.cat([t.float().view(t.size(0), -1) for t in ts], 0)
    out['batch'] = torch.cat([torch.full((g['node_feature'].size(0),), i, dtype=torch.int64) for i, g in enumerate(graphs)])
    return out


def make_graphs(sizes, f=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    graphs = []
    for i, n in enumerate(sizes):
        ei = random_graph(seed + i, n, 3 * n) if n > 1 else torch.zeros((2, 0), dtype=torch.int64)
        graphs.append(dict(node_feature=torch.randn(n, f, generator=g), edge_index=ei,
                           node_label=torch.randint(0, 7, (n,), generator=g), node_id_index=torch.arange(0, n, 2),
                           node_identity=torch.rand(n, 3, generator=g),
                           node_degree=torch.randint(0, 4, (n,), generator=g).to(torch.uint8)))
    return graphs


@pytest.mark.parametrize('sizes', [[64] * 16, [1, 7, 300, 2, 50], [5], [33] * 128])
def test_collate_matches_host_concatenation(cuda, sizes):
    graphs = make_graphs(sizes)
    want = host_collate(graphs)
    got = loader.collate([loader.GraphData(**{k: v.to(cuda) for k, v in g.items()}) for g in graphs])
    assert got.num_graphs == len(sizes) and got.num_nodes == sum(sizes)
    for k, w in want.items():
        assert got[k].shape == w.shape, k
        assert torch.equal(got[k].cpu(), w), k
    assert got['node_ptr'].tolist()[-1] == sum(sizes)


def test_preprocess_concat_columns(cuda):
    from graphgym_b200.config import cfg, reset_cfg
    from graphgym_b200.models.feature_augment import Preprocess
    reset_cfg()
    cfg.dataset.augment_feature = ['node_identity', 'node_degree']
    cfg.dataset.augment_feature_dims = [3, 1]
    graphs = make_graphs([40, 24])
    batch = loader.collate([loader.GraphData(**{k: v.to(cuda) for k, v in g.items()}) for g in graphs])
    want = torch.cat([batch[k].float().view(batch.num_nodes, -1) for k in ('node_identity', 'node_degree', 'node_feature')], 1)
    pre = Preprocess(5)
    assert pre.dim_out == 9
    out = pre(batch)
    assert out is batch and torch.equal(batch.node_feature, want)
    reset_cfg()


def test_balanced_binning_matches_reference_numpy(cuda, golden_dir):
    """Cfg-A labels: clustering coefficient (cycle kernel) -> balanced 10-way bins == the numpy pipeline of
    feature_augment.py:134-143,219-231 run by tests/golden/make_golden.py over graphs [0:16] of scalefree.pkl."""
    import os
    import numpy as np
    from graphgym_b200.contrib.transform import binning
    from graphgym_b200.contrib.transform.clustering import clustering_coefficient
    d = np.load(os.path.join(golden_dir, 'scalefree16.npz'))
    ei = torch.from_numpy(d['edge_index']).to(cuda)
    gp = torch.from_numpy(d['graph_ptr']).to(cuda)
    n = int(d['graph_ptr'][-1])
    clus = clustering_coefficient(ei, n, graph_ptr=gp.int())
    assert np.abs(clus.cpu().numpy() - d['clustering']).max() < 1e-12
    labels, edges = binning.balanced_labels(torch.from_numpy(d['clustering']).to(cuda), 10)
    assert np.array_equal(edges, d['bin_edges'])
    assert np.array_equal(labels.cpu().numpy(), d['label'])
    # float64 argsort: stable ascending, incl. negative values, zeros and ties
    g = torch.Generator().manual_seed(0)
    v = torch.cat([torch.randn(5000, generator=g, dtype=torch.float64), torch.zeros(100, dtype=torch.float64),
                   -torch.zeros(3, dtype=torch.float64), torch.randn(50, generator=g, dtype=torch.float64).repeat(20)])
    order = binning.argsort_f64(v.to(cuda)).cpu().long()
    want = torch.from_numpy(np.argsort(np.where(v.numpy() == 0, 0.0, v.numpy()), kind='stable'))
    assert torch.equal(v[order], v[want])
