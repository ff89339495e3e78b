"""Generate the golden vectors under tests/golden/ by running the REFERENCE'S OWN hot-path files.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports graphgym/contrib/layer/idconv.py, graphgym/contrib/transform/identity.py and
graphgym/models/transform.py from the reference tree (third-party helpers supplied by
oracle/pyg_shim.py), feeds them seeded inputs and stores inputs, parameters, outputs and gradients:

    pooling.npz         global_add/mean/max_pool of graphgym/models/pooling.py (torch_scatter.scatter restated in
                        oracle/pooling.py), with and without the 'ego' centre-row select, fwd + bwd
    idconv_layers.npz   the five registered ID layers (+ cfg variants of `idconv`), fwd + bwd
    identity.npz        compute_identity on K3 / P3 / C4 and two bundled fixture graphs
    egonets.npz         ego_nets (canonicalised: member sets + induced edge sets per centre) on C4 and
                        on bundled fixture graphs, radius 1..3
    contrib_layers.npz  generalconv (3 self_msg modes x agg x normalize_adj), sageinitconv, idconv(normalize_adj, mean), fwd + bwd
    scalefree16.npz     the Cfg-A workload: graphs [0:16] of datasets/scalefree.pkl as one block-diagonal
                        batch (edge lists, node offsets), the reference label of the node task — nx.clustering
                        (feature_augment.py:81-82) — its 10-way balanced binning restated from feature_augment.py:208-231
                        and :134-143 over the 16 graphs, and the size of every graph's radius-3 ego expansion as the
                        reference's own transform.py produces it
The committed .npz files are what the tests read; /root/reference is never touched at test time.
"""
import os
import pickle
import sys
import warnings

import networkx as nx
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyg_shim  # noqa: E402

REF = '/root/reference'
warnings.filterwarnings('ignore')


class _Batch:
    pass


def random_graph(seed, n, e, loops=4, dups=5):
    """Directed, asymmetric, with self loops and duplicate edges (the edge cases of the COO edits)."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    tgt = torch.randint(0, n, (e,), generator=g)
    lp = torch.randint(0, n, (loops,), generator=g)
    src = torch.cat([src, lp, src[:dups]])
    tgt = torch.cat([tgt, lp, tgt[:dups]])
    perm = torch.randperm(src.numel(), generator=g)
    return torch.stack([src[perm], tgt[perm]])


def run_layer(ns, name, cfg_kw, n, fin, fout, seed, out, suffix=''):
    ns.cfg.gnn.agg = cfg_kw.get('agg', 'add')
    ns.cfg.gnn.normalize_adj = cfg_kw.get('normalize_adj', False)
    torch.manual_seed(seed)
    layer = ns.register.layer_dict[name](fin, fout, bias=True)
    with torch.no_grad():
        for p in layer.parameters():  # non-trivial biases
            if p.dim() == 1:
                p.uniform_(-0.5, 0.5)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, fin, generator=g, requires_grad=True)
    ei = random_graph(seed + 2, n, 4 * n)
    ids = torch.randperm(n, generator=g)[: max(2, n // 5)].sort().values
    b = _Batch()
    b.node_feature, b.edge_index, b.node_id_index = x, ei, ids
    y = layer(b).node_feature
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    tag = name + ''.join('_%s-%s' % kv for kv in sorted(cfg_kw.items())) + suffix
    out[tag + '/x'] = x.detach().numpy()
    out[tag + '/edge_index'] = ei.numpy()
    out[tag + '/ids'] = ids.numpy()
    out[tag + '/y'] = y.detach().numpy()
    out[tag + '/gy'] = gy.numpy()
    out[tag + '/gx'] = x.grad.numpy()
    for pn, p in layer.named_parameters():
        out[tag + '/param/' + pn] = p.detach().numpy()
        out[tag + '/grad/' + pn] = p.grad.numpy()
    return tag


def layers(ns):
    out = {}
    tags = []
    for name, kw in [('idconv', {}), ('idconv', {'agg': 'mean'}), ('idconv', {'normalize_adj': True}),
                     ('gcnidconv', {}), ('sageidconv', {}), ('gatidconv', {}), ('ginidconv', {})]:
        tags.append(run_layer(ns, name, kw, n=37, fin=12, fout=16, seed=100 + len(tags), out=out))
    # a wider case on the vector path (F multiple of 4, > 32 lanes)
    for name in ('gcnidconv', 'sageidconv', 'gatidconv', 'ginidconv'):
        tags.append(run_layer(ns, name, {}, n=61, fin=20, fout=136, seed=300 + len(tags), out=out,
                              suffix='_wide'))
    out['tags'] = np.array(tags)
    np.savez_compressed(os.path.join(HERE, 'idconv_layers.npz'), **out)
    print('idconv_layers.npz', len(out), 'arrays')


def sym(edges):
    e = np.array(edges, dtype=np.int64).T
    return np.concatenate([e, e[::-1]], axis=1)


def fixture_graphs(name, idx):
    with open(os.path.join(REF, 'datasets', name + '.pkl'), 'rb') as f:
        graphs = pickle.load(f)
    return [graphs[i] for i in idx]


def nx_to_edge_index(G):
    e = np.array(list(G.edges()), dtype=np.int64).T.reshape(2, -1)
    return np.concatenate([e, e[::-1]], axis=1)


def identity(ns):
    out = {}
    cases = {'K3': (sym([(0, 1), (1, 2), (0, 2)]), 3), 'P3': (sym([(0, 1), (1, 2)]), 3),
             'C4': (sym([(0, 1), (1, 2), (2, 3), (3, 0)]), 4)}
    for i, G in enumerate(fixture_graphs('scalefree', [0, 7])):
        cases['scalefree%d' % i] = (nx_to_edge_index(G), G.number_of_nodes())
    for name, (ei, n) in cases.items():
        k = 4 if n <= 4 else 10
        res = ns.identity.compute_identity(torch.from_numpy(ei), n, k)
        out[name + '/edge_index'] = ei
        out[name + '/n'] = np.int64(n)
        out[name + '/k'] = np.int64(k)
        out[name + '/identity'] = res.numpy()
    out['names'] = np.array(sorted(cases))
    np.savez_compressed(os.path.join(HERE, 'identity.npz'), **out)
    print('identity.npz', len(out), 'arrays')


def pooling(ns):
    """The reference's own pooling.py, with torch_scatter.scatter supplied by oracle/pooling.py."""
    from oracle import pooling as opool
    sys.modules['torch_scatter'].scatter = opool.scatter
    pyg_shim._module('graphgym.contrib.pooling')
    ns.cfg.dataset = pyg_shim._Cfg(transform='none')
    ref = pyg_shim._load('graphgym.models.pooling', os.path.join(REF, 'graphgym/models/pooling.py'))
    out = {}
    g = torch.Generator().manual_seed(11)
    sizes = [5, 1, 0, 64, 30, 17, 0]                      # empty graphs in the middle and at the end
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    n, f = batch.numel(), 24
    x0 = torch.randn(n, f, generator=g)
    gy = torch.randn(len(sizes), f, generator=g)
    ids = torch.sort(torch.randperm(n, generator=g)[:n // 3]).values
    out['x'], out['batch'], out['ids'], out['gy'] = x0.numpy(), batch.numpy(), ids.numpy(), gy.numpy()
    out['size'] = np.int64(len(sizes))
    for transform in ('none', 'ego'):
        ns.cfg.dataset.transform = transform
        for mode in ('add', 'mean', 'max'):
            x = x0.clone().requires_grad_(True)
            y = ref.pooling_dict[mode](x, batch, ids, size=len(sizes))
            y.backward(gy)
            out['%s_%s/y' % (transform, mode)] = y.detach().numpy()
            out['%s_%s/gx' % (transform, mode)] = x.grad.numpy()
    np.savez_compressed(os.path.join(HERE, 'pooling.npz'), **out)
    print('pooling.npz', len(out), 'arrays')


class _Graph:
    def __init__(self, G):
        self.G = G
        self.num_nodes = G.number_of_nodes()


def canonical_from_nx(G_out, n, sizes):
    """member sets / induced edge sets per centre in ORIGINAL ids (node attr 'orig').  Block c is the
    centre copy c plus the id range the reference's running counter handed to ego c
    (transform.py:24-32); ``sizes[c]`` = number of non-centre members."""
    orig = nx.get_node_attributes(G_out, 'orig')
    mem_flat, mem_ptr, edge_flat, edge_ptr = [], [0], [], [0]
    bias = n
    for c in range(n):
        block = [c] + list(range(bias, bias + sizes[c]))
        bias += sizes[c]
        mem = sorted(orig[v] for v in block)
        edges = sorted((min(orig[u], orig[v]), max(orig[u], orig[v]))
                       for u, v in G_out.subgraph(block).edges())
        mem_flat += mem
        mem_ptr.append(len(mem_flat))
        edge_flat += [x for e in edges for x in e]
        edge_ptr.append(len(edge_flat) // 2)
    return (np.array(mem_flat, dtype=np.int64), np.array(mem_ptr, dtype=np.int64),
            np.array(edge_flat, dtype=np.int64).reshape(-1, 2), np.array(edge_ptr, dtype=np.int64))


def egonets(ns):
    out = {}
    graphs = {'C4': nx.cycle_graph(4)}
    for i, G in enumerate(fixture_graphs('scalefree', [0, 3])):
        graphs['scalefree%d' % i] = nx.Graph(G.edges())
    for i, G in enumerate(fixture_graphs('ba', [1])):
        graphs['ba%d' % i] = nx.Graph(G.edges())
    names = []
    for gname, G in graphs.items():
        G = nx.convert_node_labels_to_integers(G, ordering='sorted')
        n = G.number_of_nodes()
        ei = nx_to_edge_index(G)
        for radius in (1, 2, 3, 5):
            H = G.copy()
            nx.set_node_attributes(H, {v: v for v in H.nodes}, 'orig')
            g = _Graph(H)
            ns.transform.ego_nets(g, radius=radius)
            sizes = [(n if radius > 4 else len(nx.ego_graph(G, c, radius=radius))) - 1 for c in range(n)]
            assert g.G.number_of_nodes() == n + sum(sizes)
            mem, mem_ptr, edges, edge_ptr = canonical_from_nx(g.G, n, sizes)
            tag = '%s_r%d' % (gname, radius)
            names.append(tag)
            out[tag + '/edge_index'] = ei
            out[tag + '/n'] = np.int64(n)
            out[tag + '/radius'] = np.int64(radius)
            out[tag + '/num_nodes_out'] = np.int64(g.G.number_of_nodes())
            out[tag + '/num_edges_out'] = np.int64(g.G.number_of_edges())
            out[tag + '/node_id_index'] = g.node_id_index.numpy()
            out[tag + '/members'] = mem.astype(np.int32)
            out[tag + '/member_ptr'] = mem_ptr
            out[tag + '/edges'] = edges.astype(np.int32)
            out[tag + '/edge_ptr'] = edge_ptr
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'egonets.npz'), **out)
    print('egonets.npz', len(out), 'arrays')


def contrib_layers(ns):
    """generalconv (GeneralConvLayer, ref: contrib/layer/generalconv.py:12-113), sageinitconv (sageinitconv.py:12-115) and
    idconv with normalize_adj + agg='mean', run through the reference's own files."""
    root = os.path.join(REF, 'graphgym')
    gen = pyg_shim._load('graphgym.contrib.layer.generalconv', os.path.join(root, 'contrib/layer/generalconv.py'))
    pyg_shim._load('graphgym.contrib.layer.sageinitconv', os.path.join(root, 'contrib/layer/sageinitconv.py'))

    class GeneralConv(torch.nn.Module):           # ref: graphgym/models/layer.py:188-196 (imports pyg at module level)
        def __init__(self, dim_in, dim_out, bias=False, **kwargs):
            super().__init__()
            self.model = gen.GeneralConvLayer(dim_in, dim_out, bias=bias)

        def forward(self, batch):
            batch.node_feature = self.model(batch.node_feature, batch.edge_index)
            return batch
    ns.register.layer_dict['generalconv'] = GeneralConv
    out, tags = {}, []
    for self_msg in ('none', 'add', 'concat'):
        for kw in ({}, {'agg': 'mean'}, {'normalize_adj': True}, {'agg': 'mean', 'normalize_adj': True}):
            ns.cfg.gnn.self_msg = self_msg
            tags.append(run_layer(ns, 'generalconv', kw, n=41, fin=12, fout=16, seed=500 + len(tags), out=out,
                                  suffix='_selfmsg-' + self_msg))
    tags.append(run_layer(ns, 'sageinitconv', {}, n=53, fin=20, fout=24, seed=600, out=out))
    # gaddconv / gmulconv (attconv.py:14-240), default cfg (att_heads = 1, agg add, no normalisation); a wide case too
    pyg_shim._load('graphgym.contrib.layer.attconv', os.path.join(root, 'contrib/layer/attconv.py'))
    for name in ('gaddconv', 'gmulconv'):
        tags.append(run_layer(ns, name, {}, n=47, fin=12, fout=16, seed=700 + len(tags), out=out))
        tags.append(run_layer(ns, name, {}, n=61, fin=20, fout=136, seed=720 + len(tags), out=out, suffix='_wide'))
    tags.append(run_layer(ns, 'idconv', {'agg': 'mean', 'normalize_adj': True}, n=37, fin=12, fout=16, seed=601, out=out))
    out['tags'] = np.array(tags)
    np.savez_compressed(os.path.join(HERE, 'contrib_layers.npz'), **out)
    print('contrib_layers.npz', len(tags), 'cases')


def driver_cfg():
    """The keys of the reference's config/idgcn_tf/idgcn_node_scalefree.yaml the main_zd-style driver reads, re-emitted as
    a fixture (the reference tree is absent at test time)."""
    import yaml
    with open(os.path.join(REF, 'config/idgcn_tf/idgcn_node_scalefree.yaml')) as f:
        c = yaml.safe_load(f)
    keep = {'dataset': ['name', 'task', 'task_type', 'split', 'augment_feature', 'augment_feature_dims', 'augment_feature_repr',
                        'augment_label', 'augment_label_dims', 'transform'],
            'train': ['batch_size'], 'gnn': None, 'optim': ['optimizer', 'base_lr']}
    out = {sec: {k: v for k, v in c[sec].items() if keys is None or k in keys} for sec, keys in keep.items()}
    with open(os.path.join(HERE, 'idgcn_node_scalefree.yaml'), 'w') as f:
        f.write('# derived from the reference config file config/idgcn_tf/idgcn_node_scalefree.yaml by tests/golden/make_golden.py\n')
        yaml.safe_dump(out, f, sort_keys=True)
    print('idgcn_node_scalefree.yaml', out['gnn'])


def scalefree16(ns):
    graphs = fixture_graphs('scalefree', range(16))
    out, eis, ptr, clus, ego_nodes, ego_edges = {}, [], [0], [], [], []
    for G in graphs:
        G = nx.convert_node_labels_to_integers(nx.Graph(G.edges()), ordering='sorted')
        n = G.number_of_nodes()
        eis.append(nx_to_edge_index(G) + ptr[-1])
        ptr.append(ptr[-1] + n)
        c = nx.clustering(G)
        clus += [c[v] for v in range(n)]                      # feature_augment.py:81-82: list(nx.clustering(G).values())
        g = _Graph(G.copy())
        ns.transform.ego_nets(g, radius=3)                    # the reference's own transform (cfg.gnn.layers_mp = 3)
        ego_nodes.append(g.G.number_of_nodes())
        ego_edges.append(g.G.number_of_edges())
    arr = np.array(clus)
    # balanced binning, feature_dim = 10 (feature_augment.py:219-231), then np.digitize(arr, bins) - 1 (:139-140)
    sorted_arr = np.sort(arr)
    bin_indices = np.linspace(0, len(arr), num=10, endpoint=False).astype(int)
    bins = np.unique(sorted_arr[bin_indices])
    out['edge_index'] = np.concatenate(eis, axis=1)
    out['graph_ptr'] = np.array(ptr, dtype=np.int64)
    out['clustering'] = arr
    out['bin_edges'] = bins
    out['label'] = (np.digitize(arr, bins) - 1).astype(np.int64)
    out['ego_nodes'] = np.array(ego_nodes, dtype=np.int64)
    out['ego_undirected_edges'] = np.array(ego_edges, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, 'scalefree16.npz'), **out)
    print('scalefree16.npz', {k: v.shape for k, v in out.items()}, 'ego nodes', sum(ego_nodes))


if __name__ == '__main__':
    ns = pyg_shim.install(REF)
    if len(sys.argv) > 1 and sys.argv[1] in ('scalefree16', 'contrib_layers'):
        {'scalefree16': scalefree16, 'contrib_layers': contrib_layers}[sys.argv[1]](ns)
        sys.exit(0)
    layers(ns)
    identity(ns)
    egonets(ns)
    pooling(ns)
    scalefree16(ns)
    contrib_layers(ns)
    driver_cfg()
