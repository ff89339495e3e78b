"""``main_zd.py``-style front door for the accelerated path (ref: main_zd.py:260-311, graphgym/train.py:15-32,67-81).

    python -m graphgym_b200.main_zd --cfg /path/to/config/idgcn_tf/idgcn_node_scalefree.yaml \\
        --graphs tests/golden/scalefree16.npz [--epochs 20] [--repeat 1]

What the reference driver does per experiment, kept in the same order: merge the yaml into ``cfg``; seed; build the dataset
(``create_dataset``: label augmentation, 80/20 graph split, ``transform_after_split`` = ego-nets of radius
``cfg.gnn.layers_mp`` when ``dataset.transform == 'ego'``); ``create_loader`` (batches of ``train.batch_size`` graphs,
collated block-diagonally); pick the model from ``cfg.gnn.layer_type`` (the ``Tfg-*`` names of main_zd.py:299-308 resolve to
the registered layers); Adam with ``optim.base_lr``; train / eval epochs; report accuracy.

Everything between the edge lists and the logits runs on the device through the gg_* path: clustering-coefficient labels
(cycle kernel) and balanced binning, ego-net extraction, batch collation, the layers with fused post-ops, the node head.
Torch supplies the optimizer, the cross-entropy loss and autograd plumbing — they are outside SURVEY §8's hot path.
The dataset comes from an ``.npz`` (``edge_index`` [2,E] symmetric directed + ``graph_ptr`` [G+1], e.g. the committed
``tests/golden/scalefree16.npz`` made from the reference's ``datasets/scalefree.pkl``) or from the seeded BA generator.
"""
import argparse
import json
import random
import time

import numpy as np
import torch
import torch.nn.functional as F

from graphgym_b200 import loader
from graphgym_b200.config import cfg, load_cfg, reset_cfg
from graphgym_b200.contrib.transform import binning
from graphgym_b200.contrib.transform.clustering import clustering_coefficient
from graphgym_b200.contrib.transform.identity import compute_identity
from graphgym_b200.models import transform as gtr
from graphgym_b200.models.feature_augment import Preprocess
from graphgym_b200.models.gnn import GNN


def load_graphs(path, dev, graphs=64, nodes=64, seed=0):
    if path:
        d = np.load(path)
        return torch.from_numpy(d['edge_index']).to(dev), torch.from_numpy(d['graph_ptr']).to(dev)
    g = torch.Generator().manual_seed(seed)                 # syn_graph.py-like BA batch, m = 4
    srcs, ptr = [], [0]
    for i in range(graphs):
        t = torch.arange(4, nodes).repeat_interleave(4)
        tgt = (torch.rand(t.numel(), generator=g) ** 2 * t).long().clamp(max=nodes - 1)
        tgt = torch.minimum(tgt, t - 1)
        code = torch.unique(torch.minimum(t, tgt) * nodes + torch.maximum(t, tgt))
        a, b = code // nodes + ptr[-1], code % nodes + ptr[-1]
        srcs.append(torch.stack([torch.cat([a, b]), torch.cat([b, a])]))
        ptr.append(ptr[-1] + nodes)
    return torch.cat(srcs, 1).to(dev), torch.tensor(ptr, device=dev)


def create_dataset(edge_index, graph_ptr, dev):
    """-> list of GraphData (one per graph, tensors in HBM) with node_feature, edge_index, node_label (+ node_id_index and
    ego expansion when cfg.dataset.transform == 'ego'; + node_identity when 'node_identity' is an augment feature)."""
    n = int(graph_ptr[-1])
    label_dim = int(cfg.dataset.get('augment_label_dims', 10))
    clus = clustering_coefficient(edge_index, n, graph_ptr=graph_ptr.int())          # nx.clustering (feature_augment.py:81-82)
    labels, edges = binning.balanced_labels(clus, label_dim)                         # balanced binning (:208-245)
    feats = {}
    if 'node_identity' in cfg.dataset.augment_feature:                               # ID-GNN Fast (identity.py:25-35)
        k = int(cfg.dataset.augment_feature_dims[list(cfg.dataset.augment_feature).index('node_identity')])
        feats['node_identity'] = compute_identity(edge_index, n, k, graph_ptr=graph_ptr.int())
    gp = graph_ptr.tolist()
    src_graph = torch.bucketize(edge_index[0], graph_ptr[1:], right=True)
    graphs = []
    ego = cfg.dataset.transform == 'ego'
    for g in range(len(gp) - 1):
        lo, hi = gp[g], gp[g + 1]
        ei = edge_index[:, src_graph == g] - lo
        data = dict(node_feature=torch.ones(hi - lo, 1, device=dev), edge_index=ei.contiguous(),
                    node_label=labels[lo:hi].contiguous())
        for k_, v in feats.items():
            data[k_] = v[lo:hi].contiguous()
        if ego:                                                                      # transform_after_split (loader.py:175-180)
            res = gtr.ego_nets_batch(ei.contiguous(), hi - lo, int(cfg.gnn.layers_mp))
            orig = res['orig_id']
            data['node_feature'] = data['node_feature'].index_select(0, orig)
            for k_ in feats:
                data[k_] = data[k_].index_select(0, orig)
            data['edge_index'] = res['edge_index']
            data['node_id_index'] = res['node_id_index']
        graphs.append(loader.GraphData(**data))
    return graphs, len(edges)


def run_epoch(model, pre, batches, optimizer=None):
    tot_loss = correct = count = 0
    for graphs in batches:
        batch = pre(loader.collate(graphs))                                          # on-device collation + Preprocess concat
        true = batch.node_label
        if optimizer is not None:
            optimizer.zero_grad()
        pred, _ = model(batch)
        loss = F.cross_entropy(pred, true)                                           # compute_loss (loss.py): cross_entropy
        if optimizer is not None:
            loss.backward()
            optimizer.step()
        tot_loss += float(loss) * true.numel()
        correct += int((pred.argmax(1) == true).sum())
        count += true.numel()
    return tot_loss / max(count, 1), correct / max(count, 1)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default=None, help="one of the reference's config/*_tf/*.yaml files")
    ap.add_argument('--model', default=None, help='layer type when no yaml is given (e.g. Tfg-idgcn, gcnconv)')
    ap.add_argument('--graphs', default=None, help='.npz with edge_index [2,E] and graph_ptr [G+1]; default: seeded BA batch')
    ap.add_argument('--epochs', type=int, default=20)
    ap.add_argument('--repeat', type=int, default=1)
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit('graphgym_b200.main_zd needs a CUDA device: the accelerated path has no CPU fallback')
    dev = torch.device('cuda', torch.cuda.current_device())
    results = []
    for rep in range(args.repeat):
        reset_cfg()
        cfg.optim = {'base_lr': 0.01}
        if args.cfg:
            load_cfg(args.cfg)
        if args.model:
            cfg.gnn.layer_type = args.model
        seed = rep + 1                                                               # main_zd.py:288-291
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        t0 = time.time()
        edge_index, graph_ptr = load_graphs(args.graphs, dev, seed=seed)
        graphs, num_labels = create_dataset(edge_index, graph_ptr, dev)
        perm = torch.randperm(len(graphs)).tolist()
        split = max(1, int(0.8 * len(graphs)))                                       # dataset.split = [0.8, 0.2]
        train_g, val_g = [graphs[i] for i in perm[:split]], [graphs[i] for i in perm[split:]] or [graphs[perm[0]]]
        bs = int(cfg.train.batch_size)
        chunks = lambda gs: [gs[i:i + bs] for i in range(0, len(gs), bs)]
        pre = Preprocess(1)
        model = GNN(pre.dim_out, num_labels).to(dev)
        optimizer = torch.optim.Adam(model.parameters(), lr=float(cfg.optim['base_lr']))
        torch.cuda.synchronize()
        prep_s = time.time() - t0
        t0 = time.time()
        for epoch in range(args.epochs):
            model.train()
            loss, acc = run_epoch(model, pre, chunks(train_g), optimizer)
        model.eval()
        with torch.no_grad():
            val_loss, val_acc = run_epoch(model, pre, chunks(val_g))
        torch.cuda.synchronize()
        results.append({'repeat': rep, 'layer_type': cfg.gnn.layer_type, 'graphs': len(graphs), 'labels': num_labels,
                        'train_loss': round(loss, 4), 'train_acc': round(acc, 4), 'val_loss': round(val_loss, 4),
                        'val_acc': round(val_acc, 4), 'dataset_s': round(prep_s, 2),
                        's_per_epoch': round((time.time() - t0) / max(args.epochs, 1), 4)})
        print(json.dumps(results[-1]))
    reset_cfg()
    return results


if __name__ == '__main__':
    main()
