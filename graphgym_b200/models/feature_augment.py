"""The consumer side of feature augmentation (ref: graphgym/models/feature_augment.py:316-333).

``Preprocess`` keeps the reference's name, ``dim_dict`` / ``dim_out`` and ``forward(batch) -> batch``: the augmented
feature blocks named in ``cfg.dataset.augment_feature`` and ``node_feature`` are concatenated column-wise into
``batch.node_feature``.  The reference does ``torch.cat([batch[name].float() ...], dim=1)``; here every block is placed at
its column offset by one segmented-copy launch (csrc/collate.cu), converting int64 / uint8 blocks on the fly.
"""
import torch.nn as nn

from graphgym_b200 import loader
from graphgym_b200.config import cfg


class Preprocess(nn.Module):
    def __init__(self, dim_in):
        super().__init__()
        self.dim_dict = {name: dim for name, dim in zip(cfg.dataset.augment_feature, cfg.dataset.augment_feature_dims)}
        self.dim_dict['node_feature'] = dim_in
        self.dim_out = sum(self.dim_dict.values())

    def extra_repr(self):
        return '\n'.join(['{}: dim_out={}'.format(name, dim) for name, dim in self.dim_dict.items()]
                         + ['Total: dim_out={}'.format(self.dim_out)])

    def forward(self, batch):
        batch.node_feature = loader.concat_columns([batch[name] for name in self.dim_dict])
        return batch
