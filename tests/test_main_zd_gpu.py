"""The main_zd-style front door (graphgym_b200/main_zd.py; ref: main_zd.py:260-311) on Cfg-A: the reference's idgcn_tf
config (fixture derived from config/idgcn_tf/idgcn_node_scalefree.yaml) on the bundled ScaleFree graphs [0:16]."""
import os

import pytest

pytestmark = pytest.mark.gpu


def test_driver_trains_idgcn_on_the_scalefree_fixture(cuda, golden_dir):
    from graphgym_b200 import main_zd
    res = main_zd.main(['--cfg', os.path.join(golden_dir, 'idgcn_node_scalefree.yaml'),
                        '--graphs', os.path.join(golden_dir, 'scalefree16.npz'), '--epochs', '8'])
    r = res[0]
    assert r['layer_type'] == 'Tfg-idgcn' and r['graphs'] == 16 and r['labels'] == 10
    assert r['train_loss'] == r['train_loss'] and r['train_loss'] < 2.3     # finite, below ln(10): it learns
    first = main_zd.main(['--cfg', os.path.join(golden_dir, 'idgcn_node_scalefree.yaml'),
                          '--graphs', os.path.join(golden_dir, 'scalefree16.npz'), '--epochs', '1'])[0]
    assert r['train_loss'] < first['train_loss']


def test_driver_plain_layer_on_synthetic_batch(cuda):
    from graphgym_b200 import main_zd
    r = main_zd.main(['--model', 'Tfg-sageconv', '--epochs', '2'])[0]
    assert r['graphs'] == 64 and r['train_loss'] == r['train_loss']
