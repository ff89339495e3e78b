"""graphgym_b200 — B200-native drop-in for GraphGym's message-passing layer path.

Public surface (mirrors the reference):
    from graphgym_b200.register import register_layer, layer_dict      # graphgym/register.py
    from graphgym_b200.models.layer import layer_dict, GeneralLayer     # graphgym/models/layer.py
    from graphgym_b200.contrib.transform.identity import compute_identity
    from graphgym_b200.models.transform import ego_nets
Everything below the layer API is hand-written sm_100a CUDA behind the C-ABI in include/gg_b200.h
(graphgym_b200/libgg_b200.so).  There is no CPU fallback: ops raise if the library is missing or
if they are handed CPU tensors.
"""
__version__ = "0.1.0"
