"""Layer registry with the reference's contract (ref: graphgym/register.py:6-10,32-34).

``register_layer(key, module)`` adds to ``layer_dict``; a duplicate key raises
``KeyError('Key {} is already pre-defined.')`` exactly like the reference.
"""


def register(key, module, module_dict):
    if key in module_dict:
        raise KeyError('Key {} is already pre-defined.'.format(key))
    else:
        module_dict[key] = module


layer_dict = {}


def register_layer(key, module):
    register(key, module, layer_dict)
