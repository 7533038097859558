"""utils/modelutils.py mirror (reference: 15 lines): DEV and find_layers."""
import torch
import torch.nn as nn

DEV = torch.device("cuda:0")


def find_layers(module, layers=(nn.Conv2d, nn.Linear), name=""):
    """{qualified name: module} for every Linear / Conv2d below `module` (utils/modelutils.py:8-15)"""
    if isinstance(module, tuple(layers)):
        return {name: module}
    found = {}
    for child_name, child in module.named_children():
        found.update(find_layers(child, layers=layers, name=name + "." + child_name if name != "" else child_name))
    return found
