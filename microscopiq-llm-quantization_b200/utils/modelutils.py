"""Model helpers used by the layer-sequential driver: the default device and the lookup of the quantizable
sub-modules of a decoder layer (same call signature and result as the reference's utils/modelutils.find_layers)."""
import torch
from torch import nn

DEV = torch.device("cuda", 0)

#: module types GPTQ / RTN replace
QUANTIZABLE = (nn.Conv2d, nn.Linear)


def find_layers(module, layers=QUANTIZABLE, name=""):
    """{qualified name: sub-module} of every sub-module whose type is EXACTLY one of `layers` (utils/modelutils.py:8-16:
    `type(module) in layers`, so subclasses such as NonDynamicallyQuantizableLinear or MXLinear are skipped, as in
    the reference), in registration order; matched modules are not descended into and a module shared between two
    parents is listed under both names."""
    kinds = tuple(layers)
    found = {}
    stack = [(name, module)]
    while stack:
        qual, mod = stack.pop()
        if type(mod) in kinds:
            found[qual] = mod
            continue
        children = [((qual + "." + n) if qual else n, c) for n, c in mod.named_children()]
        stack.extend(reversed(children))
    return found
