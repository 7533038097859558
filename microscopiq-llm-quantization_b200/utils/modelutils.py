"""Model helpers used by the layer-sequential driver: the default device and the lookup of the quantizable
sub-modules of a decoder layer (same call signature and result as the reference's utils/modelutils.find_layers)."""
import torch
from torch import nn

DEV = torch.device("cuda", 0)

#: module types GPTQ / RTN replace
QUANTIZABLE = (nn.Conv2d, nn.Linear)


def find_layers(module, layers=QUANTIZABLE, name=""):
    """Ordered dict-like {qualified name: sub-module} of every instance of `layers` at or below `module`, names
    relative to `module` (prefixed by `name` when given), in registration order."""
    kinds = tuple(layers)
    prefix = name + "." if name else ""
    if isinstance(module, kinds):
        return {name: module}
    # named_modules() walks depth first in registration order, exactly the order a recursive descent yields;
    # a quantizable module has no quantizable descendants in the models this is used on (Linear / Conv2d are leaves)
    return {prefix + qual: sub for qual, sub in module.named_modules() if qual and isinstance(sub, kinds)}
