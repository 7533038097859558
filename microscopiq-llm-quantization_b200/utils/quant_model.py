"""quantize_model -- mirror of the reference's utils/quant_model.py:15-72: walk a pretrained model and replace every
`nn.Linear` by an `MXLinear` that runs the MX (fake-)quantized forward of number_system/mx/linear.py on the B200
kernels.  The reference also swaps Conv2d / LayerNorm-free modules for their mx counterparts; on this path only the
linears are quantized (the GPTQ / RTN drivers touch nothing else, llm/llama.py:128-159), convolutions are left as
they are.

    qmodel = quantize_model(model, mx_specs)                       # weights quantized on the fly, packed once
    qmodel = quantize_model(model, mx_specs, packed=load_packed(path))   # weights taken from a packed checkpoint
"""
import copy

import torch.nn as nn

from ..mx.linear import MXLinear


def _convert_linear(mod, mx_specs, packed, name, fast_activations=False):
    pw = None if packed is None else packed.get(name + ".weight", packed.get(name))
    if pw is not None and hasattr(pw, "dequantize"):
        bias = None
        if packed is not None and (name + ".bias") in packed:
            bias = packed[name + ".bias"]
        elif mod.bias is not None:
            bias = mod.bias.data
        return MXLinear.from_packed(pw, bias=bias, mx_specs=mx_specs, name=name, fast_activations=fast_activations)
    q = MXLinear(mod.in_features, mod.out_features, mod.bias is not None, mx_specs, name=name)
    q = q.to(device=mod.weight.device, dtype=mod.weight.dtype)
    q.weight.data = mod.weight.data.clone()
    if mod.bias is not None:
        q.bias.data = mod.bias.data.clone()
    q.fast_activations = bool(fast_activations)
    return q


def quantize_model(model, mx_specs, packed=None, skip=("lm_head",), _prefix="", fast_activations=False):
    """Recursively replace nn.Linear (exact type, like the reference's `type(model) == nn.Linear`) by MXLinear.
    `skip`: attribute names left untouched (the reference skips 'base_model' and 'lm_head', quant_model.py:69).
    `packed`: optional {qualified name(.weight): PackedMX} -- those linears are built with MXLinear.from_packed and run
    the packed-weight GEMM directly on the stored planes.
    `fast_activations`: serving mode of every converted linear (MXLinear.fast_activations: one-launch plain-MX activation
    quantizer + block-scaled GEMM instead of the reference's outlier-aware activation quantizer)."""
    if type(model) == nn.Linear:
        return _convert_linear(model, mx_specs, packed, _prefix, fast_activations)
    if isinstance(model, (nn.Sequential, nn.ModuleList)):
        mods = [quantize_model(m, mx_specs, packed, skip, (_prefix + "." if _prefix else "") + n, fast_activations)
                for n, m in model.named_children()]
        return type(model)(mods) if isinstance(model, nn.ModuleList) else nn.Sequential(*mods)
    q_model = copy.copy(model)
    q_model._modules = dict(model._modules)
    for n, child in model.named_children():
        if n in skip or n == "base_model":
            continue
        q_model._modules[n] = quantize_model(child, mx_specs, packed, skip, (_prefix + "." if _prefix else "") + n,
                                             fast_activations)
    return q_model
