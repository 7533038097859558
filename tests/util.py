"""shared test helpers"""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k] for k in z.files}


def load_json(name):
    return json.load(open(os.path.join(GOLD, name)))


def mismatches(a, b):
    """count positions where a and b differ (NaN == NaN, +0 == -0, inf must match)"""
    a = torch.as_tensor(a).detach().cpu().float()
    b = torch.as_tensor(b).detach().cpu().float()
    assert a.shape == b.shape, (a.shape, b.shape)
    both_nan = torch.isnan(a) & torch.isnan(b)
    return int(((a != b) & ~both_nan).sum())


def assert_exact(a, b, what=""):
    n = mismatches(a, b)
    assert n == 0, "%s: %d mismatching elements of %d" % (what, n, torch.as_tensor(a).numel())
