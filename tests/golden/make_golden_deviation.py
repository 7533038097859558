"""Enumerate, at FULL bench size, every element where the REFERENCE (utils/quant.py:147 quantize_mx_outlier_v1, imported
in place from /root/reference, torch CPU) and the oracle / kernels (canonical pairwise-tree fp32 statistics, DESIGN.md
section 2 deviation class 2) disagree on bench.py's workload: the seven LLaMA-2-7B decoder-layer linears, seed 1000.

    python tests/golden/make_golden_deviation.py        (authoring container only; ~2 min)

Writes tests/golden/deviation_bench.json: per matrix the flat indices of the differing elements, the input value, the
reference's output and the oracle's output.  The GPU tier asserts that the kernel equals the oracle at full size and
therefore differs from the reference in exactly this set.
"""
import json
import os
import sys
import warnings

import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [REF + "/number_system", REF, REF + "/llm", os.path.join(ROOT, "oracle")]

from utils.quant import quantize_mx_outlier_v1  # noqa: E402
import msq_oracle as O  # noqa: E402

LAYER_SHAPES = [("q_proj", 4096, 4096), ("k_proj", 4096, 4096), ("v_proj", 4096, 4096), ("o_proj", 4096, 4096),
                ("gate_proj", 11008, 4096), ("up_proj", 11008, 4096), ("down_proj", 4096, 11008)]


def main():
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(1000)                      # bench.py make_weights(dev, 1000 + rank), rank 0
    out = {"seed": 1000, "quantizer": "fp4_e2m1/fp8_e4m3 sb 8/8 std_dev 2 axes=[0] block 16", "matrices": []}
    total = 0
    for name, r, c in LAYER_SHAPES:
        W = torch.randn(r, c, generator=g) * 0.02
        ref = quantize_mx_outlier_v1(W, 8, 8, "fp4_e2m1", "fp8_e4m3", "max", 2, [0], 16, "nearest", False, False)
        orc = O.quantize_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16)
        idx = torch.nonzero((ref != orc).reshape(-1)).reshape(-1)
        ent = {"name": name, "rows": r, "cols": c, "index": idx.tolist(),
               "x": [float(v) for v in W.reshape(-1)[idx]], "reference": [float(v) for v in ref.reshape(-1)[idx]],
               "oracle": [float(v) for v in orc.reshape(-1)[idx]],
               "oracle_sha1_of_sum": float(orc.double().sum())}
        total += idx.numel()
        out["matrices"].append(ent)
        print(name, idx.numel(), flush=True)
    out["total_differing"] = total
    out["total_elements"] = sum(r * c for _, r, c in LAYER_SHAPES)
    json.dump(out, open(os.path.join(HERE, "deviation_bench.json"), "w"), indent=1)
    print("total", total)


if __name__ == "__main__":
    main()
