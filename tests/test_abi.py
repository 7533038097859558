"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/msq_b200.h
declares; the header compiles as plain C; the Python bindings declare the same set; the product
never imports the oracle and refuses CPU tensors."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msq_b200.h")
PKG = os.path.join(ROOT, "microscopiq-llm-quantization_b200")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import msq_b200
    lib = ctypes.CDLL(msq_b200._native.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    lib.msq_version.restype = ctypes.c_int
    assert lib.msq_version() >= 100                      # a call that needs no GPU
    assert lib.msq_seg_blocks() == 1024
    assert lib.msq_packed_code_width(2, 3) == 4 and lib.msq_packed_hi_width(0, 2, 2, 3) == 2


def test_python_bindings_cover_the_header():
    import msq_b200
    from msq_b200 import _misc, _tc, _uniform, packed  # noqa: F401  (register their signatures)
    from msq_b200.llm import gptq  # noqa: F401
    assert set(msq_b200._native.declared_symbols()) == set(header_symbols())


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "msq_b200.h"\nint main(void) { return MSQ_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                           str(tmp_path / "t.o")])


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: no product module may import, load or exec anything under oracle/"""
    pat = re.compile(r"^\s*(from|import)\s+\S*(msq_oracle|build_ref|oracle)\b|CDLL\([^)]*oracle|open\([^)]*oracle", re.M)
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(base, f)).read()
                assert not pat.search(txt), f
    shim = open(os.path.join(ROOT, "msq_b200", "__init__.py")).read()
    assert not pat.search(shim)


def test_no_cpu_fallback():
    import msq_b200
    x = torch.randn(32, 32)
    with pytest.raises(RuntimeError):
        msq_b200._quantize_mx(x, 8, "fp4_e2m1", axes=[1], block_size=32)
    with pytest.raises(RuntimeError):
        msq_b200._quantize_elemwise_core(x, 9, 8, 3e38)
    with pytest.raises(RuntimeError):
        msq_b200.pack_mx(x, 8, "fp4_e2m1", axes=[1], block_size=32)
    with pytest.raises(RuntimeError):
        msq_b200.quantize(x, torch.ones(32, 1), torch.zeros(32, 1), 15)
    layer = torch.nn.Linear(32, 32)
    g = msq_b200.GPTQ(layer)
    with pytest.raises(RuntimeError):
        g.add_batch(torch.randn(1, 8, 32), None)


def test_api_surface_matches_reference_names():
    """same names / positional order as utils/quant.py, mx/mx_ops.py, llm/gptq.py"""
    import inspect
    import msq_b200
    from msq_b200.utils import quant
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(quant.quantize_mx_outlier_v1) == ["A", "inlier_scale_bits", "outlier_scale_bits", "inlier_elem_format",
                                                  "outlier_elem_format", "shared_exp_method", "std_dev", "axes", "block_size",
                                                  "round", "flush_fp32_subnorms", "custom_cuda"]
    assert sig(quant.quantize_mx_outlier_hessian)[-2:] == ["prune_inliers", "custom_cuda"]
    assert sig(quant.MXQuantizer.configure)[1:5] == ["inlier_scale_bits", "outlier_scale_bits", "inlier_elem_format",
                                                     "outlier_elem_format"]
    assert sig(msq_b200.GPTQ.fasterquant) == ["self", "blocksize", "percdamp", "groupsize", "actorder", "static_groups"]
    assert sig(msq_b200._quantize_mx)[:9] == ["A", "scale_bits", "elem_format", "shared_exp_method", "axes", "block_size",
                                              "round", "flush_fp32_subnorms", "custom_cuda"]
    q = quant.MXQuantizer()
    assert q.ready() is True and q.enabled() is None and q.find_params(None) is None
    u = quant.Quantizer()
    assert set(dict(u.named_buffers())) == {"maxq", "scale", "zero"}
    specs = msq_b200.finalize_mx_specs({"w_elem_format": "fp6_e3m2", "a_elem_format": "fp6_e3m2", "block_size": 32})
    assert specs["w_elem_format_bp"] == "fp6_e3m2" and specs["round_mx_output"] == "nearest"
    assert msq_b200.finalize_mx_specs({}) is None


def test_magic_divider_is_exact():
    """host-side construction of the kernels' exact divider: floor(n / d) == (n * m) >> k for 0 <= n < 2^31"""
    import ctypes as C
    import random
    import msq_b200  # noqa: F401  (registers the signatures)
    from msq_b200 import _native as N
    L = N.lib()
    rnd = random.Random(7)
    ds = [1, 2, 3, 5, 7, 16, 31, 32, 33, 4096, 11008, 4095, 4097, 65535, 65536, 65537, 2 ** 30, 2 ** 31 - 2]
    ds += [rnd.randrange(1, 2 ** 31 - 1) for _ in range(200)] + [rnd.randrange(1, 100000) for _ in range(200)]
    for d in ds:
        m, k = C.c_uint(0), C.c_int(0)
        assert L.msq_magic_div(d, C.byref(m), C.byref(k)) == 0
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1, 2 ** 31 - 2, (2 ** 31 - 1) // d * d, (2 ** 31 - 1) // d * d - 1]
        ns += [rnd.randrange(0, 2 ** 31) for _ in range(50)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                assert (n * m.value) >> k.value == n // d, (n, d, m.value, k.value)
    assert L.msq_magic_div(0, C.byref(m), C.byref(k)) != 0
