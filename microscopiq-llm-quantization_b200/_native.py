"""ctypes binding of libmsq_b200.so (the C ABI declared in include/msq_b200.h).

There is no fallback: if the library is missing or a tensor is not a CUDA tensor the
call raises.  PyTorch is used for device memory and streams only.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsq_b200.so")
_lib = None

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_ALIGN = 0, -1, -2, -3, -4
_ERR = {-1: "bad argument", -2: "unsupported geometry/format for this path", -3: "CUDA launch error",
        -4: "buffer not 16-byte aligned"}

ST_NONFINITE_INPUT, ST_EXP_IN_OVERFLOW, ST_EXP_OUT_OVERFLOW = 1, 2, 4
ROUND = {"nearest": 0, "floor": 1, "even": 2}

_p, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
_SIGS = {
    "msq_version": (_i, []),
    "msq_seg_blocks": (_i, []),
    "msq_packed_code_width": (_i, [_i, _i]),
    "msq_packed_hi_width": (_i, [_i, _i, _i, _i]),
    "msq_packed_hi_seg_stride": (_ll, [_i, _i]),
    "msq_quantize_elemwise": (_i, [_p, _p, _ll, _i, _i, _i, _f, _i, _i, _i, _i, _p]),
    "msq_quantize_mx_f32": (_i, [_p, _p, _ll, _ll, _ll, _i, _i, _i, _i, _i, _f, _i, _i, _i, _p, _p, _p, _p,
                                 C.POINTER(_i), _p]),
    "msq_quantize_mx_outlier_f32": (_i, [_p, _p, _ll, _ll, _ll, _i, _i, _i,
                                         _i, _i, _i, _f, _i, _i, _i, _f,
                                         _f, _i, _i, _i, _p, _p,
                                         _p, _p, _p, _p, _p, _p, _p, _p, _p, C.POINTER(_i), _p]),
    "msq_outlier_bounds_across_blocks_f32": (_i, [_p, _p, _p, _ll, _ll, _ll, _i, _f, _p]),
}


def declared_symbols():
    return sorted(_SIGS)


def lib():
    """Load (once) and return the C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmsq_b200.so is not built (%s). Run `python __graft_entry__.py build` "
                "or `python microscopiq-llm-quantization_b200/build.py`; there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def register(name, restype, argtypes):
    """Used by the other binding modules (gptq, gemm, pack) to add their signatures."""
    _SIGS[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s (code %d)" % (what, _ERR.get(rc, "unknown"), rc))


def require_cuda(t, what="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the B200 kernels have no CPU fallback" % what)
    return t


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
