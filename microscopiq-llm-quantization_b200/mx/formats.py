"""Element-format table of the MX number system.

Mirror of the reference's number_system/mx/formats.py (names, enum values and the
5-tuple returned by `_get_format_params` are the API contract, formats.py:65-129); the
table itself is data-driven here.
"""
from enum import Enum, IntEnum

FP32_EXPONENT_BIAS = 127
FP32_MIN_NORMAL = 2 ** (-FP32_EXPONENT_BIAS + 1)


class RoundingMode(IntEnum):
    """int value == `rmode` of the C ABI (include/msq_b200.h) == cpp/common.cuh:130-134"""
    nearest = 0
    floor = 1
    even = 2

    @staticmethod
    def string_enums():
        return [m.name for m in RoundingMode]


class ElemFormat(Enum):
    int8 = 1
    int4 = 2
    int2 = 3
    fp8_e5m2 = 4
    fp8_e4m3 = 5
    fp6_e3m2 = 6
    fp6_e2m3 = 7
    fp4 = 8
    fp4_e2m1 = 8
    float16 = 9
    fp16 = 9
    bfloat16 = 10
    bf16 = 10

    @staticmethod
    def from_str(s):
        assert s is not None, "String elem_format == None"
        key = s.lower()
        if key in ElemFormat.__members__:
            return ElemFormat[key]
        raise Exception("Undefined elem format", s)


# format -> (exponent bits, mantissa bits incl. sign and implicit one, has Inf/NaN codes)
_LAYOUT = {
    ElemFormat.int8: (0, 8, False), ElemFormat.int4: (0, 4, False), ElemFormat.int2: (0, 2, False),
    ElemFormat.fp8_e5m2: (5, 4, True), ElemFormat.fp8_e4m3: (4, 5, False),
    ElemFormat.fp6_e3m2: (3, 4, False), ElemFormat.fp6_e2m3: (2, 5, False),
    ElemFormat.fp4: (2, 3, False), ElemFormat.float16: (5, 12, True), ElemFormat.bfloat16: (8, 9, True),
}
_CACHE = {}


def _get_min_norm(ebits):
    """smallest normal magnitude of a float format with `ebits` exponent bits"""
    return 0 if ebits == 0 else 2 ** (2 - 2 ** (ebits - 1))


def _get_max_norm(ebits, mbits):
    """largest normal of an IEEE-style format whose top exponent encodes Inf/NaN"""
    assert ebits >= 5, "invalid for floats that don't define NaN"
    emax = 2 ** (ebits - 1) - 1
    return 2 ** emax * float(2 ** (mbits - 1) - 1) / 2 ** (mbits - 2)


def _get_format_params(fmt):
    """-> (ebits, mbits, emax, max_norm, min_norm); mbits counts sign + implicit bit."""
    if isinstance(fmt, str):
        fmt = ElemFormat.from_str(fmt)
    hit = _CACHE.get(fmt)
    if hit is not None:
        return hit
    if fmt not in _LAYOUT:
        raise Exception("Unknown element format %s" % fmt)
    ebits, mbits, has_inf = _LAYOUT[fmt]
    if ebits == 0:
        emax = 0
    else:
        emax = 2 ** (ebits - 1) - (1 if has_inf else 0)
    if fmt is ElemFormat.fp8_e4m3:
        max_norm = 2 ** emax * 1.75          # S.1111.111 is NaN, so the top mantissa is .110
    else:
        max_norm = 2 ** emax * float(2 ** (mbits - 1) - 1) / 2 ** (mbits - 2)
    out = (ebits, mbits, emax, max_norm, _get_min_norm(ebits))
    _CACHE[fmt] = out
    return out
