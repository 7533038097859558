"""Known-answer tables the REFERENCE's own test-suite holds for the MX core,
transcribed as data with the reference file:line of each table.  (Paths are relative to
number_system/mx/tests/ in the reference repository.)

MODE: which of the reference's two disagreeing arithmetic modes satisfies the table --
"spec" (native funcs.cpp path) always does; "python" (mx_ops.py:444, `2^e + 1e-6`) only
where noted.
"""
INF = float("inf")
NAN = float("nan")

# test_formats.py:12-95  -> (ebits, emax, max_norm, min_norm) ; mbits where asserted
FORMAT_PARAMS = {
    "float16": (5, 15, 2 ** 15 * (1 + 1023 / 1024), 2 ** -14, None),
    "fp16": (5, 15, 2 ** 15 * (1 + 1023 / 1024), 2 ** -14, None),
    "bfloat16": (8, 127, 2 ** 127 * (1 + 127 / 128), 2 ** -126, None),
    "bf16": (8, 127, 2 ** 127 * (1 + 127 / 128), 2 ** -126, None),
    "fp8_e5m2": (5, 15, 2 ** 15 * 1.75, 2 ** -14, None),
    "fp8_e4m3": (4, 8, 2 ** 8 * 1.75, 2 ** -6, None),
    "fp6_e3m2": (3, 4, 2 ** 4 * 1.75, 2 ** -2, None),
    "fp6_e2m3": (2, 2, 2 ** 2 * 1.875, 2 ** 0, None),
    "fp4_e2m1": (2, 2, 2 ** 2 * 1.5, 2 ** 0, None),
    "fp4": (2, 2, 2 ** 2 * 1.5, 2 ** 0, None),
    "int8": (0, 0, 1 + 63 / 64, 0, 8),
    "int4": (0, 0, 1.75, 0, 4),
}

# test_corners_mx.py:27-56: NaN/Inf poison the whole block (round=floor, block = whole vector, scale_bits 8)
MX_NAN_INPUTS = [[NAN, 0, 2 ** 127, 0], [INF, 0, 2 ** 127, 0], [-INF, 0, 2 ** 127, 0]]
MX_NAN_FORMATS = ["fp8_e4m3", "fp4_e2m1", "int4"]

# test_corners_mx.py:60-79: int4 never emits -2.0 (block_size=0, axes=1, nearest) ; both modes
MX_INT4_X = [[-1.0, -1.75, -1.99, 1.99], [1.0, -1.75, -1.99, 1.99]]
MX_INT4_Y = [[-1.0, -1.75, -1.75, 1.75], [1.0, -1.75, -1.75, 1.75]]

# test_corners_mx.py:82-124: int8, block 10, axes=1, nearest ; spec mode only
MX_HW_X = [
    [1.0] * 10,
    [1.0] * 5 + [2.0] * 5,
    [-1.0] * 5 + [-2.0] * 5,
    [1.0] * 5 + [-2.0] * 5,
    [1.015625, 1.0234375, 1.03125, 1.0390625, 1.25, 1.2578125, 1.9375, 1.9453125, 1.984375, 1.9921875],
    [-1.984375, -1.9765625, -1.96875, -1.9609375, -1.9375, -1.9296875, -1.75, -1.7421875, -1.0, -1.9921875],
    [1.99609375, 1.98828125, 0.0, 0.00390625, 0.0078125, 0.01171875, -0.015625, -0.01171875, -0.0078125, -0.00390625],
]
MX_HW_Y = [
    [1.0] * 10,
    [1.0] * 5 + [2.0] * 5,
    [-1.0] * 5 + [-2.0] * 5,
    [1.0] * 5 + [-2.0] * 5,
    [1.015625, 1.03125, 1.03125, 1.046875, 1.25, 1.265625, 1.9375, 1.953125, 1.984375, 1.984375],
    [-1.984375, -1.984375, -1.96875, -1.96875, -1.9375, -1.9375, -1.75, -1.75, -1.0, -1.984375],
    [1.984375, 1.984375, 0.0, 0.0, 0.015625, 0.015625, -0.015625, -0.015625, -0.015625, 0.0],
]

# test_e5m0_scale.py:23-27: scale_bits=5, int2 elements, block = row of 8 ; spec mode only
E5M0_HEADS = [2 ** 16, 2 ** 15, 2 ** -16, 2 ** -17, 2 ** -15, 2 ** -14, 1, 0]
E5M0_TARGET_HEADS = [INF, 2 ** 15, 2 ** -15, 0, 2 ** -15, 2 ** -14, 1, 0]   # INF row = whole row non-finite
E5M0_X = [[h] + [0] * 7 for h in E5M0_HEADS]
E5M0_T = [[t] + [INF if t == INF else 0] * 7 for t in E5M0_TARGET_HEADS]

# test_fp8_e4m3_fix.py:16-17: 448 saturation, block = whole vector, scale_bits 8 ; spec mode only
FP8_E4M3_X = [485, 475, 400, -485, -475, -450, 2 ** -6, -(2 ** -6)]
FP8_E4M3_T = [448, 448, 416, -448, -448, -448, 2 ** -6, -(2 ** -6)]

# test_fp8_e4m3_fix.py:91-119: MXFP8 rounding table, block 32, axes=0, round nearest AND even,
# custom_cuda False AND True (i.e. both modes)
MXFP8_X = [-0.582733273506, -0.256973713636, 0.506033003330, 0.400039970875,
           -0.437093466520, 1.275019764900, -2.123294353485, 1.514909625053,
           -2.660086154938, -0.200791791081, -0.060985822231, -0.209203109145,
           2.385987281799, 0.062245476991, 0.217003762722, -0.857734560966,
           0.507945835590, 0.896152675152, -0.751049160957, -0.488164335489,
           0.805381953716, -0.172028362751, 0.271137863398, -0.503807783127,
           1.879478693008, -0.294227510691, 0.968807995319, 0.670037031174,
           -0.871595799923, 0.304561793804, -0.567594051361, -1.265962004662]
MXFP8_T = [-0.5625, -0.25, 0.5, 0.40625, -0.4375, 1.25, -2.0, 1.5,
           -2.75, -0.203125, -0.0625, -0.203125, 2.5, 0.0625, 0.21875, -0.875,
           0.5, 0.875, -0.75, -0.5, 0.8125, -0.171875, 0.28125, -0.5,
           1.875, -0.28125, 1.0, 0.6875, -0.875, 0.3125, -0.5625, -1.25]

# test_quantize_elemwise.py:51-58: bfloat16, round even, tol 1e-8
BF16_RNE_X = [65535., -65535., -1.9985847, 0.4999000132083893, -0.49968934059143066,
              -0.0018959, 0.0018959, -2.52075195e-02, 0.0156078, 0.00780624,
              -0.00048747, -0.01561374, -0.00780595, -0.0039039, 0.01557725,
              -0.0077924, 0.00389144, -0.00779678, -0.01556631, -0.00194835]
BF16_RNE_T = [65536., -65536., -2.0, 0.5, -0.5,
              -0.00189209, 0.00189209, -2.51464844e-02, 0.015625, 0.0078125,
              -0.00048828, -0.015625, -0.0078125, -0.00390625, 0.01556396,
              -0.00778198, 0.00389099, -0.00778198, -0.01556396, -0.0019455]

# test_corners_elemwise.py:28-52: bf16 max_norm / rounds-to-max / rounds-to-Inf (nearest)
BF16_MAX_X = [2 ** 127 * (2 ** 8 - 1) / 2 ** 7, -2 ** 127 * (2 ** 8 - 1) / 2 ** 7,
              2 ** 127 * (2 ** 10 - 3) / 2 ** 9, -2 ** 127 * (2 ** 10 - 3) / 2 ** 9,
              2 ** 127 * (2 ** 10 - 1) / 2 ** 9, -2 ** 127 * (2 ** 10 - 1) / 2 ** 9]
BF16_MAX_FINITE = [True, True, True, True, False, False]
# test_corners_elemwise.py:55-85: fp16 (mantissa_bits=10, exp_bits=5)
FP16_MAX_X = [65504., -65504., 65519., -65519., 65520., -65520.]
FP16_MAX_FINITE = [True, True, True, True, False, False]
# test_corners_elemwise.py:88-116: bfloat/fp keep NaN and Inf where they are
ELEM_NAN_X = [0., NAN, INF, 0.]


def all_encodings(e, m, encodes_infs=True):
    """every value of an eXmY format (common_lib.py:142-180), as a list of python floats"""
    import numpy as np
    x = np.zeros(2 ** (e + m + 1), dtype=np.float32)
    for i in range(2 ** (e + m)):
        if e > 0:
            ex = i >> m
            if encodes_infs and ex == 2 ** e - 1:
                continue
            if ex == 0:
                ex, lead = 1 - (2 ** (e - 1) - 1), 0.0
            else:
                ex, lead = ex - (2 ** (e - 1) - 1), 1.0
            frac = (i & (2 ** m - 1)) / 2 ** m
            v = 0 if (e == 4 and m == 3 and ex == 8 and frac == 0.875) else 2.0 ** ex * (lead + frac)
        else:
            v = i / 2 ** (m - 1)
        x[i] = v
        x[i + 2 ** (e + m)] = -v
    return x
