// msq_common.cuh -- shared device helpers for the MicroScopiQ B200 kernels.
//
// Arithmetic contract (what "bit-exact" is measured against): the reference's
// PyTorch path, utils/quant.py:23-266 + number_system/mx/elemwise_ops.py:84-174,
// restated in oracle/msq_oracle.py.  Every multiply/add that the reference does
// as a separate fp32 op is a separate fp32 op here (this translation unit is
// compiled with -fmad=false; fused ops are only ever requested explicitly).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MSQ_OK 0
#define MSQ_ERR_BAD_ARG (-1)
#define MSQ_ERR_UNSUPPORTED (-2)
#define MSQ_ERR_CUDA (-3)
#define MSQ_ERR_ALIGN (-4)

// status word bits written by the quantizers (device int32, OR-accumulated)
#define MSQ_ST_NONFINITE_INPUT 1   // NaN/Inf in a block  (reference: assert quant.py:479/103)
#define MSQ_ST_EXP_IN_OVERFLOW 2   // inlier shared exponent above the scale range (quant.py:88)
#define MSQ_ST_EXP_OUT_OVERFLOW 4  // outlier shared exponent above the scale range (quant.py:119)

#define MSQ_SEG_BLOCKS 1024        // blocks per packed segment (layout constant)
#define MSQ_EXP_NAN (-128)         // int8 marker for a NaN shared exponent

namespace msq {

enum Round { RD_NEAREST = 0, RD_FLOOR = 1, RD_EVEN = 2 };   // formats.py:15-18

// element format as the kernels see it (formats.py:65-129 digested on the host)
struct Fmt {
  int ebits;      // 0 => integer format
  int bits;       // "mbits" of the reference: sign + implicit + trailing mantissa
  int emax;       // exponent of max_norm
  int min_exp;    // 2 - 2^(ebits-1)
  int width;      // storage container width of one code: 2, 4 or 8
  float max_norm;
  float up;       // 2^(bits-2)
  float down;     // 2^-(bits-2)
  float L, H;     // float-domain exponent clamp of the element rounding: 2^min_exp, 2^127 (ints: 1, 1)
};

__host__ inline Fmt make_fmt(int ebits, int mbits, int emax, float max_norm) {
  Fmt f;
  f.ebits = ebits; f.bits = mbits; f.emax = emax; f.max_norm = max_norm;
  f.min_exp = ebits > 0 ? 2 - (1 << (ebits - 1)) : 0;
  int cw = ebits > 0 ? 1 + ebits + (mbits - 2) : mbits;   // code width
  f.width = cw <= 2 ? 2 : (cw <= 4 ? 4 : 8);
  f.up = 1.0f; f.down = 1.0f;
  for (int i = 0; i < mbits - 2; ++i) { f.up *= 2.0f; f.down *= 0.5f; }
  f.L = 1.0f; f.H = 1.0f;
  if (ebits > 0) {
    for (int i = 0; i < -f.min_exp; ++i) f.L *= 0.5f;
    for (int i = 0; i < f.min_exp; ++i) f.L *= 2.0f;
    f.H = 1.70141183e38f;      // 2^127
  }
  return f;
}

// exact 2^e for e in [-149, 127]
__device__ __forceinline__ float pow2i(int e) {
  int b = e + 127;
  return __int_as_float(b > 0 ? (b << 23) : (0x00400000 >> (-b)));
}

// exact floor(log2(a)) for finite a > 0 (subnormals included)
__device__ __forceinline__ int floor_log2f(float a) {
  int u = __float_as_int(a);
  int be = (u >> 23) & 0xFF;
  if (be != 0) return be - 127;
  return 31 - __clz(u & 0x7FFFFF) - 149;
}

// _round_mantissa (elemwise_ops.py:47-81) on |y|; sign applied by caller
template <int RMODE>
__device__ __forceinline__ float round_mag(float a) {
  if (RMODE == RD_FLOOR) return floorf(a);
  float r = floorf(__fadd_rn(a, 0.5f));
  if (RMODE == RD_EVEN) {
    float t = __fadd_rn(a, -0.5f);
    float h = __fmul_rn(t, 0.5f);
    if (t >= 0.0f && h == floorf(h)) r = __fadd_rn(r, -1.0f);
  }
  return r;
}

// _quantize_elemwise_core, float formulation (elemwise_ops.py:131-167).
template <int RMODE>
__device__ __forceinline__ float qelem(float x, int ebits, int min_exp, float up, float down,
                                       float max_norm, bool saturate, bool allow_denorm) {
  float ax = fabsf(x);
  float out = x;
  float y, p2 = 1.0f;
  if (ebits != 0) {
    int be = (__float_as_int(ax) >> 23);
    int p = max(be - 127, min_exp);
    if (!allow_denorm && !(ax >= pow2i(min_exp))) out = __fmul_rn(0.0f, x);
    p2 = pow2i(min(p, 127));
    y = __fmul_rn(__fmul_rn(out, pow2i(-min(p, 127))), up);
  } else {
    y = __fmul_rn(out, up);
  }
  float r = copysignf(round_mag<RMODE>(fabsf(y)), y);
  y = __fmul_rn(r, down);
  if (ebits != 0) y = __fmul_rn(y, p2);
  if (saturate || ebits == 0) {
    y = (y > max_norm) ? max_norm : y;      // comparisons keep NaN, like torch.clamp
    y = (y < -max_norm) ? -max_norm : y;
  } else if (fabsf(y) > max_norm) {
    y = copysignf(__int_as_float(0x7F800000), y);
  }
  if (ax == __int_as_float(0x7F800000)) y = x;
  return y;
}

template <int RMODE>
__device__ __forceinline__ float qelem(float x, const Fmt& f, bool saturate, bool allow_denorm) {
  return qelem<RMODE>(x, f.ebits, f.min_exp, f.up, f.down, f.max_norm, saturate, allow_denorm);
}

// value on the format grid (element units) -> sign-magnitude code (Appendix B)
__device__ __forceinline__ uint32_t encode_elem(float q, int ebits, int bits) {
  float a = fabsf(q);
  uint32_t sign = (a != 0.0f) ? ((uint32_t)__float_as_int(q) >> 31) : 0u;
  uint32_t mag;
  int cw;
  if (ebits == 0) {
    mag = (uint32_t)__float2int_rz(__fmul_rn(a, pow2i(bits - 2)));
    cw = bits;
  } else {
    int tm = bits - 2;
    int bias = (1 << (ebits - 1)) - 1;
    int u = __float_as_int(a);
    int be = ((u >> 23) & 0xFF) - 127 + bias;
    if (be >= 1) mag = ((uint32_t)be << tm) | ((uint32_t)(u & 0x7FFFFF) >> (23 - tm));
    else mag = (uint32_t)__float2int_rz(__fmul_rn(a, pow2i(tm - 1 + bias)));
    cw = 1 + ebits + tm;
  }
  return mag | (sign << (cw - 1));
}

__device__ __forceinline__ float decode_elem(uint32_t code, int ebits, int bits) {
  if (ebits == 0) {
    uint32_t mag = code & ((1u << (bits - 1)) - 1u);
    float v = __fmul_rn((float)mag, pow2i(-(bits - 2)));
    return ((code >> (bits - 1)) & 1u) ? -v : v;
  }
  int tm = bits - 2;
  int cw = 1 + ebits + tm;
  int bias = (1 << (ebits - 1)) - 1;
  uint32_t ef = (code >> tm) & ((1u << ebits) - 1u);
  uint32_t mf = code & ((1u << tm) - 1u);
  float v;
  if (ef > 0) v = __int_as_float((int)(((ef - bias + 127) << 23) | (mf << (23 - tm))));
  else v = __fmul_rn((float)mf, pow2i(1 - bias - tm));
  return ((code >> (cw - 1)) & 1u) ? -v : v;
}


// ---- fast element encoder --------------------------------------------------------------------
// For a value q that already lies on the format grid, the sign-magnitude code is a shifted slice of
// the fp32 bits of |q| (floats, normal range), of |q| + min_norm (floats, subnormal range) or of
// |q| + 2 (integer formats):   mag = (bits(|q| + add) >> sh) - bias_const.
struct EncC {
  float thresh;   // |q| < thresh -> "small" path (floats: min_norm; ints: +inf)
  float addc;     // added on the small path (floats: min_norm; ints: 2.0)
  int sh;         // 23 - trailing mantissa bits of the code
  int c_norm;     // subtracted on the normal path
  int c_sub;      // subtracted on the small path
  int signbit;    // 1 << (code width - 1)
};
__host__ __device__ inline EncC make_enc(int ebits, int bits) {
  EncC e;
  if (ebits == 0) {
    e.thresh = 3.0e38f; e.addc = 2.0f; e.sh = 23 - (bits - 1); e.c_norm = 0; e.c_sub = 128 << (bits - 1);
    e.signbit = 1 << (bits - 1);
  } else {
    const int tm = bits - 2, bias = (1 << (ebits - 1)) - 1;
    float mn = 1.0f;
    for (int i = 0; i < bias - 1; ++i) mn *= 0.5f;        // 2^(1-bias)
    e.thresh = mn; e.addc = mn; e.sh = 23 - tm; e.c_norm = (127 - bias) << tm; e.c_sub = (128 - bias) << tm;
    e.signbit = 1 << (ebits + tm);
  }
  return e;
}
__device__ __forceinline__ uint32_t encode_fast(float q, const EncC& e) {
  const float aq = fabsf(q);
  const bool small = aq < e.thresh;
  const int u = __float_as_int(__fadd_rn(aq, small ? e.addc : 0.0f));
  const int mag = (u >> e.sh) - (small ? e.c_sub : e.c_norm);
  return (uint32_t)mag | ((q < 0.0f) ? (uint32_t)e.signbit : 0u);
}

// shared-exponent clamp (quant.py:85-89 / mx_ops.py:270-273).  Returns MSQ_EXP_NAN on overflow.
__device__ __forceinline__ int clamp_scale_exp(int e, int scale_bits, bool under_to_minus20) {
  int S = (1 << (scale_bits - 1)) - 1;
  if (e > S) return MSQ_EXP_NAN;
  if (e < -S) return (under_to_minus20 && S > 20) ? -20 : -S;
  return e;
}

__device__ __forceinline__ float qnan() { return __int_as_float(0x7FC00000); }

}  // namespace msq
