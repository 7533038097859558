// msq_quant_core.cuh -- per-block arithmetic of the MX(+outlier) quantizers.
//
// A block of `bs` values is held by LPB cooperating lanes, N values each
// (bs == N * LPB).  Reductions are in-thread over N then xor-butterflies over
// the LPB lanes (commutative => every lane ends with the identical value).
//
// core_outlier : utils/quant.py:23-266 (variant 0) and
//                number_system/mx/mx_ops.py:210-330 (variant 1, bounds supplied)
// core_plain   : number_system/mx/mx_ops.py:426-451 (mode 0, incl. the +1e-6 of :444)
//                and cpp/mx.cuh:37-52 + cpp/shared_exp.cuh:14-53 (mode 1)
#pragma once
#include "msq_common.cuh"

namespace msq {

struct QCfg {
  Fmt fin, fout;
  int sb_in, sb_out;
  float std_dev;
  int flush;
  int variant;   // outlier: 0 = utils/quant (under -> -20), 1 = mx_ops (bounds external, under -> -S)
  int mode;      // plain  : 0 = python (+1e-6), 1 = spec/native
  int rmode;     // Round
};

template <int LPB> __device__ __forceinline__ double grp_sum(double v) {
#pragma unroll
  for (int o = LPB / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPB> __device__ __forceinline__ float grp_max(float v) {
#pragma unroll
  for (int o = LPB / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPB> __device__ __forceinline__ int grp_add(int v) {
#pragma unroll
  for (int o = LPB / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPB> __device__ __forceinline__ int grp_or(int v) {
#pragma unroll
  for (int o = LPB / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Rounding selected at run time without duplicating the element code:
// addend = 0.5 (nearest / even) or 0 (floor); `even` subtracts one on exact .5 ties
// whose floor is even (elemwise_ops.py:59-72).
struct RoundSel { float addend; bool even; };
__device__ __forceinline__ RoundSel make_round(int rmode) {
  RoundSel r; r.addend = (rmode == RD_FLOOR) ? 0.0f : 0.5f; r.even = (rmode == RD_EVEN); return r;
}

// _quantize_elemwise_core with saturate_normals=True, allow_denorm=True (the only way the MX
// quantizers call it, quant.py:96-99), format chosen per element by `isf/min_exp/up/down/mx`.
__device__ __forceinline__ float qelem_mx(float x, bool isf, int min_exp, float up, float down, float mx,
                                          const RoundSel rs) {
  const float ax = fabsf(x);
  const int be = __float_as_int(ax) >> 23;
  const int p = isf ? min(max(be - 127, min_exp), 127) : 0;
  float y = __fmul_rn(__fmul_rn(x, pow2i(-p)), up);
  const float a = fabsf(y);
  float r = floorf(__fadd_rn(a, rs.addend));
  if (rs.even) {
    const float t = __fadd_rn(a, -0.5f);
    const float h = __fmul_rn(t, 0.5f);
    if (t >= 0.0f && h == floorf(h)) r = __fadd_rn(r, -1.0f);
  }
  y = __fmul_rn(__fmul_rn(copysignf(r, y), down), pow2i(p));
  y = (y > mx) ? mx : y;
  y = (y < -mx) ? -mx : y;
  return (ax == __int_as_float(0x7F800000)) ? x : y;
}

__device__ __forceinline__ float qelem_rt(float v, const Fmt& f, int rmode) {
  return qelem_mx(v, f.ebits != 0, f.min_exp, f.up, f.down, f.max_norm, make_round(rmode));
}

// Same arithmetic for FINITE inputs with |x| < 2^127 (true for every scaled block element:
// |x / 2^e| <= 2^(emax+1)), which lets both powers of two be built with one integer op each
// and the clamp use min/max.
__device__ __forceinline__ float qelem_mx_bounded(float x, bool isf, int min_exp, float up, float down, float mx,
                                                  const RoundSel rs) {
  const int be = (__float_as_int(x) >> 23) & 0xFF;
  const int p = isf ? max(be - 127, min_exp) : 0;
  float y = __fmul_rn(__fmul_rn(x, __int_as_float((127 - p) << 23)), up);
  const float a = fabsf(y);
  float r = floorf(__fadd_rn(a, rs.addend));
  if (rs.even) {
    const float t = __fadd_rn(a, -0.5f);
    const float h = __fmul_rn(t, 0.5f);
    if (t >= 0.0f && h == floorf(h)) r = __fadd_rn(r, -1.0f);
  }
  y = __fmul_rn(__fmul_rn(copysignf(r, y), down), __int_as_float((p + 127) << 23));
  return fminf(fmaxf(y, -mx), mx);
}

struct BlockOut {
  uint32_t mask;   // bit i set: x[i] is an outlier
  int e_in, e_out; // clamped shared exponents (MSQ_EXP_NAN on overflow)
  int cnt;         // outliers in the whole block (all lanes)
  int st;          // status bits raised by this block
};

// x[] in: block values (padding = 0).  x[] out: dequantized values.
// code[] out (if CODES): element codes, inlier or outlier format according to mask.
// EXT: detector bounds come from lo_e/hi_e (variant 1) instead of block statistics.
template <int N, int LPB, bool CODES, bool EXT>
__device__ __forceinline__ BlockOut core_outlier(float (&x)[N], uint32_t (&code)[N], const QCfg& c,
                                                 const float* lo_e, const float* hi_e) {
  constexpr int BS = N * LPB;
  BlockOut o;
  o.st = 0;
  int bad = 0;
  float l = 0.0f, h = 0.0f;
  if (!EXT) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float a = fabsf(x[i]);
      bad |= !(a <= 3.402823466e38f);
      s += (double)a;
    }
    s = grp_sum<LPB>(s);
    const double mean64 = s / (double)BS;
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double d = (double)fabsf(x[i]) - mean64;
      q = fma(d, d, q);
    }
    q = grp_sum<LPB>(q);
    const float mean = (float)mean64;
    const float sd = (float)sqrt(q / (double)BS);
    const float w = __fmul_rn(c.std_dev, sd);
    l = __fsub_rn(mean, w); h = __fadd_rn(mean, w);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) bad |= !(fabsf(x[i]) <= 3.402823466e38f);
  }
  if (LPB > 1) bad = grp_or<LPB>(bad);

  // outlier split.  A*(1-mask) / A*mask of quant.py:70-71 are exact products with 0/1, so the
  // maxima below are the reference's maxima.  ALU-pipe work is kept minimal: 0/1 blending and
  // exponent arithmetic run on the FMA pipe (FMUL / IMAD).
  uint32_t mask = 0;
  float m_in = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float li = EXT ? lo_e[i] : l, hi_i = EXT ? hi_e[i] : h;
    const bool isout = (x[i] < li) | (x[i] > hi_i);
    mask |= (isout ? 1u : 0u) << i;
    m_in = fmaxf(m_in, __fmul_rn(fabsf(x[i]), isout ? 0.0f : 1.0f));
  }
  o.mask = mask;
  o.cnt = grp_add<LPB>(__popc(mask));
  m_in = grp_max<LPB>(m_in);

  const int E_in = (m_in == 0.0f) ? -126 : floor_log2f(m_in);
  const bool flush = c.flush && (E_in <= -127);
  o.e_in = clamp_scale_exp(E_in - c.fin.emax, c.sb_in, c.variant == 0);
  if (bad) o.st |= MSQ_ST_NONFINITE_INPUT;
  if (o.e_in == MSQ_EXP_NAN) o.st |= MSQ_ST_EXP_IN_OVERFLOW;
  const bool dead_in = bad || o.e_in == MSQ_EXP_NAN;
  const int e_in = dead_in ? 0 : o.e_in;
  const float s_in = pow2i(e_in), inv_s_in = pow2i(-e_in);

  float m_out = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float t = fabsf(__fmul_rn(x[i], s_in));
    m_out = fmaxf(m_out, __fmul_rn(t, ((mask >> i) & 1u) ? 1.0f : 0.0f));
  }
  m_out = grp_max<LPB>(m_out);
  const int E_out = (m_out == 0.0f) ? -126 : floor_log2f(m_out);
  o.e_out = dead_in ? MSQ_EXP_NAN : clamp_scale_exp(E_out - c.fout.emax, c.sb_out, c.variant == 0);
  if (!dead_in && o.e_out == MSQ_EXP_NAN) o.st |= MSQ_ST_EXP_OUT_OVERFLOW;
  const bool dead = dead_in || o.e_out == MSQ_EXP_NAN;
  const int e_out = dead ? 0 : o.e_out;
  const float s_out = pow2i(e_out), inv_s_out = pow2i(-e_out);
  const RoundSel rs = make_round(c.rmode);
  const float inv_s_in_eff = flush ? 0.0f : inv_s_in;            // quant.py:80 (x*0 keeps the sign of x)

  // Element rounding with the quantum exponent p clamped in the BIASED domain:
  //   p' = clamp(biased_exp(v), lo', hi')   (floats: lo' = min_exp+127, hi' = 254; ints: 127, 127)
  //   y  = v * 2^(mant - p)  ->  round  ->  * 2^(p - mant)
  // (v * 2^-p) * 2^mant == v * 2^(mant-p) exactly here: |v| <= 2^(emax+1) and every element
  // format on this path has ebits <= 5, so no intermediate leaves the normal range inexactly.
  const int mant_i = c.fin.bits - 2, mant_o = c.fout.bits - 2;
  const int lo_i = c.fin.ebits ? c.fin.min_exp + 127 : 127, hi_i = c.fin.ebits ? 254 : 127;
  const int lo_o = c.fout.ebits ? c.fout.min_exp + 127 : 127, hi_o = c.fout.ebits ? 254 : 127;
  const int d_lo = lo_o - lo_i, d_hi = hi_o - hi_i;
  const int c1_i = (254 + mant_i) << 23, d_c1 = (mant_o - mant_i) << 23;   // s1 = c1 - p'<<23
  const int c2_i = -(mant_i << 23), d_c2 = -((mant_o - mant_i) << 23);     // s2 = c2 + p'<<23
  const float mx_i = c.fin.max_norm, mx_o = c.fout.max_norm;

#pragma unroll
  for (int i = 0; i < N; ++i) {
    const bool isout = (mask >> i) & 1u;
    const int mi = isout ? 1 : 0;
    // inlier : v = x * 2^-e_in               ; r = q * 2^e_in                  quant.py:92,102
    // outlier: v = (x * 2^e_in) * 2^-e_out   ; r = (q * 2^e_out) * 2^-e_in     quant.py:94,125,136
    const float vin = __fmul_rn(x[i], inv_s_in_eff);
    const float vout = __fmul_rn(__fmul_rn(x[i], s_in), inv_s_out);
    const float v = isout ? vout : vin;
    const int be = (__float_as_int(v) >> 23) & 0xFF;
    const int pb = min(max(be, lo_i + mi * d_lo), hi_i + mi * d_hi);
    const float s1 = __int_as_float((c1_i + mi * d_c1) - (pb << 23));
    const float s2 = __int_as_float((c2_i + mi * d_c2) + (pb << 23));
    const float y = __fmul_rn(v, s1);
    const float a = fabsf(y);
    float r = floorf(__fadd_rn(a, rs.addend));
    if (rs.even) {
      const float t = __fadd_rn(a, -0.5f);
      const float hh = __fmul_rn(t, 0.5f);
      if (t >= 0.0f && hh == floorf(hh)) r = __fadd_rn(r, -1.0f);
    }
    const float mx = isout ? mx_o : mx_i;
    const float qv = fminf(fmaxf(__fmul_rn(copysignf(r, y), s2), -mx), mx);
    if (CODES) code[i] = encode_elem(qv, isout ? c.fout.ebits : c.fin.ebits, isout ? c.fout.bits : c.fin.bits);
    const float rin = __fmul_rn(qv, s_in);
    const float rout = __fmul_rn(__fmul_rn(qv, s_out), inv_s_in);
    x[i] = __fadd_rn(isout ? rout : rin, 0.0f);                                 // quant.py:140 (x + 0)
  }
  if (dead) {      // NaN/Inf input or a shared exponent above the scale range: the reference asserts
#pragma unroll
    for (int i = 0; i < N; ++i) { x[i] = qnan(); if (CODES) code[i] = 0u; }
  }
  return o;
}

// cpp/quantize.cuh:88-149 restated: integer-mantissa rounding of one fp32 value.
__device__ __forceinline__ float qelem_native(float x, int bits, int ebits, float max_norm, int rmode,
                                              bool saturate, bool allow_denorm) {
  const uint32_t u = (uint32_t)__float_as_int(x);
  const uint32_t sign = u >> 31;
  int be = (int)((u >> 23) & 0xFFu);
  int man = (int)(u & 0x7FFFFFu);
  const int mb = bits - 1;                       // magnitude bits incl. implicit one
  const bool is_int = ebits == 0;
  const int nbias = is_int ? 1 : (1 << (ebits - 1)) - 1;
  const int nbe = be - 127 + nbias;
  if (!is_int && !allow_denorm && nbe < 1) return 0.0f;
  int ediff = nbe <= 0 ? 1 - nbe : 0;
  ediff = ediff > 24 ? 24 : ediff;
  const bool sub = be == 0;
  const int sig = sub ? 23 : 24;
  if (!sub) man += (1 << 23);
  bool tie = false, ev = false;
  if (rmode == RD_EVEN) {
    const int tb = ediff + (sig - mb);
    tie = (man & ((1 << (tb - 1)) - 1)) == 0;
    ev = (man & (1 << tb)) == 0;
  }
  man >>= ediff;
  man >>= (sig - mb - 1);
  if (rmode != RD_FLOOR && (!is_int || man != ((1 << (mb + 1)) - 1)) && !(tie && ev)) man += 1;
  man >>= 1;
  if (man == 0) return 0.0f;
  man <<= (sig - mb + ediff);
  const bool ovf = man >= (1 << sig);
  if (ovf && !sub) man >>= 1;
  man &= 0x7FFFFF;
  if (ovf) be += 1;
  float out = __int_as_float((int)((sign << 31) | ((uint32_t)be << 23) | (uint32_t)man));
  if (fabsf(out) > max_norm) {
    if (is_int || saturate) out = sign ? -max_norm : max_norm;
    else out = __int_as_float((int)((sign << 31) | 0x7F800000u));
  }
  return out;
}

struct PlainOut { int e; int st; };

template <int N, int LPB>
__device__ __forceinline__ PlainOut core_plain(float (&x)[N], uint32_t (&code)[N], bool want_codes,
                                               const QCfg& c, const float* ext_max) {
  PlainOut o; o.st = 0;
  int bad = 0;
  float m = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) { bad |= !(fabsf(x[i]) <= 3.402823466e38f); m = fmaxf(m, fabsf(x[i])); }
  bad = grp_or<LPB>(bad);
  m = grp_max<LPB>(m);
  if (ext_max) { m = fabsf(*ext_max); bad = !(m <= 3.402823466e38f); }
  const Fmt& f = c.fin;
  if (c.mode == 0) {
    const int E = (m == 0.0f) ? -126 : floor_log2f(m);
    const bool flush = c.flush && (E <= -127);
    int e = bad ? MSQ_EXP_NAN : clamp_scale_exp(E - f.emax, c.sb_in, false);
    o.e = e;
    if (e == MSQ_EXP_NAN) {
      o.st |= bad ? MSQ_ST_NONFINITE_INPUT : MSQ_ST_EXP_IN_OVERFLOW;
#pragma unroll
      for (int i = 0; i < N; ++i) { x[i] = qnan(); code[i] = 0; }
      return o;
    }
    const float s = pow2i(e);
    const float den = __fadd_rn(s, 1e-6f);                       // mx_ops.py:444
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float v = __fdiv_rn(flush ? __fmul_rn(x[i], 0.0f) : x[i], den);
      float qv = qelem_rt(v, f, c.rmode);
      if (want_codes) code[i] = encode_elem(qv, f.ebits, f.bits);
      x[i] = __fmul_rn(qv, s);
    }
  } else {
    int be = bad ? 255 : ((__float_as_int(m) >> 23) & 0xFF);
    const bool flush = (be == 0) && c.flush;
    int sh = (be != 255) ? be - f.emax : 255;                     // shared_exp.cuh:39-41
    const int S = c.sb_in != 0 ? (1 << (c.sb_in - 1)) - 1 : 255;
    const int ub = sh - 127;
    sh = ub > S ? 255 : sh;
    sh = ub < -S ? 127 - S : sh;
    const int mant = (sh == 0 || sh == 255) ? 0x400000 : 0;
    const float s = __int_as_float((sh << 23) | mant);
    o.e = sh == 255 ? MSQ_EXP_NAN : sh - 127;
    if (sh == 255) o.st |= bad ? MSQ_ST_NONFINITE_INPUT : MSQ_ST_EXP_IN_OVERFLOW;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float v = flush ? 0.0f : __fdiv_rn(x[i], s);
      float qv = qelem_native(v, f.bits, f.ebits, f.max_norm, c.rmode, true, true);
      if (want_codes) code[i] = (qv == qv) ? encode_elem(qv, f.ebits, f.bits) : 0;
      x[i] = __fmul_rn(qv, s);
    }
  }
  return o;
}

}  // namespace msq
