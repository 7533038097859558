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
#include <string.h>

namespace msq {

struct QCfg {
  Fmt fin, fout;
  EncC enc_i, enc_o;   // code encoders of fin / fout, filled by finish_cfg() on the host
  int sb_in, sb_out;
  float std_dev;
  int flush;
  int variant;   // outlier: 0 = utils/quant (under -> -20), 1 = mx_ops (bounds external, under -> -S)
  int mode;      // plain  : 0 = python (+1e-6), 1 = spec/native
  int rmode;     // Round
  float fzero;   // 0.0f the compiler cannot see (finish_cfg): addend that keeps a product an FFMA of its own
  // float-domain encoder (core_outlier, packed path): code = (bits(|q| * sc) | sign bit at tb) >> (bits(ka) & 31).
  // Each constant is chosen so that an exact 0/1 FFMA blend of the inlier and the outlier value yields the element's
  // own constant as a BIT PATTERN (enc2_consts)
  float sc_i, sc_o, ka_i, ka_o, tb_i, tb_o;
  int enc2_ok;   // both formats representable (code width <= 8, ebits <= 5)
  // hi-field accumulation of the packing kernels (HI template flag of core_outlier): bit patterns 1 (inlier) and
  // 1 << hi_w (outlier) blended per element like the constants above; hi_mulhi = 2^(32 - slot width) turns
  // "code >> slot width" into a multiply-high
  float mu_i, mu_o;
  unsigned int hi_mulhi;
  // core_outlier_cvt: |shared exponent| limits of the conversion path (min(40, S_in) / min(80, S_out)) and the clamped
  // exponents the reference gives a block whose inlier / outlier maximum is zero
  int cvt_lim_in, cvt_lim_out, e_zero_in, e_zero_out;
};

// |q| on the format grid times sc lands in fp32's own subnormal / lowest-normal range in such a way that the fp32 bit
// pattern, shifted right by ksh2, IS the magnitude code: floats: sc = 2^(-126 - min_exp) maps the format's smallest
// normal onto fp32's (exponent field 1, subnormals onto subnormals), ksh2 = 23 - trailing mantissa bits; integers:
// sc = 2^-127 puts k = |q| / 2^-(bits-2) into the mantissa at bit 24 - bits.  The sign goes to bit ksh2 + width - 1 of the
// unshifted pattern (an exponent bit that is always zero there).
__host__ inline bool enc2_consts(const Fmt& f, float* sc, float* ka, float* tb) {
  int ksh2, tbit, e2;
  if (f.ebits == 0) {
    if (f.bits < 2 || f.bits > 8) return false;
    ksh2 = 24 - f.bits; tbit = 23; e2 = -127;
  } else {
    const int tm = f.bits - 2;
    if (f.ebits > 5 || 1 + f.ebits + tm > 8 || tm < 0) return false;
    ksh2 = 23 - tm; tbit = 23 + f.ebits; e2 = -126 - f.min_exp;
  }
  float v = 1.0f;
  for (int i = 0; i < -e2; ++i) v *= 0.5f;
  *sc = v;
  *ka = 8388608.0f + (float)ksh2;
  const uint32_t tb_bits = 1u << tbit;
  memcpy(tb, &tb_bits, 4);
  return true;
}
__host__ inline void finish_cfg(QCfg& c) {
  c.enc_i = make_enc(c.fin.ebits, c.fin.bits);
  c.enc_o = make_enc(c.fout.ebits, c.fout.bits);
  c.fzero = 0.0f;
  const bool a = enc2_consts(c.fin, &c.sc_i, &c.ka_i, &c.tb_i), b = enc2_consts(c.fout, &c.sc_o, &c.ka_o, &c.tb_o);
  c.enc2_ok = (a && b) ? 1 : 0;
  {
    const int d = c.fout.width - c.fin.width;
    const int hi_w = d <= 0 ? 0 : (d <= 2 ? 2 : (d <= 4 ? 4 : 8));
    const uint32_t one = 1u, mo = 1u << hi_w;
    memcpy(&c.mu_i, &one, 4);
    memcpy(&c.mu_o, &mo, 4);
    c.hi_mulhi = 1u << (32 - c.fin.width);
  }
  {
    // floor_log2_or_min(0) = -126; clamp_scale_sel(e, S, U): e < -S -> U, e > S -> NaN marker
    const int S_in = (1 << (c.sb_in - 1)) - 1, S_out = (1 << (c.sb_out - 1)) - 1;
    const int U_in = (c.variant == 0 && S_in > 20) ? -20 : -S_in, U_out = (c.variant == 0 && S_out > 20) ? -20 : -S_out;
    c.cvt_lim_in = S_in < 40 ? S_in : 40;
    c.cvt_lim_out = S_out < 80 ? S_out : 80;
    const int z_in = -126 - c.fin.emax, z_out = -126 - c.fout.emax;
    c.e_zero_in = z_in < -S_in ? U_in : z_in;
    c.e_zero_out = z_out < -S_out ? U_out : z_out;
  }
}

// fp32 sum over the LPB lanes in canonical tree order: partner distance 1, 2, 4, ... continues the
// in-thread adjacent pairwise tree (commutativity makes every lane end with the same value)
template <int LPB> __device__ __forceinline__ float grp_sum_tree(float v) {
#pragma unroll
  for (int o = 1; o < LPB; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// in-thread adjacent pairwise tree over N (power of two) values; destroys t[]
template <int N, int START = 1> __device__ __forceinline__ float tree_sum_regs(float (&t)[N]) {
#pragma unroll
  for (int stride = START; stride < N; stride <<= 1)
#pragma unroll
    for (int i = 0; i + stride < N; i += 2 * stride) t[i] = __fadd_rn(t[i], t[i + stride]);
  return t[0];
}
// the same adjacent pairwise tree with two independent additions per packed FADD2:
// level `s` adds t[i+s] into t[i] for i = 0, 2s, 4s, ...; additions (i, i+2s) share one instruction
template <int N, int START = 1> __device__ __forceinline__ float tree_sum_regs2(float (&t)[N]) {
#pragma unroll
  for (int stride = START; stride < N; stride <<= 1) {
#pragma unroll
    for (int i = 0; i + stride < N; i += 4 * stride) {
      if (i + 3 * stride < N) {
        const float2 r = __fadd2_rn(make_float2(t[i], t[i + 2 * stride]), make_float2(t[i + stride], t[i + 3 * stride]));
        t[i] = r.x; t[i + 2 * stride] = r.y;
      } else {
        t[i] = __fadd_rn(t[i], t[i + stride]);
      }
    }
  }
  return t[0];
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

// floor(log2(a)) for finite a >= 0 with the reference's "0 -> FP32_MIN_NORMAL" (quant.py:520-523), branch-free
__device__ __forceinline__ int floor_log2_or_min(float a) {
  const int u = __float_as_int(a);
  const int be = u >> 23;
  const int sub = (u ? 31 - __clz(u) : 23) - 149;     // subnormals; u == 0 gives -126
  return be ? be - 127 : sub;
}
// clamp_scale_exp without branches; S = 2^(scale_bits-1) - 1, U = value taken below -S
__device__ __forceinline__ int clamp_scale_sel(int e, int S, int U) {
  e = (e < -S) ? U : e;
  return (e > S) ? MSQ_EXP_NAN : e;
}

struct BlockOut {
  uint32_t mask;   // bit i set: x[i] is an outlier
  int e_in, e_out; // clamped shared exponents (MSQ_EXP_NAN on overflow)
  int cnt;         // outliers in the whole block (all lanes)
  int st;          // status bits raised by this block
  // HI: compacted hi fields (outlier code >> slot width) of elements 8g .. 8g+7, first outlier in the lowest bits
  uint32_t hacc[4];
};

// ---------------------------------------------------------------------------------------------
// core_outlier: the MX+outlier quantizer on one block.
// x[] in: block values (padding = 0).  x[] out: dequantized values.
// code[] out (if CODES): element codes, inlier or outlier format according to o.mask.
// EXT: detector bounds come from lo_e/hi_e (variant 1) instead of block statistics.
//
// Instruction budget matters here (this is the HBM-bound headline kernel), so the element path is
// written for the Blackwell pipes: pairs of elements go through packed fp32x2 instructions
// (FMUL2/FFMA2/FADD2, IEEE-rn per lane, so results are identical to scalar code), the
// inlier/outlier choice is an exact 0/1 blend  a*(1-m) + b*m  on the FMA pipe instead of selects,
// and the quantum exponent is clamped in the float domain (2^p = clamp(2^E, 2^lo, 2^hi)).
// Every blend is exact: one operand is multiplied by exactly 0, the other by exactly 1, and the
// format constants being blended (2^lo, 2^hi, 2^-mant, max_norm) are small dyadic numbers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
// a*(1-m) + b*m with m in {0,1}, nm = 1-m
__device__ __forceinline__ float2 blend2(float2 a, float2 b, float2 m, float2 nm) {
  return __ffma2_rn(b, m, __fmul2_rn(a, nm));
}
// c_i + m*(c_o - c_i) for exactly representable constants
__device__ __forceinline__ float2 cblend2(float ci, float dco, float2 m) { return __ffma2_rn(f2s(dco), m, f2s(ci)); }

// comparison straight to a 1.0f / 0.0f value (SASS FSET.BF): keeps the 16 per-block comparisons out of the
// 7 predicate registers (the compiler otherwise spills them through a bit-mask register)
__device__ __forceinline__ float fset_gt(float a, float b) {
  float d;
  asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
// comparison to an all-ones / all-zeros integer mask (SASS FSET without .BF)
__device__ __forceinline__ uint32_t fset_lt_u32(float a, float b) {
  uint32_t d;
  asm("set.lt.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float fset_lt(float a, float b) {
  float d;
  asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// (x < lo) || (x > hi) as 1.0f / 0.0f in TWO logic-pipe instructions (FSETP + FSET with the predicate as third operand)
__device__ __forceinline__ float fset_outside(float x, float lo, float hi) {
  float d;
  asm("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\nset.gt.or.f32.f32 %0, %1, %3, p;\n}\n" : "=f"(d) : "f"(x), "f"(lo), "f"(hi));
  return d;
}

struct ElemConsts {      // per-format constants of the float-domain rounding
  float L, H;            // 2^lo, 2^hi  (ints: 1, 1)
  float down;            // 2^-(bits-2)
  float mx;              // max_norm
};
__device__ __forceinline__ ElemConsts elem_consts(const Fmt& f) {
  ElemConsts e;
  e.L = f.L; e.H = f.H; e.down = f.down; e.mx = f.max_norm;    // precomputed on the host (make_fmt)
  return e;
}

// scalar version of the bounded element rounding (used for odd N and for round-to-even)
__device__ __forceinline__ float qelem_pair_scalar(float v, float L, float H, float down, float mx, const RoundSel rs) {
  const float P2 = fminf(fmaxf(__int_as_float(__float_as_int(v) & 0x7F800000), L), H);
  const float s2 = __fmul_rn(P2, down);
  const float s1 = __int_as_float(0x7F000000 - __float_as_int(s2));
  const float y = __fmul_rn(v, s1);
  float r;
  if (!rs.even) {
    r = truncf(__fadd_rn(y, copysignf(rs.addend, y)));
  } else {
    const float a = fabsf(y);
    r = floorf(__fadd_rn(a, 0.5f));
    const float t = __fadd_rn(a, -0.5f), hh = __fmul_rn(t, 0.5f);
    if (t >= 0.0f && hh == floorf(hh)) r = __fadd_rn(r, -1.0f);
    r = copysignf(r, y);
  }
  return fminf(fmaxf(__fmul_rn(r, s2), -mx), mx);
}

// ---- element rounding of a pair on the packed fp32x2 path -------------------------------------------------
// quantum = max(2^E(v) * Dp, LD)   (Dp = 2^-mant for float formats, 0 for integer formats; LD = 2^min_exp * 2^-mant)
__device__ __forceinline__ float2 quantum2(float2 v, float2 Dp, float2 LD) {
  const float2 PD = __fmul2_rn(f2(__int_as_float(__float_as_int(v.x) & 0x7F800000),
                                  __int_as_float(__float_as_int(v.y) & 0x7F800000)), Dp);
  return f2(fmaxf(PD.x, LD.x), fmaxf(PD.y, LD.y));
}
// torch-path rounding (elemwise_ops.py:59-72, nearest = floor(|y| + 0.5) with the fp32 add, floor): |q| clamped to M
// t = 2^23*quantum + k*quantum: the low mantissa bits of t are the integer k = floor(|v|/quantum + addend)
__device__ __forceinline__ float2 round_mag2_py_t(float2 v, float2 qn, float addend, float2 M, float2& t) {
  const float2 Cq = __fmul2_rn(qn, f2s(8388608.0f));
  const float ax = __fmaf_rn(qn.x, addend, fabsf(v.x)), ay = __fmaf_rn(qn.y, addend, fabsf(v.y));
  t = f2(__fadd_rz(ax, Cq.x), __fadd_rz(ay, Cq.y));
  return f2(fminf(__fsub_rn(t.x, Cq.x), M.x), fminf(__fsub_rn(t.y, Cq.y), M.y));
}
__device__ __forceinline__ float2 round_mag2_py(float2 v, float2 qn, float addend, float2 M) {
  float2 t;
  return round_mag2_py_t(v, qn, addend, M, t);
}
// integer-mantissa rounding of cpp/quantize.cuh:88-149 on the exact value: nearest = half away from zero on the
// truncated half-quantum count, even = ties to even, floor = truncation
__device__ __forceinline__ float2 round_mag2_native(float2 v, float2 qn, int rmode, float2 M) {
  const float2 Cq = __fmul2_rn(qn, f2s(8388608.0f));
  float qx, qy;
  if (rmode == RD_EVEN) {
    qx = __fsub_rn(__fadd_rn(fabsf(v.x), Cq.x), Cq.x);
    qy = __fsub_rn(__fadd_rn(fabsf(v.y), Cq.y), Cq.y);
  } else if (rmode == RD_FLOOR) {
    qx = __fsub_rn(__fadd_rz(fabsf(v.x), Cq.x), Cq.x);
    qy = __fsub_rn(__fadd_rz(fabsf(v.y), Cq.y), Cq.y);
  } else {
    const float2 hq = __fmul2_rn(qn, f2s(0.5f)), Ch = __fmul2_rn(qn, f2s(4194304.0f));
    const float tx = __fadd_rn(__fsub_rn(__fadd_rz(fabsf(v.x), Ch.x), Ch.x), hq.x);    // floor(2a/q) half quanta + one
    const float ty = __fadd_rn(__fsub_rn(__fadd_rz(fabsf(v.y), Ch.y), Ch.y), hq.y);
    qx = __fsub_rn(__fadd_rz(tx, Cq.x), Cq.x);
    qy = __fsub_rn(__fadd_rz(ty, Cq.y), Cq.y);
  }
  return f2(fminf(qx, M.x), fminf(qy, M.y));
}

template <int N, int LPB, bool CODES, bool EXT, bool CNT = true, bool HI = false>
__device__ __forceinline__ BlockOut core_outlier(float (&x)[N], uint32_t (&code)[N], const QCfg& c,
                                                 const float* lo_e, const float* hi_e) {
  constexpr int BS = N * LPB;
  BlockOut o;
  o.st = 0;
  o.mask = 0;
  o.cnt = 0;
  int bad = 0;
  float l = 0.0f, h = 0.0f;
  if (!EXT) {
    // mean(|A|), std(|A|) of utils/quant.py:477-478: fp32, two-pass, canonical pairwise-tree order
    float t[N];
    if (N >= 2) {     // first tree level with |.| as an operand modifier
#pragma unroll
      for (int i = 0; i + 1 < N; i += 2) t[i] = __fadd_rn(fabsf(x[i]), fabsf(x[i + 1]));
    } else {
      t[0] = fabsf(x[0]);
    }
    const float s = grp_sum_tree<LPB>(tree_sum_regs2<N, (N >= 2 ? 2 : 1)>(t));
    bad = !(s <= 3.402823466e38f);    // an Inf/NaN element (or an overflowing sum) makes it non-finite
    // BS is a power of two here, so multiplying by 1/BS is the division of utils/quant.py:477 exactly
    constexpr float inv_bs = 1.0f / (float)BS;
    static_assert((BS & (BS - 1)) == 0, "fast mappings use power-of-two blocks");
    const float mean = __fmul_rn(s, inv_bs);
    if (N % 4 == 0) {
      // ptxas (12.9) contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 even with -fmad=false: the first tree level
      // would become fma(d0, d0, d1*d1) -- one rounding less than the canonical order and a different detector bound in
      // ~1 block out of 10^7 (found by the full-size parity test).  The squares are therefore FFMA2 themselves,
      // d*d + z with z = 0.0f from the parameter block (a value the compiler cannot fold): rn(d*d + 0) == rn(d*d)
      // exactly, the instruction count stays that of FMUL2, and an FFMA result cannot be fused into the following add.
      const float2 z2 = f2s(c.fzero);
#pragma unroll
      for (int i = 0; i < N; i += 4) {      // squares paired the way the first tree level consumes them
        const float d0 = __fsub_rn(fabsf(x[i]), mean), d1 = __fsub_rn(fabsf(x[i + 1]), mean);
        const float d2 = __fsub_rn(fabsf(x[i + 2]), mean), d3 = __fsub_rn(fabsf(x[i + 3]), mean);
        const float2 a = __ffma2_rn(make_float2(d0, d2), make_float2(d0, d2), z2);
        const float2 b = __ffma2_rn(make_float2(d1, d3), make_float2(d1, d3), z2);
        t[i] = a.x; t[i + 2] = a.y; t[i + 1] = b.x; t[i + 3] = b.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float d = __fsub_rn(fabsf(x[i]), mean);
        t[i] = __fmul_rn(d, d);
      }
    }
    const float q = grp_sum_tree<LPB>(tree_sum_regs2<N>(t));
    const float sd = __fsqrt_rn(__fmul_rn(q, inv_bs));
    const float w = __fmul_rn(c.std_dev, sd);
    l = __fsub_rn(mean, w); h = __fadd_rn(mean, w);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) bad |= !(fabsf(x[i]) <= 3.402823466e38f);
  }
  if (EXT && LPB > 1) bad = grp_or<LPB>(bad);

  // outlier split exactly as quant.py:70-71: inlier_val = A*(1-mask), outlier_val = A*mask with an
  // exact 0/1 float mask.  From here on x[] holds the inlier part and t1[] the outlier part.
  float mf[N], t1[N];
  float m_in = 0.0f, fcnt = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float li = EXT ? lo_e[i] : l, hi_i = EXT ? hi_e[i] : h;
    mf[i] = fmaxf(fset_lt(x[i], li), fset_gt(x[i], hi_i));      // 0/1 float mask without predicates
    if (CNT && !(CODES && N % 2 == 0)) fcnt = __fadd_rn(fcnt, mf[i]);
  }
  if (CODES && N % 2 == 0) {
    // the block's outlier bits as an integer: sum of mask_i * 2^i, exact in fp32 for 16 bits at a time, on the FMA
    // pipe (two elements per FFMA2); the count is its popcount
    static_assert(N <= 32, "mask is one 32-bit word");
    float2 lo2 = f2s(0.0f), hi2 = f2s(0.0f);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const int sh = i & 15;
      const float2 wgt = f2((float)(1u << sh), (float)(2u << sh));
      if (i < 16) lo2 = __ffma2_rn(f2(mf[i], mf[i + 1]), wgt, lo2);
      else hi2 = __ffma2_rn(f2(mf[i], mf[i + 1]), wgt, hi2);
    }
    o.mask = __float2uint_rz(__fadd_rn(lo2.x, lo2.y));
    if (N > 16) o.mask |= __float2uint_rz(__fadd_rn(hi2.x, hi2.y)) << 16;
  }
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const float2 m = f2(mf[i], mf[i + 1]), xv = f2(x[i], x[i + 1]);
      const float2 nm = __ffma2_rn(m, f2s(-1.0f), f2s(1.0f));       // 1 - mask, exact
      const float2 to = __fmul2_rn(xv, m), ti = __fmul2_rn(xv, nm);
      t1[i] = to.x; t1[i + 1] = to.y; x[i] = ti.x; x[i + 1] = ti.y;
      m_in = fmaxf(m_in, fmaxf(fabsf(ti.x), fabsf(ti.y)));
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      t1[i] = __fmul_rn(x[i], mf[i]);
      x[i] = __fmul_rn(x[i], __fsub_rn(1.0f, mf[i]));
      m_in = fmaxf(m_in, fabsf(x[i]));
    }
  }
  if (CNT) o.cnt = grp_add<LPB>((CODES && N % 2 == 0) ? __popc(o.mask) : (int)fcnt);
  m_in = grp_max<LPB>(m_in);

  const int S_in = (1 << (c.sb_in - 1)) - 1, S_out = (1 << (c.sb_out - 1)) - 1;
  const int U_in = (c.variant == 0 && S_in > 20) ? -20 : -S_in, U_out = (c.variant == 0 && S_out > 20) ? -20 : -S_out;
  const int E_in = floor_log2_or_min(m_in);
  const bool flush = c.flush && (E_in <= -127);
  o.e_in = clamp_scale_sel(E_in - c.fin.emax, S_in, U_in);
  if (bad) o.st |= MSQ_ST_NONFINITE_INPUT;
  if (o.e_in == MSQ_EXP_NAN) o.st |= MSQ_ST_EXP_IN_OVERFLOW;
  const bool dead_in = bad || o.e_in == MSQ_EXP_NAN;
  const int e_in = dead_in ? 0 : o.e_in;
  const float s_in = pow2i(e_in), inv_s_in = pow2i(-e_in);

  // level-1 scaling of the outliers, t = outlier_val * 2^e_in (quant.py:94); their maximum gives e_out
  float m_out = 0.0f;
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const float2 ts = __fmul2_rn(f2(t1[i], t1[i + 1]), f2s(s_in));
      t1[i] = ts.x; t1[i + 1] = ts.y;
      m_out = fmaxf(m_out, fmaxf(fabsf(ts.x), fabsf(ts.y)));
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      t1[i] = __fmul_rn(t1[i], s_in);
      m_out = fmaxf(m_out, fabsf(t1[i]));
    }
  }
  m_out = grp_max<LPB>(m_out);
  const int E_out = floor_log2_or_min(m_out);
  o.e_out = dead_in ? MSQ_EXP_NAN : clamp_scale_sel(E_out - c.fout.emax, S_out, U_out);
  if (!dead_in && o.e_out == MSQ_EXP_NAN) o.st |= MSQ_ST_EXP_OUT_OVERFLOW;
  const bool dead = dead_in || o.e_out == MSQ_EXP_NAN;
  const int e_out = dead ? 0 : o.e_out;
  const float s_out = pow2i(e_out), inv_s_out = pow2i(-e_out);
  const RoundSel rs = make_round(c.rmode);
  const float inv_s_in_eff = flush ? 0.0f : inv_s_in;            // quant.py:80 (x*0 keeps the sign of x)
  const ElemConsts ci = elem_consts(c.fin), co = elem_consts(c.fout);
  const float dM = __fsub_rn(co.mx, ci.mx);
  // quantum = clamp(2^E(v), L, H) * D.  H = 2^127 never binds for float formats (v is finite), integer formats
  // have L = H = 1; both cases are  max(2^E(v) * D', L*D)  with D' = D (floats) or 0 (integers).
  const float Dp_i = c.fin.ebits ? ci.down : 0.0f, Dp_o = c.fout.ebits ? co.down : 0.0f;
  const float LD_i = __fmul_rn(ci.L, ci.down), LD_o = __fmul_rn(co.L, co.down);
  const float dDp = __fsub_rn(Dp_o, Dp_i), dLD = __fsub_rn(LD_o, LD_i);
  const float dSC = __fsub_rn(c.sc_o, c.sc_i), dKA = __fsub_rn(c.ka_o, c.ka_i), dTB = __fsub_rn(c.tb_o, c.tb_i);
  const float dMU = __fsub_rn(c.mu_o, c.mu_i);
  uint32_t hw = 1u;          // HI: weight of the next hi field inside its group of eight elements
  if (HI) { o.hacc[0] = o.hacc[1] = o.hacc[2] = o.hacc[3] = 0u; }
  const EncC enc_i = c.enc_i, enc_o = c.enc_o;
  auto encode_sel = [&](float qv, bool isout) {
    EncC e;
    e.thresh = isout ? enc_o.thresh : enc_i.thresh; e.addc = isout ? enc_o.addc : enc_i.addc;
    e.sh = isout ? enc_o.sh : enc_i.sh; e.c_norm = isout ? enc_o.c_norm : enc_i.c_norm;
    e.c_sub = isout ? enc_o.c_sub : enc_i.c_sub; e.signbit = isout ? enc_o.signbit : enc_i.signbit;
    return encode_fast(qv, e);
  };

  // inlier : v = in * 2^-e_in           ; r = q * 2^e_in                  quant.py:92,102
  // outlier: v = t * 2^-e_out           ; r = (q * 2^e_out) * 2^-e_in     quant.py:125,136
  // exactly one of the two scaled parts is non-zero, so v = vin + vout is exact and one rounding
  // with the element's own format replaces the reference's two (the other one rounds an exact 0).
  if ((N % 2 == 0) && !rs.even) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const float2 m = f2(mf[i], mf[i + 1]);
      const float2 vin = __fmul2_rn(f2(x[i], x[i + 1]), f2s(inv_s_in_eff));
      const float2 vout = __fmul2_rn(f2(t1[i], t1[i + 1]), f2s(inv_s_out));
      const float2 v = __fadd2_rn(vin, vout);
      const float2 Dp2 = cblend2(Dp_i, dDp, m), LD2 = cblend2(LD_i, dLD, m), M2 = cblend2(ci.mx, dM, m);
      // quantum 2^(p - mant) with p = clamp(exponent(v), lo, hi), done on powers of two in the float
      // domain.  floor(|v| / quantum + addend) * quantum is computed without leaving the FMA pipe:
      //   a = |v| + addend*quantum   (the reference's fp32 add, scaled by an exact power of two)
      //   floor to a multiple of the quantum = (a (+rz) C) - C  with C = 2^23 * quantum
      const float2 qn = quantum2(v, Dp2, LD2);
      float2 t;
      float2 q = round_mag2_py_t(v, qn, rs.addend, M2, t);
      q.x = copysignf(q.x, v.x);
      q.y = copysignf(q.y, v.y);
      if (CODES) {
        // float-domain encoder (enc2_consts): the element's own scale / shift / sign-bit constants are exact 0/1 blends
        // on the FMA pipe, |q| (already clamped to max_norm) is scaled into fp32's lowest binades where its bit
        // pattern is the magnitude code; the sign is set only for q < 0 (a rounded-to-zero negative keeps code 0)
        const float2 aq = f2(fabsf(q.x), fabsf(q.y));
        const float2 y = __fmul2_rn(aq, cblend2(c.sc_i, dSC, m));
        const float2 KA = cblend2(c.ka_i, dKA, m), TB = cblend2(c.tb_i, dTB, m);
        // sign bit: q + 0 turns a -0 into +0, so the sign of the sum is "q < 0"; bit 31 -> 0/1 by a multiply-high
        // and into place by a multiply-add (FMA pipe; the logic pipe is the busy one in the packing kernels)
        const float2 qz = __fadd2_rn(q, f2s(0.0f));
        const uint32_t s0 = __umulhi((uint32_t)__float_as_int(qz.x), 2u), s1 = __umulhi((uint32_t)__float_as_int(qz.y), 2u);
        const uint32_t p0 = s0 * (uint32_t)__float_as_int(TB.x) + (uint32_t)__float_as_int(y.x);
        const uint32_t p1 = s1 * (uint32_t)__float_as_int(TB.y) + (uint32_t)__float_as_int(y.y);
        code[i] = __funnelshift_r(p0, 0u, (uint32_t)__float_as_int(KA.x));          // shift count = low 5 bits
        code[i + 1] = __funnelshift_r(p1, 0u, (uint32_t)__float_as_int(KA.y));
        if (HI) {
          // compacted hi fields without predicates or variable shifts: field * weight summed, the weight multiplied by
          // 2^hi_w after every outlier and by 1 after an inlier (whose field is zero); all on the FMA pipe
          if ((i & 7) == 0) hw = 1u;
          const float2 MU = cblend2(c.mu_i, dMU, m);
          uint32_t& acc = o.hacc[(i >> 3) & 3];
          acc = __umulhi(code[i], c.hi_mulhi) * hw + acc;
          hw *= (uint32_t)__float_as_int(MU.x);
          acc = __umulhi(code[i + 1], c.hi_mulhi) * hw + acc;
          hw *= (uint32_t)__float_as_int(MU.y);
        }
      }
      x[i] = q.x; x[i + 1] = q.y;
    }
    // scale back: inlier q * 2^e_in, outlier (q * 2^e_out) * 2^-e_in, then the "+ 0" of quant.py:140.
    // When all exponents are far from the fp32 range limits both products are exact, so the outlier's two
    // multiplications equal one multiplication by 2^(e_out - e_in) and the "+ 0" folds into an FMA.
    // With the two scales within 22 binades of each other the per-element scale is an exact 0/1 blend as well.
    const bool safe = (abs(e_in) <= 100) & (abs(e_out) <= 100) & (abs(e_out - e_in) <= 100) & (abs(e_out - 2 * e_in) <= 22);
    if (safe) {
      const float s_oi = pow2i(e_out - e_in);
      const float dS = __fsub_rn(s_oi, s_in);
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        const float2 A = cblend2(s_in, dS, f2(mf[i], mf[i + 1]));
        const float2 r = __ffma2_rn(f2(x[i], x[i + 1]), A, f2s(0.0f));
        x[i] = r.x; x[i + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        const float2 q = f2(x[i], x[i + 1]);
        const float2 rin = __fmul2_rn(q, f2s(s_in));
        const float2 rout = __fmul2_rn(__fmul2_rn(q, f2s(s_out)), f2s(inv_s_in));
        const float2 r = __fadd2_rn(f2(mf[i] != 0.0f ? rout.x : rin.x, mf[i + 1] != 0.0f ? rout.y : rin.y), f2s(0.0f));
        x[i] = r.x; x[i + 1] = r.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const bool isout = mf[i] != 0.0f;
      const float v = __fadd_rn(__fmul_rn(x[i], inv_s_in_eff), __fmul_rn(t1[i], inv_s_out));
      const float qv = qelem_pair_scalar(v, isout ? co.L : ci.L, isout ? co.H : ci.H, isout ? co.down : ci.down,
                                         isout ? co.mx : ci.mx, rs);
      if (CODES) {
        code[i] = encode_sel(qv, isout);
        o.mask |= (isout ? 1u : 0u) << i;
        if (HI) {
          if ((i & 7) == 0) hw = 1u;
          o.hacc[(i >> 3) & 3] += __umulhi(code[i], c.hi_mulhi) * hw;
          hw *= isout ? (uint32_t)__float_as_int(c.mu_o) : 1u;
        }
      }
      const float r = isout ? __fmul_rn(__fmul_rn(qv, s_out), inv_s_in) : __fmul_rn(qv, s_in);
      x[i] = __fadd_rn(r, 0.0f);
    }
  }
  if (dead) {      // NaN/Inf input or a shared exponent above the scale range: the reference asserts
#pragma unroll
    for (int i = 0; i < N; ++i) { x[i] = qnan(); if (CODES) code[i] = 0u; }
    if (HI) { o.hacc[0] = o.hacc[1] = o.hacc[2] = o.hacc[3] = 0u; }
  }
  return o;
}

// ---------------------------------------------------------------------------------------------
// core_outlier_cvt: the packing kernels' block arithmetic for the drivers' own pairing -- fp4_e2m1 inliers, fp8_e4m3
// outliers, round-to-nearest, 16-element blocks held by one thread (llm/llama.py:130-137 of the reference) -- with the
// element rounding AND the code done by Blackwell's narrow-format conversion instructions
// (cvt.rn.satfinite.e2m1x2.f32 / .e4m3x2.f32, two elements each) instead of ~20 fp32 / integer instructions per element.
//
// Why the results are the reference's, bit for bit (elemwise_ops.py:47-81,131-167 with saturate_normals, allow_denorm):
//  * the scaled value v = x * 2^-e (inliers) / x * 2^(e_in - e_out) (outliers) is exact while the shared exponents are
//    far from the fp32 range limits (checked per block; an fp32-subnormal intermediate only occurs for |v| < 2^-46,
//    which both roundings send to zero);
//  * the reference rounds |y| = |v| / quantum to floor(fl(|y| + 0.5)) -- ties away from zero -- and clamps to max_norm;
//    the hardware rounds to nearest EVEN and saturates to the same max_norm.  Setting the last mantissa bit of x
//    (x | 1) before the (exact) scaling moves every exact tie just above the tie and changes nothing else: ties have
//    that bit clear (the format grids are >= 2^20 ulps coarse), no other value can cross or land on one;
//  * fl(|y| + 0.5) is itself rounded: for exactly one magnitude per format, the fp32 just below half the smallest
//    subnormal (|y| = 0.5 - 2^-25), the sum rounds up to 1.0 and the reference returns the smallest subnormal where
//    round-to-nearest gives zero (every other binade adds exactly, or lands where the floor does not change).  That
//    magnitude is compared for exactly (block-uniform constants) and patched in a branch that real data never takes;
//  * the reference's code for a value rounded to zero carries no sign (encode_elem): the sign bit of zero-magnitude
//    codes is cleared (SWAR on the packed words).
// Everything outside these conditions -- non-finite input, a shared exponent outside [-40, 40] (inliers) / [-80, 80]
// (outliers) or clamped by the scale width, an all-outlier block, flush-to-zero -- takes core_outlier unchanged.
// Outputs are the packed words the staging code wants: 4-bit slots of elements 0-7 / 8-15 and the compacted hi nibbles.
// ---------------------------------------------------------------------------------------------
struct BlockPk {
  uint32_t s0, s1;   // slot nibbles of elements 0..7 / 8..15 (element j of a group in nibble j)
  uint32_t h0, h1;   // compacted hi nibbles of the outliers of each group, first outlier in the lowest nibble
  uint32_t mask;
  int e_in, e_out, cnt, st;
};
// eight fp32 -> eight E2M1 codes, element j in nibble j
__device__ __forceinline__ uint32_t cvt8_e2m1(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  uint32_t d;
  asm("{\n"
      ".reg .b8 b0, b1, b2, b3;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b0, %2, %1;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b1, %4, %3;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b2, %6, %5;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b3, %8, %7;\n"
      "mov.b32 %0, {b0, b1, b2, b3};\n"
      "}\n"
      : "=r"(d)
      : "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7));
  return d;
}
// four fp32 -> four E4M3 codes, element j in byte j
__device__ __forceinline__ uint32_t cvt4_e4m3(float a0, float a1, float a2, float a3) {
  uint32_t d;
  asm("{\n"
      ".reg .b16 l, h;\n"
      "cvt.rn.satfinite.e4m3x2.f32 l, %2, %1;\n"
      "cvt.rn.satfinite.e4m3x2.f32 h, %4, %3;\n"
      "mov.b32 %0, {l, h};\n"
      "}\n"
      : "=r"(d)
      : "f"(a0), "f"(a1), "f"(a2), "f"(a3));
  return d;
}
// 16 codes (<= 8 bits, low nibble = slot) -> the two slot words
__device__ __forceinline__ void slots_from_codes16(const uint32_t* c, uint32_t& W0, uint32_t& W1) {
  // four codes per word by multiply-adds (sums of disjoint bytes), low nibbles kept, adjacent nibbles merged:
  // byte 0 / byte 2 of t = elements (4j, 4j+1) / (4j+2, 4j+3)
  uint32_t t[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t w = ((c[4 * j + 3] * 256u + c[4 * j + 2]) * 256u + c[4 * j + 1]) * 256u + c[4 * j];
    const uint32_t l = w & 0x0F0F0F0Fu;
    t[j] = l | (l >> 4);
  }
  W0 = __byte_perm(t[0], t[1], 0x6420);
  W1 = __byte_perm(t[2], t[3], 0x6420);
}
// bits 0..7 -> nibbles 0..7 (0xF where the bit is set)
__device__ __forceinline__ uint32_t nibble_mask8(uint32_t m8) {
  uint32_t x = (m8 * 0x1001u) & 0x000F000Fu;       // bits 0-3 | bits 4-7 at bit 16
  x = (x * 0x41u) & 0x03030303u;
  x = (x * 0x9u) & 0x11111111u;
  return x * 15u;
}

__device__ __forceinline__ BlockPk core_outlier_cvt(const float (&x)[16], const QCfg& c) {
  constexpr int N = 16;
  BlockPk o;
  o.st = 0;
  // ---- detector bounds: the statements of core_outlier (same order of fp32 operations)
  float t[N];
#pragma unroll
  for (int i = 0; i + 1 < N; i += 2) t[i] = __fadd_rn(fabsf(x[i]), fabsf(x[i + 1]));
  const float s = tree_sum_regs2<N, 2>(t);
  const bool bad = !(s <= 3.402823466e38f);
  constexpr float inv_bs = 1.0f / (float)N;
  const float mean = __fmul_rn(s, inv_bs);
  {
    const float2 z2 = f2s(c.fzero);
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float d0 = __fsub_rn(fabsf(x[i]), mean), d1 = __fsub_rn(fabsf(x[i + 1]), mean);
      const float d2 = __fsub_rn(fabsf(x[i + 2]), mean), d3 = __fsub_rn(fabsf(x[i + 3]), mean);
      const float2 a = __ffma2_rn(make_float2(d0, d2), make_float2(d0, d2), z2);
      const float2 b = __ffma2_rn(make_float2(d1, d3), make_float2(d1, d3), z2);
      t[i] = a.x; t[i + 2] = a.y; t[i + 1] = b.x; t[i + 3] = b.y;
    }
  }
  const float q = tree_sum_regs2<N>(t);
  const float sd = __fsqrt_rn(__fmul_rn(q, inv_bs));
  const float w = __fmul_rn(c.std_dev, sd);
  const float l = __fsub_rn(mean, w), h = __fadd_rn(mean, w);
  // ---- mask, the two maxima (|x * (1 - m)|, |x * m| are the magnitudes of quant.py:70-71's products)
  float mf[N];
  float m_in = 0.0f, m_out = 0.0f;
  float2 lo2 = f2s(0.0f);
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    mf[i] = fset_outside(x[i], l, h);
    mf[i + 1] = fset_outside(x[i + 1], l, h);
    const float2 m = f2(mf[i], mf[i + 1]), xv = f2(x[i], x[i + 1]);
    lo2 = __ffma2_rn(m, f2((float)(1u << i), (float)(2u << i)), lo2);
    const float2 nm = __ffma2_rn(m, f2s(-1.0f), f2s(1.0f));
    const float2 to = __fmul2_rn(xv, m), ti = __fmul2_rn(xv, nm);
    m_in = fmaxf(m_in, fmaxf(fabsf(ti.x), fabsf(ti.y)));
    m_out = fmaxf(m_out, fmaxf(fabsf(to.x), fabsf(to.y)));
  }
  o.mask = __float2uint_rz(__fadd_rn(lo2.x, lo2.y));
  o.cnt = __popc(o.mask);
  // ---- shared exponents exactly as core_outlier; decide whether this block is on the conversion path
  // ---- shared exponents.  On the conversion path both maxima are fp32 normals well inside the range, so floor(log2) is the
  // exponent field and the clamp of the scale width cannot bind (range limits c.cvt_lim_* = min(40 / 80, S), finish_cfg);
  // a block of zeros (padding) and a block without outliers get the clamped "exponent of zero" the reference computes
  const int er_in = (__float_as_int(m_in) >> 23) - 127 - 2;                     // fp4_e2m1: emax = 2
  const bool zero_in = (m_in == 0.0f) && (o.mask == 0u);       // nothing but zeros: every code is 0 whatever the scale
  bool fast = !bad && ((er_in >= -c.cvt_lim_in && er_in <= c.cvt_lim_in) || zero_in);
  const int e_in = zero_in ? 0 : er_in;
  o.e_in = zero_in ? c.e_zero_in : er_in;
  // max |t * 2^e_in| = (max |t|) * 2^e_in: scaling by a power of two is monotone (and exact in this range)
  const int er_out = (__float_as_int(__fmul_rn(m_out, pow2i(fast ? e_in : 0))) >> 23) - 127 - 8;      // fp8_e4m3: emax = 8
  o.e_out = o.mask == 0u ? c.e_zero_out : er_out;
  fast = fast && (o.mask == 0u || (er_out >= -c.cvt_lim_out && er_out <= c.cvt_lim_out));
  if (!fast) {
    float xc[N];
    uint32_t code[N];
#pragma unroll
    for (int i = 0; i < N; ++i) xc[i] = x[i];
    const BlockOut b = core_outlier<N, 1, true, false, true, true>(xc, code, c, nullptr, nullptr);
    slots_from_codes16(code, o.s0, o.s1);
    o.h0 = b.hacc[0]; o.h1 = b.hacc[1];
    o.mask = b.mask; o.e_in = b.e_in; o.e_out = b.e_out; o.cnt = b.cnt; o.st = b.st;
    return o;
  }
  const int e_out = o.mask ? er_out : e_in;
  const float si = pow2i(-e_in), so = pow2i(e_in - e_out);
  // the one magnitude per format where fl(|y| + 0.5) rounds up to 1 (see above): half the smallest subnormal minus an ulp
  const float c_in = __int_as_float(__float_as_int(pow2i(e_in + c.fin.min_exp - (c.fin.bits - 2) - 1)) - 1);
  const float c_out = __int_as_float(__float_as_int(pow2i(e_out - e_in + c.fout.min_exp - (c.fout.bits - 2) - 1)) - 1);
  float vi[N], vo[N];
  // any |x| equal to one of the two magnitudes?  (|x| - c_in) * (|x| - c_out) is zero exactly then (both differences are
  // exact zeros or non-zero, the factors are finite); an underflowing product of two tiny differences only sends the block
  // through the exact branch below for nothing.  Differences and products on the FMA pipe, one 3-input minimum per pair on
  // the logic pipe (the pipe that bounds this kernel) instead of four comparisons
  float pmin = 1.0f;
  const float2 nci = f2s(-c_in), nco = f2s(-c_out);
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    const float2 ax = f2(fabsf(x[i]), fabsf(x[i + 1]));
    const float2 pr = __fmul2_rn(__fadd2_rn(ax, nci), __fadd2_rn(ax, nco));
    pmin = fminf(pmin, fminf(fabsf(pr.x), fabsf(pr.y)));
    const float2 xb = f2(__int_as_float(__float_as_int(x[i]) | 1), __int_as_float(__float_as_int(x[i + 1]) | 1));
    const float2 a = __fmul2_rn(xb, f2s(si)), b = __fmul2_rn(xb, f2s(so));
    vi[i] = a.x; vi[i + 1] = a.y; vo[i] = b.x; vo[i + 1] = b.y;
  }
  uint32_t in0 = cvt8_e2m1(vi[0], vi[1], vi[2], vi[3], vi[4], vi[5], vi[6], vi[7]);
  uint32_t in1 = cvt8_e2m1(vi[8], vi[9], vi[10], vi[11], vi[12], vi[13], vi[14], vi[15]);
  uint32_t ow[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) ow[j] = cvt4_e4m3(vo[4 * j], vo[4 * j + 1], vo[4 * j + 2], vo[4 * j + 3]);
  if (pmin == 0.0f) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const uint32_t sgn = (uint32_t)__float_as_int(x[i]) >> 31;
      if (fabsf(x[i]) == c_in) {
        uint32_t& wd = (i < 8) ? in0 : in1;
        wd = (wd & ~(0xFu << (4 * (i & 7)))) | ((1u | (sgn << 3)) << (4 * (i & 7)));
      }
      if (fabsf(x[i]) == c_out) {
        uint32_t& wd = ow[i >> 2];
        wd = (wd & ~(0xFFu << (8 * (i & 3)))) | ((1u | (sgn << 7)) << (8 * (i & 3)));
      }
    }
  }
  // ---- a zero magnitude carries no sign
  {
    const uint32_t a0 = (in0 & 0x77777777u) + 0x77777777u, a1 = (in1 & 0x77777777u) + 0x77777777u;
    in0 &= (a0 & 0x88888888u) | 0x77777777u;
    in1 &= (a1 & 0x88888888u) | 0x77777777u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t a = (ow[j] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
      ow[j] &= (a & 0x80808080u) | 0x7F7F7F7Fu;
    }
  }
  // ---- low / high nibbles of the outlier bytes as nibble words, slot select by the mask
  uint32_t lp[4], hp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t y4 = ow[j] >> 4, y8 = ow[j] >> 8;
    lp[j] = (ow[j] & 0x0F0F0F0Fu) | (y4 & 0xF0F0F0F0u);      // byte 0 = low nibbles of elements (0, 1), byte 2 = (2, 3)
    hp[j] = (y4 & 0x0F0F0F0Fu) | (y8 & 0xF0F0F0F0u);         // same for the high nibbles
  }
  const uint32_t nm0 = nibble_mask8(o.mask & 0xFFu), nm1 = nibble_mask8(o.mask >> 8);
  const uint32_t l0 = __byte_perm(lp[0], lp[1], 0x6420), l1 = __byte_perm(lp[2], lp[3], 0x6420);
  const uint32_t g0 = __byte_perm(hp[0], hp[1], 0x6420) & nm0, g1 = __byte_perm(hp[2], hp[3], 0x6420) & nm1;
  o.s0 = (in0 & ~nm0) | (l0 & nm0);
  o.s1 = (in1 & ~nm1) | (l1 & nm1);
  // ---- hi nibbles compacted: from the top element down, (acc : next nibble) is funnel-shifted left by 4 for an outlier
  // and by 0 for an inlier (the shift count is the 4-bit of the element's mask nibble; SHF takes it modulo 32)
  {
    const uint32_t a0 = nm0 & 0x44444444u, a1 = nm1 & 0x44444444u;
    uint32_t acc0 = 0u, acc1 = 0u;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      acc0 = __funnelshift_l(g0 << (4 * (7 - i)), acc0, a0 >> (4 * i));
      acc1 = __funnelshift_l(g1 << (4 * (7 - i)), acc1, a1 >> (4 * i));
    }
    o.h0 = acc0; o.h1 = acc1;
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
  // float-domain constants of the element rounding (see core_outlier)
  const float Dp = f.ebits ? f.down : 0.0f, LD = __fmul_rn(f.L, f.down);
  if (c.mode == 0) {
    const int E = floor_log2_or_min(m);
    const bool flush = c.flush && (E <= -127);
    const int S = (1 << (c.sb_in - 1)) - 1;
    int e = bad ? MSQ_EXP_NAN : clamp_scale_sel(E - f.emax, S, -S);
    o.e = e;
    if (e == MSQ_EXP_NAN) {
      o.st |= bad ? MSQ_ST_NONFINITE_INPUT : MSQ_ST_EXP_IN_OVERFLOW;
#pragma unroll
      for (int i = 0; i < N; ++i) { x[i] = qnan(); code[i] = 0; }
      return o;
    }
    const float s = pow2i(e);
    const float den = __fadd_rn(s, 1e-6f);                       // mx_ops.py:444
    const float keep = flush ? 0.0f : 1.0f;                      // x * 0 keeps the sign like the reference's flush
    if ((N % 2 == 0) && c.rmode != RD_EVEN) {
      const float addend = (c.rmode == RD_FLOOR) ? 0.0f : 0.5f;
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        const float2 xs = __fmul2_rn(f2(x[i], x[i + 1]), f2s(keep));
        const float2 v = f2(__fdiv_rn(xs.x, den), __fdiv_rn(xs.y, den));
        const float2 qn = quantum2(v, f2s(Dp), f2s(LD));
        float2 q = round_mag2_py(v, qn, addend, f2s(f.max_norm));
        q.x = copysignf(q.x, v.x); q.y = copysignf(q.y, v.y);
        if (want_codes) { code[i] = encode_fast(q.x, c.enc_i); code[i + 1] = encode_fast(q.y, c.enc_i); }
        const float2 r = __fmul2_rn(q, f2s(s));
        x[i] = r.x; x[i + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float v = __fdiv_rn(flush ? __fmul_rn(x[i], 0.0f) : x[i], den);
        float qv = qelem_rt(v, f, c.rmode);
        if (want_codes) code[i] = encode_fast(qv, c.enc_i);
        x[i] = __fmul_rn(qv, s);
      }
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
    if ((N % 2 == 0) && sh != 255 && sh >= 0) {
      // x / 2^k == x * 2^-k (both the correctly rounded value of the same real number)
      const float inv = flush ? 0.0f : pow2i(127 - sh);
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        float2 v = __fmul2_rn(f2(x[i], x[i + 1]), f2s(inv));
        if (flush) v = f2s(0.0f);                                 // mx.cuh:45 (plain zero, not x*0)
        const float2 qn = quantum2(v, f2s(Dp), f2s(LD));
        float2 q = round_mag2_native(v, qn, c.rmode, f2s(f.max_norm));
        q.x = copysignf(q.x, v.x); q.y = copysignf(q.y, v.y);
        q = __fadd2_rn(q, f2s(0.0f));                             // quantize.cuh returns +0 for a zero mantissa
        if (want_codes) { code[i] = encode_fast(q.x, c.enc_i); code[i + 1] = encode_fast(q.y, c.enc_i); }
        const float2 r = __fmul2_rn(q, f2s(s));
        x[i] = r.x; x[i + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float v = flush ? 0.0f : __fdiv_rn(x[i], s);
        float qv = qelem_native(v, f.bits, f.ebits, f.max_norm, c.rmode, true, true);
        if (want_codes) code[i] = (qv == qv) ? encode_fast(qv, c.enc_i) : 0;
        x[i] = __fmul_rn(qv, s);
      }
    }
  }
  return o;
}

}  // namespace msq
