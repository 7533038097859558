// msq_pack.cu -- dequantize ("unpack") the packed MicroScopiQ planes produced by the PACK variants of
// the quantizer kernels (csrc/msq_quant_strided.cu, msq_quant_inner.cu).  Layout: DESIGN.md "packed layout".
//   codes   : one inlier-width slot per element, linear element order
//   bitmap  : 1 bit per element (set = outlier), 32-bit little-endian words
//   e_in/e_out : int8 shared exponents per block, block-linear order (MSQ_EXP_NAN = NaN scale)
//   hi      : per segment (MSQ_SEG_BLOCKS blocks) the high bits of the outlier codes, compacted in
//             block-major order, hi_w bits each, starting at seg * hi_seg_stride
// unpack(pack(W)) reproduces the reference's dequantized tensor bit for bit (tests/test_gpu_pack.py).
#include "msq_quant_kernels.cuh"
#include <stdlib.h>
#include <cuda_fp16.h>

namespace msq {

struct UParams {
  long long pre, post, nblocks;
  int L, bs, nblk;
  Fmt fin, fout;
  const uint8_t* codes; const int8_t* e_in; const int8_t* e_out; const uint32_t* bitmap; const uint8_t* hi;
  long long hi_seg_stride; int hi_w;
  float* out_f32; __nv_bfloat16* out_bf16;
  int plain;          // 1: plain MX (no bitmap / outliers)
  unsigned int post_m, nblk_m; int post_k, nblk_k;      // magic dividers (k_unpack_strided_cvt)
};

__device__ __forceinline__ float dq_elem(uint32_t code, bool isout, const UParams& P, float s_in, float inv_s_in, float s_out,
                                         bool nan_blk) {
  float q = isout ? decode_elem(code, P.fout.ebits, P.fout.bits) : decode_elem(code, P.fin.ebits, P.fin.bits);
  float r = isout ? __fmul_rn(__fmul_rn(q, s_out), inv_s_in) : __fmul_rn(q, s_in);
  r = __fadd_rn(r, 0.0f);
  return nan_blk ? qnan() : r;
}

__device__ __forceinline__ uint32_t hi_field(const uint32_t* s_hi, int k, int hw) {
  const int bit = k * hw;
  return (s_hi[bit >> 5] >> (bit & 31)) & ((1u << hw) - 1u);
}

// strided layout: thread = block (column), 256 consecutive block-linear indices per sub-iteration
__global__ void __launch_bounds__(256) k_unpack_strided(const UParams P) {
  extern __shared__ uint32_t s_hi[];
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const long long seg = blockIdx.x;
  const bool has_hi = !P.plain && P.hi_w > 0;
  const int wi = P.fin.width;
  const int lane = threadIdx.x & 31;
  const int post = (int)P.post;
  if (has_hi) {
    const int words = (int)(P.hi_seg_stride >> 2);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.hi + seg * P.hi_seg_stride);
    for (int i = threadIdx.x; i < words; i += 256) s_hi[i] = __ldg(src + i);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
  }
  for (int sub = 0; sub < MSQ_SEG_BLOCKS / 256; ++sub) {
    const long long t0 = seg * MSQ_SEG_BLOCKS + (long long)sub * 256 + threadIdx.x;
    const bool active = t0 < P.nblocks;
    int base = 0, b = 0;
    if (active) {
      const int pb = (int)(t0 / post);
      const int j = (int)(t0 - (long long)pb * post);
      const int p = pb / P.nblk; b = pb - p * P.nblk;
      base = (p * P.L + b * P.bs) * post + j;
    }
    const int nv = active ? min(P.bs, P.L - b * P.bs) : 0;
    // outlier mask of this block: one bitmap word per row, shared by the warp (post % 32 == 0)
    uint32_t mask = 0;
    if (!P.plain) {
      for (int i = 0; i < nv; ++i) {
        const uint32_t w = __ldg(P.bitmap + ((base + i * post) >> 5));
        mask |= ((w >> lane) & 1u) << i;
      }
    }
    int k = 0;
    if (has_hi) {
      int tot;
      const int off = cta_excl_scan(__popc(mask), s_warp, &tot);
      k = s_base + off;
      __syncthreads();
      if (threadIdx.x == 0) s_base += tot;
      __syncthreads();
    }
    if (!active) continue;
    const int ei = P.e_in[t0], eo = P.plain ? 0 : P.e_out[t0];
    const bool nan_blk = ei == MSQ_EXP_NAN || eo == MSQ_EXP_NAN;
    const float s_in = pow2i(nan_blk ? 0 : ei), inv_s_in = pow2i(nan_blk ? 0 : -ei), s_out = pow2i(nan_blk ? 0 : eo);
    for (int i = 0; i < nv; ++i) {
      const int lin = base + i * post;
      const long long bit = (long long)lin * wi;
      uint32_t code = (P.codes[bit >> 3] >> (bit & 7)) & ((1u << wi) - 1u);
      const bool isout = (mask >> i) & 1u;
      if (isout && has_hi) { code |= hi_field(s_hi, k, P.hi_w) << wi; ++k; }
      const float r = dq_elem(code, isout, P, s_in, inv_s_in, s_out, nan_blk);
      if (P.out_f32) P.out_f32[lin] = r;
      if (P.out_bf16) P.out_bf16[lin] = __float2bfloat16_rn(r);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Fast strided unpack: lane = column, thread = one MX block, BS compile-time.
//  * decoding is one shared-memory table lookup (2^wi inlier entries followed by 2^(wi+hi_w) outlier entries,
//    built per CTA from decode_elem);
//  * code / bitmap rows are fetched as whole 32-bit words shared by the lanes that need them;
//  * the position of a warp's hi fields inside the segment stream comes from a prologue that popcounts the
//    segment's bitmap words once (32 groups of 32 blocks), so the main loop has no CTA-wide synchronisation
//    and its four sub-iterations are independent; inside a warp a shuffle scan gives the block offset and the
//    popcount of the lower mask bits the element offset (no serial dependency, no branches);
//  * only the used prefix of the hi stream is staged.
// ---------------------------------------------------------------------------------------------------------
// base + stride * i as one IMAD.WIDE
__device__ __forceinline__ const char* row_at(const void* base, int stride_bytes, int i) {
  long long a;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(a) : "r"(stride_bytes), "r"(i), "l"(reinterpret_cast<long long>(base)));
  return reinterpret_cast<const char*>(a);
}

// HWT: hi field width known at compile time (0 = no hi stream, 4), or -1 = run-time P.hi_w
template <int BS, int HWT>
__global__ void __launch_bounds__(256) k_unpack_strided_fast(const UParams P) {
  constexpr bool HAS_HI = HWT != 0;
  const int hw = HWT >= 0 ? HWT : P.hi_w;
  extern __shared__ uint32_t s_dyn[];
  __shared__ int s_off[MSQ_SEG_BLOCKS / 32 + 1];
  const int wi = P.fin.width;
  const int wo = wi + hw;                             // full outlier code width
  float* lut = reinterpret_cast<float*>(s_dyn);       // [2^wi | 2^wo]
  const int lut_n = (1 << wi) + (1 << wo);
  uint32_t* s_hi = s_dyn + lut_n;
  const int seg = (int)blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int post = (int)P.post;
  constexpr int NGRP = MSQ_SEG_BLOCKS / 32;           // 32 groups of 32 blocks (= one warp's columns)
  for (int i = threadIdx.x; i < lut_n; i += 256)
    lut[i] = (i < (1 << wi)) ? decode_elem((uint32_t)i, P.fin.ebits, P.fin.bits)
                             : decode_elem((uint32_t)(i - (1 << wi)), P.fout.ebits, P.fout.bits);
  if (HAS_HI) {
    // outliers per group: BS bitmap words each, BS consecutive threads reduce with shuffles
    for (int w = threadIdx.x; w < NGRP * BS; w += 256) {
      const int g = w / BS, i = w % BS;
      const int t = seg * MSQ_SEG_BLOCKS + g * 32;
      int cnt = 0;
      if (t < (int)P.nblocks) {
        const int pb = t / post, j = t - pb * post;
        const int p = pb / P.nblk, b = pb - p * P.nblk;
        if (i < min(BS, P.L - b * BS)) cnt = __popc(__ldg(P.bitmap + (((p * P.L + b * BS + i) * post + j) >> 5)));
      }
#pragma unroll
      for (int o = BS / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (i == 0) s_off[g] = cnt;
    }
    __syncthreads();
    if (warp == 0) {
      int tot;
      const int ex = warp_excl_scan(s_off[lane], lane, tot);
      s_off[lane] = ex;
      if (lane == 0) s_off[NGRP] = tot;
    }
    __syncthreads();
    const int words = (s_off[NGRP] * hw + 31) >> 5;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.hi + (long long)seg * P.hi_seg_stride);
    for (int i = threadIdx.x; i < words; i += 256) s_hi[i] = __ldg(src + i);
  }
  __syncthreads();
  const uint32_t cm = (1u << wi) - 1u, hm = (1u << hw) - 1u;
  const int cshift = (lane * wi) & 31;                // this lane's slot inside the row's code word
  const uint32_t* codes32 = reinterpret_cast<const uint32_t*>(P.codes) + ((lane * wi) >> 5);
  const int cstride = (post >> 5) * wi * 4;           // bytes between the code words of consecutive rows
  const int bstride = (post >> 5) * 4;                // same for bitmap words
  const int sb = post * 4;
#pragma unroll 2
  for (int sub = 0; sub < MSQ_SEG_BLOCKS / 256; ++sub) {
    const int t0 = seg * MSQ_SEG_BLOCKS + sub * 256 + (int)threadIdx.x;
    const bool active = t0 < (int)P.nblocks;
    int base = 0, b = 0;
    if (active) {
      const int pb = t0 / post;
      const int j = t0 - pb * post;
      const int p = pb / P.nblk; b = pb - p * P.nblk;
      base = (p * P.L + b * BS) * post + j;
    }
    const int nv = active ? min(BS, P.L - b * BS) : 0;
    const int ei = active ? P.e_in[t0] : 0, eo = (active && !P.plain) ? P.e_out[t0] : 0;
    const int row32 = (base - lane) >> 5;             // index of this warp's 32-column group in row 0 of the block
    const char* cp = reinterpret_cast<const char*>(codes32 + (long long)row32 * wi);
    const char* bp = reinterpret_cast<const char*>(P.bitmap + row32);
    uint32_t mask = 0, cw[BS];
    if (nv == BS) {
#pragma unroll
      for (int i = 0; i < BS; ++i) {
        cw[i] = __ldg(reinterpret_cast<const uint32_t*>(row_at(cp, cstride, i)));
        if (!P.plain) mask |= ((__ldg(reinterpret_cast<const uint32_t*>(row_at(bp, bstride, i))) >> lane) & 1u) << i;
      }
    } else {
#pragma unroll
      for (int i = 0; i < BS; ++i) {
        cw[i] = 0;
        if (i < nv) {
          cw[i] = __ldg(reinterpret_cast<const uint32_t*>(row_at(cp, cstride, i)));
          if (!P.plain) mask |= ((__ldg(reinterpret_cast<const uint32_t*>(row_at(bp, bstride, i))) >> lane) & 1u) << i;
        }
      }
    }
    int k = 0;
    if (HAS_HI) {
      int wt;
      k = s_off[sub * 8 + warp] + warp_excl_scan(__popc(mask), lane, wt);
    }
    const bool nan_blk = ei == MSQ_EXP_NAN || eo == MSQ_EXP_NAN;
    const float s_in = pow2i(nan_blk ? 0 : ei), inv_s_in = pow2i(nan_blk ? 0 : -ei), s_out = pow2i(nan_blk ? 0 : eo);
    // decode in place: cw[i] becomes the fp32 bits of element i
#pragma unroll
    for (int i = 0; i < BS; ++i) {
      uint32_t c = (cw[i] >> cshift) & cm;
      const uint32_t isout = (mask >> i) & 1u;
      if (HAS_HI) {
        const int bit = (k + __popc(mask & ((1u << i) - 1u))) * hw;
        const uint32_t f = (s_hi[bit >> 5] >> (bit & 31)) & hm;
        c |= (f & (0u - isout)) << wi;
      }
      const float q = lut[c + (isout << wi)];
      float r = isout ? __fmul_rn(__fmul_rn(q, s_out), inv_s_in) : __fmul_rn(q, s_in);
      r = nan_blk ? qnan() : __fadd_rn(r, 0.0f);
      cw[i] = __float_as_uint(r);
    }
    if (P.out_f32) {
      float* o = P.out_f32 + base;
      if (nv == BS) {
#pragma unroll
        for (int i = 0; i < BS; ++i) *reinterpret_cast<uint32_t*>(const_cast<char*>(row_at(o, sb, i))) = cw[i];
      } else {
#pragma unroll
        for (int i = 0; i < BS; ++i)
          if (i < nv) *reinterpret_cast<uint32_t*>(const_cast<char*>(row_at(o, sb, i))) = cw[i];
      }
    }
    if (P.out_bf16) {
      __nv_bfloat16* o = P.out_bf16 + base;
#pragma unroll
      for (int i = 0; i < BS; ++i)
        if (i < nv)
          *reinterpret_cast<__nv_bfloat16*>(const_cast<char*>(row_at(o, sb >> 1, i))) = __float2bfloat16_rn(__uint_as_float(cw[i]));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// k_unpack_strided_cvt: the strided unpack of the drivers' own pairing (fp4_e2m1 slots, fp8_e4m3 outliers with a
// 4-bit hi field, 16-row blocks) on packed words.  k_unpack_strided_fast spends 54 instructions per element
// (ncu: 76 M warp instructions for 45 M elements, ALU pipe 66 %, issue 71 %: instruction-bound at 0.33 of the copy
// roofline) on per-element table lookups, a popcount + shared-memory load + two shifts per hi field and per-element
// selects.  Here a thread (= column = one MX block) gathers its 16 slot nibbles and 16 mask bits by funnel shifts
// (one shift each, no masking), takes the block's hi nibbles as ONE unaligned 64-bit window of the segment stream
// and deals them to the outlier positions with funnel shifts, and decodes with the hardware conversions
// (cvt.rn.f16x2.e2m1x2 / e4m3x2, exact: every code value is an fp16 value): the inlier and the outlier decode of
// an element are both computed (one of them from a zeroed code) and added, so there is no select.
// Same prologue (segment offsets, hi staging) and the same results, bit for bit, as k_unpack_strided_fast.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 dec2_e2m1(uint32_t byte) {          // two E2M1 codes (low nibble first) -> fp32
  uint32_t h;
  asm("{\n.reg .b8 b;\ncvt.u8.u32 b, %1;\ncvt.rn.f16x2.e2m1x2 %0, b;\n}\n" : "=r"(h) : "r"(byte));
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
__device__ __forceinline__ float2 dec2_e4m3(uint32_t two_bytes) {     // two E4M3 codes (low byte first) -> fp32
  uint32_t h;
  asm("{\n.reg .b16 b;\ncvt.u16.u32 b, %1;\ncvt.rn.f16x2.e4m3x2 %0, b;\n}\n" : "=r"(h) : "r"(two_bytes));
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
// 4 nibbles (bits 0..15) -> 4 bytes (nibble j in the low half of byte j)
__device__ __forceinline__ uint32_t spread4(uint32_t n) {
  uint32_t x = (n | (n << 8)) & 0x00FF00FFu;
  return (x | (x << 4)) & 0x0F0F0F0Fu;
}
// bits 0..7 -> nibbles 0..7 (0xF where the bit is set)
__device__ __forceinline__ uint32_t nibble_mask8u(uint32_t m8) {
  uint32_t x = (m8 * 0x1001u) & 0x000F000Fu;
  x = (x * 0x41u) & 0x03030303u;
  x = (x * 0x9u) & 0x11111111u;
  return x * 15u;
}

template <bool BF16OUT>      // output type; L % 16 == 0 (no ragged row block: the table kernel takes those)
__global__ void __launch_bounds__(256) k_unpack_strided_cvt(const UParams P) {
  constexpr int BS = 16, hw = 4;
  extern __shared__ uint32_t s_dyn[];
  __shared__ int s_off[MSQ_SEG_BLOCKS / 32 + 1];
  uint32_t* s_hi = s_dyn;
  const int seg = (int)blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int post = (int)P.post;
  constexpr int NGRP = MSQ_SEG_BLOCKS / 32;
  // ---- prologue of k_unpack_strided_fast: outliers per group of 32 blocks, exclusive scan, stage the used hi prefix
  for (int w = threadIdx.x; w < NGRP * BS; w += 256) {
    const int g = w / BS, i = w % BS;
    const int t = seg * MSQ_SEG_BLOCKS + g * 32;
    int cnt = 0;
    if (t < (int)P.nblocks) {
      const int pb = t / post, j = t - pb * post;
      const int p = pb / P.nblk, b = pb - p * P.nblk;
      if (i < min(BS, P.L - b * BS)) cnt = __popc(__ldg(P.bitmap + (((p * P.L + b * BS + i) * post + j) >> 5)));
    }
#pragma unroll
    for (int o = BS / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (i == 0) s_off[g] = cnt;
  }
  __syncthreads();
  if (warp == 0) {
    int tot;
    const int ex = warp_excl_scan(s_off[lane], lane, tot);
    s_off[lane] = ex;
    if (lane == 0) s_off[NGRP] = tot;
  }
  __syncthreads();
  {
    const int words = (s_off[NGRP] * hw + 31) >> 5;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.hi + (long long)seg * P.hi_seg_stride);
    for (int i = threadIdx.x; i < words + 2; i += 256) s_hi[i] = i < words ? __ldg(src + i) : 0u;      // + 2: the window below
  }
  __syncthreads();
  const int cshift = (lane * 4) & 31;
  const uint32_t* codes32 = reinterpret_cast<const uint32_t*>(P.codes) + ((lane * 4) >> 5);
  const int cstride = (post >> 5) * 4 * 4;
  const int bstride = (post >> 5) * 4;
  const int sb = post * 4;
#pragma unroll 1
  for (int sub = 0; sub < MSQ_SEG_BLOCKS / 256; ++sub) {
    const int t0 = seg * MSQ_SEG_BLOCKS + sub * 256 + (int)threadIdx.x;
    const bool active = t0 < (int)P.nblocks;
    int base = lane, b = 0;                   // inactive (past the last block): column `lane` of the first row block
    if (active) {
      const int pb = fast_div(t0, P.post_m, P.post_k);
      const int j = t0 - pb * post;
      const int p = fast_div(pb, P.nblk_m, P.nblk_k); b = pb - p * P.nblk;
      base = (p * P.L + b * BS) * post + j;
    }
    // (an inactive thread -- past the last block -- reads the first row block and stores nothing)
    const int ei = active ? P.e_in[t0] : 0, eo = active ? P.e_out[t0] : 0;
    const int row32 = (base - lane) >> 5;
    const char* cp = reinterpret_cast<const char*>(codes32 + (long long)row32 * 4);
    const char* bp = reinterpret_cast<const char*>(P.bitmap + row32);
    // slot nibbles of rows 0-7 / 8-15 and the 16 mask bits: every row shifts its nibble / bit in from the top
    uint32_t s0 = 0, s1 = 0, mask = 0;
#pragma unroll
    for (int i = 0; i < BS; ++i) {
      const uint32_t cw = __ldg(reinterpret_cast<const uint32_t*>(row_at(cp, cstride, i)));
      const uint32_t bw = __ldg(reinterpret_cast<const uint32_t*>(row_at(bp, bstride, i)));
      if (i < 8) s0 = __funnelshift_r(s0, cw >> cshift, 4);
      else s1 = __funnelshift_r(s1, cw >> cshift, 4);
      mask = __funnelshift_r(mask, bw >> lane, 1);
    }
    if (!active) mask = 0;
    mask >>= 16;
    // position of this block's hi nibbles in the segment stream
    int wt;
    const int k = s_off[sub * 8 + warp] + warp_excl_scan(__popc(mask), lane, wt);
    const int bit = k * hw;
    const uint32_t w0 = s_hi[bit >> 5], w1 = s_hi[(bit >> 5) + 1], w2 = s_hi[(bit >> 5) + 2];
    uint32_t hlo = __funnelshift_r(w0, w1, bit & 31), hhi = __funnelshift_r(w1, w2, bit & 31);   // 16 nibbles, first outlier lowest
    // deal them to the outlier positions: element i takes the lowest nibble, the window moves on by 4 bits after an outlier
    const uint32_t nm0 = nibble_mask8u(mask & 0xFFu), nm1 = nibble_mask8u(mask >> 8);
    uint32_t h0 = 0, h1 = 0;
    {
      const uint32_t a0 = nm0 & 0x44444444u, a1 = nm1 & 0x44444444u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h0 = __funnelshift_r(h0, hlo, 4);
        const uint32_t sh = a0 >> (4 * i);            // 4 for an outlier, 0 for an inlier (SHF takes it modulo 32)
        hlo = __funnelshift_r(hlo, hhi, sh);
        hhi >>= (sh & 31u);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h1 = __funnelshift_r(h1, hlo, 4);
        const uint32_t sh = a1 >> (4 * i);
        hlo = __funnelshift_r(hlo, hhi, sh);
        hhi >>= (sh & 31u);
      }
    }
    h0 &= nm0; h1 &= nm1;
    const uint32_t in0 = s0 & ~nm0, in1 = s1 & ~nm1;            // inlier slots (outlier positions zeroed: decode 0)
    const uint32_t lo0 = s0 & nm0, lo1 = s1 & nm1;              // low nibbles of the outlier codes
    const bool nan_blk = ei == MSQ_EXP_NAN || eo == MSQ_EXP_NAN;
    const float s_in = pow2i(nan_blk ? 0 : ei), inv_s_in = pow2i(nan_blk ? 0 : -ei), s_out = pow2i(nan_blk ? 0 : eo);
    float r[BS];
    const float2 si2 = make_float2(s_in, s_in), so2 = make_float2(s_out, s_out), ii2 = make_float2(inv_s_in, inv_s_in);
#pragma unroll
    for (int g = 0; g < 4; ++g) {                                // four elements at a time
      const uint32_t inw = g < 2 ? in0 : in1, low = g < 2 ? lo0 : lo1, hiw = g < 2 ? h0 : h1;
      const int q = (g & 1) * 16;
      const uint32_t ob = spread4((low >> q) & 0xFFFFu) | (spread4((hiw >> q) & 0xFFFFu) << 4);     // four E4M3 bytes
      const float2 i01 = dec2_e2m1((inw >> q) & 0xFFu), i23 = dec2_e2m1((inw >> (q + 8)) & 0xFFu);
      const float2 o01 = dec2_e4m3(ob & 0xFFFFu), o23 = dec2_e4m3(ob >> 16);
      // inlier: q * 2^e_in; outlier: (q * 2^e_out) * 2^-e_in (quant.py:136); exactly one of the two terms is non-zero and
      // the sum of a product and a zero is the reference's "+ 0" (quant.py:140)
      const float2 a = __fadd2_rn(__fmul2_rn(i01, si2), __fmul2_rn(__fmul2_rn(o01, so2), ii2));
      const float2 c = __fadd2_rn(__fmul2_rn(i23, si2), __fmul2_rn(__fmul2_rn(o23, so2), ii2));
      r[4 * g + 0] = a.x; r[4 * g + 1] = a.y; r[4 * g + 2] = c.x; r[4 * g + 3] = c.y;
    }
    if (nan_blk) {
#pragma unroll
      for (int i = 0; i < BS; ++i) r[i] = qnan();
    }
    if (active) {
      if (!BF16OUT) {
        float* o = P.out_f32 + base;
#pragma unroll
        for (int i = 0; i < BS; ++i) *reinterpret_cast<float*>(const_cast<char*>(row_at(o, sb, i))) = r[i];
      } else {
        __nv_bfloat16* o = P.out_bf16 + base;
#pragma unroll
        for (int i = 0; i < BS; ++i)
          *reinterpret_cast<__nv_bfloat16*>(const_cast<char*>(row_at(o, sb >> 1, i))) = __float2bfloat16_rn(r[i]);
      }
    }
  }
}

// innermost layout: thread = 4 consecutive elements, bs/4 lanes per block
__global__ void __launch_bounds__(256) k_unpack_inner(const UParams P) {
  extern __shared__ uint32_t s_hi[];
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const long long seg = blockIdx.x;
  const bool has_hi = !P.plain && P.hi_w > 0;
  const int wi = P.fin.width;
  const int BS = P.bs;
  if (has_hi) {
    const int words = (int)(P.hi_seg_stride >> 2);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.hi + seg * P.hi_seg_stride);
    for (int i = threadIdx.x; i < words; i += 256) s_hi[i] = __ldg(src + i);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
  }
  const long long nelem = P.nblocks * BS;
  const long long seg_e0 = seg * MSQ_SEG_BLOCKS * BS;
  for (int it = 0; it < BS; ++it) {
    const long long e0 = seg_e0 + (long long)it * 1024 + threadIdx.x * 4;
    const bool active = e0 < nelem;
    uint32_t nib = 0;
    if (active && !P.plain) nib = (__ldg(P.bitmap + (e0 >> 5)) >> (e0 & 31)) & 15u;
    int k = 0;
    if (has_hi) {
      int tot;
      const int off = cta_excl_scan(__popc(nib), s_warp, &tot);
      k = s_base + off;
      __syncthreads();
      if (threadIdx.x == 0) s_base += tot;
      __syncthreads();
    }
    if (!active) continue;
    const long long t = e0 / BS;
    const int ei = P.e_in[t], eo = P.plain ? 0 : P.e_out[t];
    const bool nan_blk = ei == MSQ_EXP_NAN || eo == MSQ_EXP_NAN;
    const float s_in = pow2i(nan_blk ? 0 : ei), inv_s_in = pow2i(nan_blk ? 0 : -ei), s_out = pow2i(nan_blk ? 0 : eo);
    const long long bit0 = e0 * wi;
    uint32_t cw;
    if (wi == 2) cw = P.codes[bit0 >> 3];
    else if (wi == 4) cw = *reinterpret_cast<const uint16_t*>(P.codes + (bit0 >> 3));
    else cw = *reinterpret_cast<const uint32_t*>(P.codes + (bit0 >> 3));
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t code = (cw >> (i * wi)) & ((1u << wi) - 1u);
      const bool isout = (nib >> i) & 1u;
      if (isout && has_hi) { code |= hi_field(s_hi, k, P.hi_w) << wi; ++k; }
      r[i] = dq_elem(code, isout, P, s_in, inv_s_in, s_out, nan_blk);
    }
    if (P.out_f32) *reinterpret_cast<float4*>(P.out_f32 + e0) = make_float4(r[0], r[1], r[2], r[3]);
    if (P.out_bf16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]), b2 = __floats2bfloat162_rn(r[2], r[3]);
      uint2 v;
      v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b2);
      *reinterpret_cast<uint2*>(P.out_bf16 + e0) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fast innermost unpack (blocks along the contiguous axis, the layout of GEMM weight operands): thread = 8
// consecutive elements (one 32-bit code word for 4-bit slots, one bitmap byte, one 16/32-byte store), LUT
// decode, hi offsets from a prologue prefix sum over the segment's bitmap words (no synchronisation in the
// main loop).
// ---------------------------------------------------------------------------------------------------------
template <int HWT>
__global__ void __launch_bounds__(256) k_unpack_inner_fast(const UParams P) {
  constexpr bool HAS_HI = HWT != 0;
  const int hw = HWT >= 0 ? HWT : P.hi_w;
  extern __shared__ uint32_t s_dyn[];
  __shared__ int s_warp[8];
  const int wi = P.fin.width;
  const int wo = wi + hw;
  const int BS = P.bs;
  float* lut = reinterpret_cast<float*>(s_dyn);
  const int lut_n = (1 << wi) + (1 << wo);
  int* s_pref = reinterpret_cast<int*>(s_dyn + lut_n);          // exclusive outlier count before each bitmap word
  const int nwords = (MSQ_SEG_BLOCKS * BS) >> 5;                // bitmap words per segment (BS % 8 == 0 => integer)
  uint32_t* s_hi = s_dyn + lut_n + (HAS_HI ? nwords + 1 : 0);
  const long long seg = blockIdx.x;
  const long long nelem = P.nblocks * BS;
  const long long seg_e0 = seg * MSQ_SEG_BLOCKS * BS;
  for (int i = threadIdx.x; i < lut_n; i += 256)
    lut[i] = (i < (1 << wi)) ? decode_elem((uint32_t)i, P.fin.ebits, P.fin.bits)
                             : decode_elem((uint32_t)(i - (1 << wi)), P.fout.ebits, P.fout.bits);
  if (HAS_HI) {
    // each thread owns nwords/256 consecutive words (nwords is a multiple of 256: 32*BS words, BS % 8 == 0)
    const int per = nwords >> 8;
    const uint32_t* bm = P.bitmap + (seg_e0 >> 5);
    const long long words_total = (nelem + 31) >> 5;
    int sum = 0;
    for (int j = 0; j < per; ++j) {
      const int w = threadIdx.x * per + j;
      sum += ((seg_e0 >> 5) + w < words_total) ? __popc(__ldg(bm + w)) : 0;
    }
    int tot;
    int run = cta_excl_scan(sum, s_warp, &tot);
    for (int j = 0; j < per; ++j) {
      const int w = threadIdx.x * per + j;
      s_pref[w] = run;
      run += ((seg_e0 >> 5) + w < words_total) ? __popc(__ldg(bm + w)) : 0;
    }
    const int words = (tot * hw + 31) >> 5;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.hi + seg * P.hi_seg_stride);
    for (int i = threadIdx.x; i < words; i += 256) s_hi[i] = __ldg(src + i);
  }
  __syncthreads();
  const uint32_t cm = (1u << wi) - 1u, hm = (1u << hw) - 1u;
  const int iters = (MSQ_SEG_BLOCKS * BS) >> 11;                // 2048 elements per CTA iteration
#pragma unroll 2
  for (int it = 0; it < iters; ++it) {
    const long long e0 = seg_e0 + (long long)it * 2048 + threadIdx.x * 8;
    if (e0 >= nelem) continue;
    uint32_t bword = 0;
    if (!P.plain) bword = __ldg(P.bitmap + (e0 >> 5));
    const int sh = (int)(e0 & 31);
    const uint32_t byte = (bword >> sh) & 0xFFu;
    uint64_t cw;
    if (wi == 4) cw = __ldg(reinterpret_cast<const uint32_t*>(P.codes + (e0 >> 1)));
    else if (wi == 2) cw = __ldg(reinterpret_cast<const uint16_t*>(P.codes + (e0 >> 2)));
    else { const uint2 v = __ldg(reinterpret_cast<const uint2*>(P.codes + e0)); cw = (uint64_t)v.x | ((uint64_t)v.y << 32); }
    const long long t = e0 / BS;
    const int ei = P.e_in[t], eo = P.plain ? 0 : P.e_out[t];
    const bool nan_blk = ei == MSQ_EXP_NAN || eo == MSQ_EXP_NAN;
    const float s_in = pow2i(nan_blk ? 0 : ei), inv_s_in = pow2i(nan_blk ? 0 : -ei), s_out = pow2i(nan_blk ? 0 : eo);
    int k = 0;
    if (HAS_HI) k = s_pref[(int)((e0 - seg_e0) >> 5)] + __popc(bword & ((1u << sh) - 1u));
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t c = (uint32_t)(cw >> (j * wi)) & cm;
      const uint32_t isout = (byte >> j) & 1u;
      if (HAS_HI) {
        const int bit = (k + __popc(byte & ((1u << j) - 1u))) * hw;
        const uint32_t f = (s_hi[bit >> 5] >> (bit & 31)) & hm;
        c |= (f & (0u - isout)) << wi;
      }
      const float q = lut[c + (isout << wi)];
      const float v = isout ? __fmul_rn(__fmul_rn(q, s_out), inv_s_in) : __fmul_rn(q, s_in);
      r[j] = nan_blk ? qnan() : __fadd_rn(v, 0.0f);
    }
    if (P.out_f32) {
      float4* o = reinterpret_cast<float4*>(P.out_f32 + e0);
      o[0] = make_float4(r[0], r[1], r[2], r[3]);
      o[1] = make_float4(r[4], r[5], r[6], r[7]);
    }
    if (P.out_bf16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]), b2 = __floats2bfloat162_rn(r[2], r[3]);
      __nv_bfloat162 c2 = __floats2bfloat162_rn(r[4], r[5]), d2 = __floats2bfloat162_rn(r[6], r[7]);
      uint4 v;
      v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b2);
      v.z = *reinterpret_cast<uint32_t*>(&c2); v.w = *reinterpret_cast<uint32_t*>(&d2);
      *reinterpret_cast<uint4*>(P.out_bf16 + e0) = v;
    }
  }
}

}  // namespace msq

using namespace msq;

extern "C" {

// Inverse of the PACK outputs of msq_quantize_mx_outlier_f32 / msq_quantize_mx_f32 (plain != 0: no
// bitmap / e_out / hi).  Exactly one or both of out_f32 / out_bf16 may be given.
int msq_unpack_mx(const unsigned char* codes, const signed char* e_in, const signed char* e_out, const unsigned int* bitmap,
                  const unsigned char* hi, float* out_f32, void* out_bf16, long long pre, long long L, long long post,
                  int block_size, int in_ebits, int in_mbits, int in_emax, float in_max_norm, int out_ebits, int out_mbits,
                  int out_emax, float out_max_norm, int plain, void* stream) {
  msq::DevGuard dev_guard_(codes);
  if (!codes || !e_in || (!out_f32 && !out_bf16) || pre <= 0 || L <= 0 || post <= 0 || block_size <= 0) return MSQ_ERR_BAD_ARG;
  if (!plain && (!bitmap || !e_out)) return MSQ_ERR_BAD_ARG;
  if (pre * L * post >= 0x7fffffffLL) return MSQ_ERR_UNSUPPORTED;
  UParams P;
  memset(&P, 0, sizeof(P));
  P.pre = pre; P.L = (int)L; P.post = post; P.bs = block_size;
  P.nblk = (int)((L + block_size - 1) / block_size);
  P.nblocks = pre * P.nblk * post;
  P.fin = make_fmt(in_ebits, in_mbits, in_emax, in_max_norm);
  P.fout = plain ? P.fin : make_fmt(out_ebits, out_mbits, out_emax, out_max_norm);
  P.codes = codes; P.e_in = e_in; P.e_out = e_out; P.bitmap = bitmap; P.hi = hi; P.plain = plain;
  P.hi_w = plain ? 0 : msq_packed_hi_width(in_ebits, in_mbits, out_ebits, out_mbits);
  if (P.hi_w > 0 && !hi) return MSQ_ERR_BAD_ARG;
  P.hi_seg_stride = msq_packed_hi_seg_stride(block_size, P.hi_w);
  magic_div(post, &P.post_m, &P.post_k);
  magic_div(P.nblk, &P.nblk_m, &P.nblk_k);
  P.out_f32 = out_f32; P.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  const long long nseg = (P.nblocks + MSQ_SEG_BLOCKS - 1) / MSQ_SEG_BLOCKS;
  const size_t smem = P.hi_w > 0 ? (size_t)P.hi_seg_stride : 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (post == 1 && L % block_size == 0 && block_size % 8 == 0 && P.fin.width + P.hi_w <= 8) {
    // blocks of 8k elements: every 8-element group sits inside one block and one bitmap word
    const size_t lut_bytes = ((size_t)(1 << P.fin.width) + ((size_t)1 << (P.fin.width + P.hi_w))) * sizeof(float);
    const size_t pref_bytes = P.hi_w > 0 ? ((size_t)((MSQ_SEG_BLOCKS * block_size) >> 5) + 1) * sizeof(int) : 0;
    const size_t sm = lut_bytes + pref_bytes + smem;
    void (*k)(const UParams) = P.hi_w == 0 ? k_unpack_inner_fast<0> : (P.hi_w == 4 ? k_unpack_inner_fast<4> : k_unpack_inner_fast<-1>);
    if (sm > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k<<<(unsigned)nseg, 256, sm, s>>>(P);
  } else if (post == 1 && L % block_size == 0 && block_size % 4 == 0) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_unpack_inner, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_unpack_inner<<<(unsigned)nseg, 256, smem, s>>>(P);
  } else if (post > 1 && post % 32 == 0 && (block_size == 8 || block_size == 16 || block_size == 32) &&
             P.fin.width + P.hi_w <= 8) {
    const size_t lut_bytes = ((size_t)(1 << P.fin.width) + ((size_t)1 << (P.fin.width + P.hi_w))) * sizeof(float);
    const size_t sm = lut_bytes + smem;
    void (*k)(const UParams) = nullptr;
    const bool hh = P.hi_w > 0;
    const int hwt = P.hi_w == 0 ? 0 : (P.hi_w == 4 ? 4 : -1);
#define MSQ_PICK_U(BS_) k = hwt == 0 ? k_unpack_strided_fast<BS_, 0> : (hwt == 4 ? k_unpack_strided_fast<BS_, 4> : k_unpack_strided_fast<BS_, -1>)
    if (block_size == 16) MSQ_PICK_U(16);
    else if (block_size == 32) MSQ_PICK_U(32);
    else MSQ_PICK_U(8);
#undef MSQ_PICK_U
    (void)hh;
    // the drivers' pairing on packed words + hardware decode (MSQ_UNPACK_CVT=0 keeps the table kernel: A/B)
    const bool cvt = !plain && block_size == 16 && P.hi_w == 4 && P.fin.ebits == 2 && P.fin.bits == 3 && P.fout.ebits == 4 &&
                     P.fout.bits == 5 && L % 16 == 0 && (out_f32 == nullptr) != (out_bf16 == nullptr) &&
                     !(getenv("MSQ_UNPACK_CVT") && getenv("MSQ_UNPACK_CVT")[0] == '0');
    if (cvt) {
      const size_t smc = smem + 16;
      void (*kc)(const UParams) = out_f32 ? k_unpack_strided_cvt<false> : k_unpack_strided_cvt<true>;
      if (smc > 48 * 1024) cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc);
      kc<<<(unsigned)nseg, 256, smc, s>>>(P);
      return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
    }
    if (sm > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k<<<(unsigned)nseg, 256, sm, s>>>(P);
  } else if (post > 1 && post % 32 == 0 && block_size <= 32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_unpack_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_unpack_strided<<<(unsigned)nseg, 256, smem, s>>>(P);
  } else {
    return MSQ_ERR_UNSUPPORTED;
  }
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Packed planes from per-element codes: the GPTQ column kernel (msq_gptq_solve_f32) emits one code byte per weight
// (inlier or outlier format according to the bitmap) -- the quantizer's own integers, not a re-quantization of the
// dequantized values.  This pass lays them out as the standard planes: inlier-width slots + segment-compacted hi
// stream, so that unpack(planes) == GPTQ's Q bit for bit and the row-block GEMM can consume them.
// ---------------------------------------------------------------------------------------------------------------
namespace msq {

__global__ void __launch_bounds__(256) k_slots_from_codes(const uint8_t* __restrict__ codes8, long long n, int wi,
                                                          uint8_t* __restrict__ slots) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;      // group of 8 elements
  const long long e0 = g * 8;
  if (e0 >= n) return;
  uint32_t w = 0;
  const uint32_t m = (1u << wi) - 1u;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (e0 + i < n) w |= ((uint32_t)codes8[e0 + i] & m) << (wi * i);
  uint8_t* dst = slots + ((e0 * wi) >> 3);
  const int nbytes = wi;                                               // 8 elements * wi bits
  for (int b = 0; b < nbytes; ++b)
    if (((e0 * wi) >> 3) + b < ((n * wi + 7) >> 3)) dst[b] = (uint8_t)(w >> (8 * b));
}

__global__ void __launch_bounds__(256) k_hi_from_codes(const uint8_t* __restrict__ codes8, const uint32_t* __restrict__ bitmap,
                                                       long long pre, long long L, long long post, int bs, long long nb, int wi,
                                                       int hi_w, uint8_t* __restrict__ hi, long long hi_seg_stride,
                                                       int* __restrict__ seg_cnt) {
  __shared__ uint32_t wsum[8];
  const long long nblk = (L + bs - 1) / bs;
  const long long t0 = (long long)blockIdx.x * MSQ_SEG_BLOCKS + threadIdx.x * 4;
  uint32_t c[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long t = t0 + u;
    c[u] = 0;
    if (t < nb) {
      const long long j = t % post, pb = t / post, b = pb % nblk, p = pb / nblk;
      for (int i = 0; i < bs; ++i) {
        const long long l = b * bs + i;
        if (l >= L) break;
        const long long e = (p * L + l) * post + j;
        c[u] += (__ldg(bitmap + (e >> 5)) >> (e & 31)) & 1u;
      }
    }
  }
  const uint32_t mine = c[0] + c[1] + c[2] + c[3];
  uint32_t incl = mine;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t base = 0, total = 0;
  for (int w = 0; w < 8; ++w) { if (w < warp) base += wsum[w]; total += wsum[w]; }
  if (threadIdx.x == 0 && seg_cnt) seg_cnt[blockIdx.x] = hi_w ? (int)total : 0;
  if (!hi_w) return;
  uint32_t off = base + incl - mine;
  uint32_t* dst = reinterpret_cast<uint32_t*>(hi + (long long)blockIdx.x * hi_seg_stride);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long t = t0 + u;
    if (t >= nb || !c[u]) continue;
    const long long j = t % post, pb = t / post, b = pb % nblk, p = pb / nblk;
    for (int i = 0; i < bs; ++i) {
      const long long l = b * bs + i;
      if (l >= L) break;
      const long long e = (p * L + l) * post + j;
      if ((__ldg(bitmap + (e >> 5)) >> (e & 31)) & 1u) {
        const uint32_t field = ((uint32_t)codes8[e] >> wi) & ((1u << hi_w) - 1u);
        const uint32_t bit = off * hi_w;
        atomicOr(dst + (bit >> 5), field << (bit & 31));            // hi_w divides 32: a field never straddles words
        ++off;
      }
    }
  }
}

}  // namespace msq

extern "C" int msq_pack_from_codes(const unsigned char* codes8, const unsigned int* bitmap, long long pre, long long L,
                                   long long post, int block_size, int in_ebits, int in_mbits, int out_ebits, int out_mbits,
                                   unsigned char* slots, unsigned char* hi, int* seg_cnt, void* stream) {
  msq::DevGuard dev_guard_(codes8);
  if (!codes8 || !bitmap || !slots || pre <= 0 || L <= 0 || post <= 0 || block_size <= 0) return MSQ_ERR_BAD_ARG;
  using namespace msq;
  const Fmt fi = make_fmt(in_ebits, in_mbits, 0, 1.0f);
  const int wi = fi.width, hi_w = msq_packed_hi_width(in_ebits, in_mbits, out_ebits, out_mbits);
  if (hi_w && !hi) return MSQ_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = pre * L * post;
  const long long nblk = (L + block_size - 1) / block_size, nb = pre * nblk * post;
  const long long nseg = (nb + MSQ_SEG_BLOCKS - 1) / MSQ_SEG_BLOCKS;
  const long long stride = msq_packed_hi_seg_stride(block_size, hi_w);
  if (hi_w && cudaMemsetAsync(hi, 0, (size_t)(nseg * stride), s) != cudaSuccess) return MSQ_ERR_CUDA;
  if (wi == 8) {
    if (cudaMemcpyAsync(slots, codes8, (size_t)n, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return MSQ_ERR_CUDA;
  } else {
    k_slots_from_codes<<<(unsigned)((n / 8 + 256) / 256), 256, 0, s>>>(codes8, n, wi, slots);
    if (cudaGetLastError() != cudaSuccess) return MSQ_ERR_CUDA;
  }
  k_hi_from_codes<<<(unsigned)nseg, 256, 0, s>>>(codes8, bitmap, pre, L, post, block_size, nb, wi, hi_w, hi, stride, seg_cnt);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}
