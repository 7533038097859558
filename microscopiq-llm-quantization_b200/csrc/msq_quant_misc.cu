// msq_quant_misc.cu -- generic (any geometry) quantizer kernel, detector statistics of the
// mx_ops variant, element-wise rounding, and the C entry points + kernel dispatch.
#include "msq_quant_kernels.cuh"

namespace msq {

// --------------------------------------------------------------------------
// generic mapping: one thread per block, any bs / ragged tail / any post.
// Supports deq, mask, e_in, e_out, blk_cnt, status (no packed outputs).
// --------------------------------------------------------------------------
template <bool OUTLIER>
__global__ void __launch_bounds__(256) k_generic(const QParams P) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= P.nblocks) return;
  const QCfg& c = P.c;
  const long long pb = t / P.post, j = t - pb * P.post;
  const long long p = pb / P.nblk;
  const int b = (int)(pb - p * P.nblk);
  const long long base = (p * P.L + (long long)b * P.bs) * P.post + j;
  const int nv = min(P.bs, P.L - b * P.bs);          // valid rows; the rest is zero padding
  const int bs = P.bs;
  int st = 0;
  int bad = 0;
  float m = 0.0f;
  for (int i = 0; i < nv; ++i) {
    float a = fabsf(__ldg(P.in + base + (long long)i * P.post));
    bad |= !(a <= 3.402823466e38f);
    m = fmaxf(m, a);
  }
  if (!OUTLIER) {
    // plain MX through the shared core, one element at a time
    if (P.ext_max) { m = fabsf(P.ext_max[t]); bad = !(m <= 3.402823466e38f); }
    float mm[1] = {bad ? qnan() : m};
    for (int i = 0; i < nv; ++i) {
      float xv[1] = {__ldg(P.in + base + (long long)i * P.post)};
      uint32_t cd[1];
      PlainOut po = core_plain<1, 1>(xv, cd, false, c, mm);
      st |= po.st;
      if (P.deq) P.deq[base + (long long)i * P.post] = xv[0];
      if (i == 0 && P.e_in) P.e_in[t] = (int8_t)po.e;
    }
    if (nv == 0 && P.e_in) P.e_in[t] = 0;
    if (P.status && st) atomicOr(P.status, st);
    return;
  }
  float lo = 0.f, hi = 0.f;
  if (c.variant == 0) {
    // canonical fp32 pairwise-tree sums over the bs block entries (binary-counter form: identical
    // association to the adjacent tree with the summands zero-padded to a power of two)
    float lvl[32];
    auto tree_total = [&](auto term) {
      unsigned cnt = 0;
      for (int i = 0; i < bs; ++i) {
        float carry = term(i);
        int l = 0;
        for (unsigned cc = cnt; cc & 1u; cc >>= 1, ++l) carry = __fadd_rn(lvl[l], carry);
        lvl[l] = carry;
        ++cnt;
      }
      float acc = 0.0f;
      bool have = false;
      for (int l = 0; l < 32; ++l)
        if ((cnt >> l) & 1u) { acc = have ? __fadd_rn(lvl[l], acc) : lvl[l]; have = true; }
      return acc;
    };
    auto absval = [&](int i) { return i < nv ? fabsf(__ldg(P.in + base + (long long)i * P.post)) : 0.0f; };
    const float ssum = tree_total(absval);
    if (!(ssum <= 3.402823466e38f)) bad = 1;
    const float mean = __fdiv_rn(ssum, (float)bs);
    const float q = tree_total([&](int i) { const float d = __fsub_rn(absval(i), mean); return __fmul_rn(d, d); });
    const float sd = __fsqrt_rn(__fdiv_rn(q, (float)bs));
    const float w = __fmul_rn(c.std_dev, sd);
    lo = __fsub_rn(mean, w); hi = __fadd_rn(mean, w);
  }
  auto bounds = [&](int i, float& l, float& h) {
    if (c.variant == 0) { l = lo; h = hi; }
    else { long long o = (p * bs + i) * P.post + j; l = P.lo_ext[o]; h = P.hi_ext[o]; }
  };
  int cnt = 0;
  float m_in = 0.0f;
  for (int i = 0; i < bs; ++i) {
    float xv = i < nv ? __ldg(P.in + base + (long long)i * P.post) : 0.0f;
    float l, h; bounds(i, l, h);
    bool isout = (xv < l) | (xv > h);
    cnt += isout;
    if (!isout) m_in = fmaxf(m_in, fabsf(xv));
  }
  const int E_in = (m_in == 0.0f) ? -126 : floor_log2f(m_in);
  const bool flush = c.flush && (E_in <= -127);
  int e_in = clamp_scale_exp(E_in - c.fin.emax, c.sb_in, c.variant == 0);
  if (bad) st |= MSQ_ST_NONFINITE_INPUT;
  if (e_in == MSQ_EXP_NAN) st |= MSQ_ST_EXP_IN_OVERFLOW;
  int e_out = MSQ_EXP_NAN;
  float s_in = qnan(), inv_s_in = qnan(), s_out = qnan(), inv_s_out = qnan();
  bool dead = bad || e_in == MSQ_EXP_NAN;
  if (!dead) {
    s_in = pow2i(e_in); inv_s_in = pow2i(-e_in);
    float m_out = 0.0f;
    for (int i = 0; i < bs; ++i) {
      float xv = i < nv ? __ldg(P.in + base + (long long)i * P.post) : 0.0f;
      float l, h; bounds(i, l, h);
      if ((xv < l) | (xv > h)) m_out = fmaxf(m_out, fabsf(__fmul_rn(xv, s_in)));
    }
    const int E_out = (m_out == 0.0f) ? -126 : floor_log2f(m_out);
    e_out = clamp_scale_exp(E_out - c.fout.emax, c.sb_out, c.variant == 0);
    if (e_out == MSQ_EXP_NAN) { st |= MSQ_ST_EXP_OUT_OVERFLOW; dead = true; }
    else { s_out = pow2i(e_out); inv_s_out = pow2i(-e_out); }
  }
  for (int i = 0; i < nv; ++i) {
    const long long lin = base + (long long)i * P.post;
    float xv = __ldg(P.in + lin);
    float l, h; bounds(i, l, h);
    const bool isout = (xv < l) | (xv > h);
    float r;
    if (dead) r = qnan();
    else if (isout) {
      float v = __fmul_rn(__fmul_rn(xv, s_in), inv_s_out);
      r = __fmul_rn(__fmul_rn(qelem_rt(v, c.fout, c.rmode), s_out), inv_s_in);
    } else {
      float v = __fmul_rn(flush ? __fmul_rn(xv, 0.0f) : xv, inv_s_in);
      r = __fmul_rn(qelem_rt(v, c.fin, c.rmode), s_in);
    }
    if (P.deq) P.deq[lin] = __fadd_rn(r, 0.0f);
    if (P.mask) P.mask[lin] = isout ? 1 : 0;
  }
  if (P.e_in) P.e_in[t] = (int8_t)e_in;
  if (P.e_out) P.e_out[t] = (int8_t)e_out;
  if (P.blk_cnt) P.blk_cnt[t] = (uint8_t)cnt;
  if (P.status && st) atomicOr(P.status, st);
}

// --------------------------------------------------------------------------
// variant-1 detector statistics (mx_ops.py:62-83 as called at :248): mean and
// UNBIASED std of signed A over the block-count axis, per in-block position.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bounds_across_blocks(const float* __restrict__ in, float* __restrict__ lo,
                                                              float* __restrict__ hi, long long pre, int L, long long post,
                                                              int bs, int nblk, float std_dev) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;   // over [pre, bs, post]
  if (t >= pre * bs * post) return;
  const long long j = t % post;
  const int i = (int)((t / post) % bs);
  const long long p = t / (post * bs);
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) {
    int l = b * bs + i;
    if (l < L) s += (double)__ldg(in + (p * L + l) * post + j);
  }
  const double mean64 = s / (double)nblk;
  double q = 0.0;
  for (int b = 0; b < nblk; ++b) {
    int l = b * bs + i;
    double d = (l < L ? (double)__ldg(in + (p * L + l) * post + j) : 0.0) - mean64;
    q = fma(d, d, q);
  }
  const float mean = (float)mean64;
  const float sd = (float)sqrt(q / (double)(nblk - 1));
  const float w = __fmul_rn(std_dev, sd);
  lo[t] = __fsub_rn(mean, w);
  hi[t] = __fadd_rn(mean, w);
}

// --------------------------------------------------------------------------
// element-wise rounding (quantize_elemwise_func_cuda, funcs.cpp:178-195)
// --------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, int VEC>
__global__ void __launch_bounds__(256) k_elemwise(const T* __restrict__ in, T* __restrict__ out, long long n, Fmt f,
                                                  int rmode, int saturate, int allow_denorm, int native) {
  const long long stride = (long long)gridDim.x * 256 * VEC;
  for (long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC; i0 < n; i0 += stride) {
    T v[VEC];
    if (i0 + VEC <= n) {
      if (sizeof(T) * VEC == 16) *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(in + i0));
      else if (sizeof(T) * VEC == 8) *reinterpret_cast<uint2*>(v) = __ldg(reinterpret_cast<const uint2*>(in + i0));
      else
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = in[i0 + k];
    } else {
      for (int k = 0; k < VEC; ++k) v[k] = (i0 + k < n) ? in[i0 + k] : from_f32<T>(0.0f);
    }
    auto slow = [&](float x) {
      if (native) return qelem_native(x, f.bits, f.ebits, f.max_norm, rmode, saturate, allow_denorm);
      if (rmode == RD_NEAREST) return qelem<RD_NEAREST>(x, f, saturate, allow_denorm);
      if (rmode == RD_FLOOR) return qelem<RD_FLOOR>(x, f, saturate, allow_denorm);
      return qelem<RD_EVEN>(x, f, saturate, allow_denorm);
    };
    if (VEC % 2 == 0 && (native || rmode != RD_EVEN)) {
      // packed fp32x2 rounding (see core_outlier); values the magic-number rounding cannot hold (|x| >= 2^100,
      // Inf, NaN) take the literal path
      const float Dp = f.ebits ? f.down : 0.0f, LD = __fmul_rn(f.L, f.down);
      const float addend = (rmode == RD_FLOOR) ? 0.0f : 0.5f;
      const float inf = __int_as_float(0x7F800000);
#pragma unroll
      for (int k = 0; k < VEC; k += 2) {
        const float2 x = f2(to_f32<T>(v[k]), to_f32<T>(v[k + 1]));
        float2 y;
        if (fabsf(x.x) < 1.2676506e30f && fabsf(x.y) < 1.2676506e30f) {
          float2 a = x;
          if (!allow_denorm && f.ebits) {                      // elemwise_ops.py:139-141 / quantize.cuh:104
            a.x = (fabsf(a.x) >= f.L) ? a.x : (native ? 0.0f : __fmul_rn(0.0f, a.x));
            a.y = (fabsf(a.y) >= f.L) ? a.y : (native ? 0.0f : __fmul_rn(0.0f, a.y));
          }
          const float2 qn = quantum2(a, f2s(Dp), f2s(LD));
          const float2 big = f2s(3.0e38f);
          float2 q = native ? round_mag2_native(a, qn, rmode, big) : round_mag2_py(a, qn, addend, big);
          if (saturate || f.ebits == 0) { q.x = fminf(q.x, f.max_norm); q.y = fminf(q.y, f.max_norm); }
          else { q.x = q.x > f.max_norm ? inf : q.x; q.y = q.y > f.max_norm ? inf : q.y; }
          y.x = copysignf(q.x, a.x); y.y = copysignf(q.y, a.y);
          if (native) y = __fadd2_rn(y, f2s(0.0f));            // quantize.cuh returns +0 for a zero mantissa
        } else {
          y.x = slow(x.x); y.y = slow(x.y);
        }
        v[k] = from_f32<T>(y.x); v[k + 1] = from_f32<T>(y.y);
      }
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = from_f32<T>(slow(to_f32<T>(v[k])));
    }
    if (i0 + VEC <= n) {
      if (sizeof(T) * VEC == 16) *reinterpret_cast<uint4*>(out + i0) = *reinterpret_cast<uint4*>(v);
      else if (sizeof(T) * VEC == 8) *reinterpret_cast<uint2*>(out + i0) = *reinterpret_cast<uint2*>(v);
      else
#pragma unroll
        for (int k = 0; k < VEC; ++k) out[i0 + k] = v[k];
    } else {
      for (int k = 0; k < VEC; ++k) if (i0 + k < n) out[i0 + k] = v[k];
    }
  }
}

// --------------------------------------------------------------------------
// host-side dispatch
// --------------------------------------------------------------------------
cudaError_t launch_generic(const QParams& P, bool outlier, cudaStream_t stream) {
  const unsigned g = (unsigned)((P.nblocks + 255) / 256);
  if (outlier) k_generic<true><<<g, 256, 0, stream>>>(P);
  else k_generic<false><<<g, 256, 0, stream>>>(P);
  return cudaGetLastError();
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int launch_quant(QParams& P, bool outlier, cudaStream_t stream, int* path_out) {
  const bool packing = P.codes || P.bitmap || P.hi;
  if (packing) {
    // the packed planes are produced together: codes + exponents (+ bitmap, + hi stream when
    // outlier codes are wider than the inlier slot)
    if (!P.codes || !P.e_in) return MSQ_ERR_BAD_ARG;
    if (outlier && (!P.bitmap || !P.e_out || (P.hi_w > 0 && !P.hi))) return MSQ_ERR_BAD_ARG;
    if (!aligned16(P.codes) || !aligned16(P.bitmap) || !aligned16(P.hi)) return MSQ_ERR_ALIGN;
    if (outlier && !P.c.enc2_ok) return MSQ_ERR_UNSUPPORTED;     // codes wider than 8 bits have no packed form
  } else if (!P.deq && !P.mask && !P.e_in && !P.blk_cnt) {
    return MSQ_ERR_BAD_ARG;
  }
  size_t smem = (packing && outlier && P.hi_w > 0) ? (size_t)P.hi_seg_stride : 0;
  const bool inout_al = aligned16(P.in) && (!P.deq || aligned16(P.deq));
  const bool small = P.pre * (long long)P.L * P.post < 0x7fffffffLL;     // fast paths use 32-bit offsets
  const bool fast_ok = small && !P.mask && (packing || P.deq);
  int path = 0;
  cudaError_t err;
  // 2-D weight blocked along its rows, dequantized values only (what MXQuantizer.quantize asks for): persistent
  // TMA-staged kernel (msq_quant_tma.cu); path 4
  if (outlier && fast_ok && !packing && P.pre == 1 && P.post > 1 && (P.post & 3) == 0 && inout_al && P.c.variant == 0 &&
      !P.e_in && !P.e_out && !P.blk_cnt && (P.bs == 8 || P.bs == 16 || P.bs == 32)) {
    const char* e = getenv("MSQ_QUANT_TMA");
    if (!(e && e[0] == '0')) {
      const float* in1[1] = {P.in};
      float* out1[1] = {P.deq};
      const long long rows1[1] = {P.L}, cols1[1] = {P.post};
      const int rc = launch_strided_tma(1, in1, out1, rows1, cols1, P.bs, P.c, P.status, stream);
      if (rc != MSQ_ERR_UNSUPPORTED) {
        if (path_out) *path_out = 4;
        return rc;
      }
    }
  }
  // quantize + pack of a 2-D weight (16-row blocks, 4-bit slots): TMA-staged segment kernel; path 5
  if (outlier && packing && fast_ok && P.pre == 1 && P.post > 1) {
    const char* e = getenv("MSQ_QUANT_TMA");
    if (!(e && e[0] == '0')) {
      const int rc = launch_pack_tma(P, stream);
      if (rc != MSQ_ERR_UNSUPPORTED) {
        if (path_out) *path_out = 5;
        return rc;
      }
    }
  }
  if (fast_ok && P.post == 1 && P.L % P.bs == 0 && inout_al && !P.ext_max &&
      (P.bs == 8 || P.bs == 16 || P.bs == 32 || P.bs == 64 || P.bs == 128)) {
    path = 2;
    err = launch_inner(P, outlier, smem, stream);
  } else if (fast_ok && P.post > 1 && (P.bs == 8 || P.bs == 16 || P.bs == 32) && (!packing || P.post % 32 == 0) &&
             (packing || P.pre * (long long)P.nblk <= 65535)) {
    path = 1;
    err = launch_strided(P, outlier, smem, stream);
  } else if (fast_ok && !packing && P.post > 1 && (P.post & 3) == 0 && inout_al && (P.bs == 64 || P.bs == 128) &&
             !(outlier && P.c.variant != 0) && P.pre * (long long)P.nblk * ((P.post + 63) / 64) < 0x7fffffffLL) {
    path = 6;
    err = launch_strided_big(P, outlier, stream);
  } else {
    if (packing) return MSQ_ERR_UNSUPPORTED;
    path = 3;
    err = launch_generic(P, outlier, stream);
  }
  if (path_out) *path_out = path;
  return err == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // namespace msq

using namespace msq;

static int fill_geometry(QParams& P, long long pre, long long L, long long post, int bs) {
  if (pre <= 0 || L <= 0 || post <= 0 || bs <= 0 || L > 0x7fffffffLL) return MSQ_ERR_BAD_ARG;
  P.pre = pre; P.L = (int)L; P.post = post; P.bs = bs;
  P.nblk = (int)((L + bs - 1) / bs);
  P.nblocks = pre * P.nblk * post;
  P.post_m = P.nblk_m = 0; P.post_k = P.nblk_k = 0;
  if (post < 0x7fffffffLL) magic_div(post, &P.post_m, &P.post_k);
  magic_div(P.nblk, &P.nblk_m, &P.nblk_k);
  return MSQ_OK;
}

extern "C" {

int msq_packed_hi_width(int in_ebits, int in_mbits, int out_ebits, int out_mbits) {
  Fmt fi = make_fmt(in_ebits, in_mbits, 0, 1.0f), fo = make_fmt(out_ebits, out_mbits, 0, 1.0f);
  int d = fo.width - fi.width;
  return d <= 0 ? 0 : (d <= 2 ? 2 : (d <= 4 ? 4 : 8));
}

int msq_packed_code_width(int ebits, int mbits) { return make_fmt(ebits, mbits, 0, 1.0f).width; }

long long msq_packed_hi_seg_stride(int block_size, int hi_w) {
  long long b = ((long long)MSQ_SEG_BLOCKS * block_size * hi_w) >> 3;
  return (b + 15) & ~15LL;
}

int msq_seg_blocks(void) { return MSQ_SEG_BLOCKS; }

// (m, k) with floor(n / d) == (n * m) >> k for every 0 <= n < 2^31: the divider the packing kernels use for their
// block-linear index arithmetic (exported so that the host-side construction can be tested without a GPU)
int msq_magic_div(long long d, unsigned int* m, int* k) {
  if (d <= 0 || d >= 0x7fffffffLL || !m || !k) return MSQ_ERR_BAD_ARG;
  magic_div(d, m, k);
  return MSQ_OK;
}

int msq_quantize_mx_outlier_f32(
    const float* in, float* deq, long long pre, long long L, long long post, int block_size,
    int in_scale_bits, int out_scale_bits,
    int in_ebits, int in_mbits, int in_emax, float in_max_norm,
    int out_ebits, int out_mbits, int out_emax, float out_max_norm,
    float std_dev, int flush_fp32_subnorms, int rmode, int variant,
    const float* lo_ext, const float* hi_ext,
    unsigned char* codes, signed char* e_in, signed char* e_out, unsigned int* bitmap,
    unsigned char* hi, int* seg_cnt, unsigned char* blk_cnt, unsigned char* mask,
    int* status, int* path, void* stream) {
  msq::DevGuard dev_guard_(in);
  if (!in || in_scale_bits <= 0 || out_scale_bits <= 0 || in_scale_bits > 8 || out_scale_bits > 8) return MSQ_ERR_BAD_ARG;
  if (rmode < 0 || rmode > 2 || (variant != 0 && variant != 1)) return MSQ_ERR_BAD_ARG;
  if (variant == 1 && (!lo_ext || !hi_ext)) return MSQ_ERR_BAD_ARG;
  QParams P;
  memset(&P, 0, sizeof(P));
  int rc = fill_geometry(P, pre, L, post, block_size);
  if (rc) return rc;
  P.c.fin = make_fmt(in_ebits, in_mbits, in_emax, in_max_norm);
  P.c.fout = make_fmt(out_ebits, out_mbits, out_emax, out_max_norm);
  P.c.sb_in = in_scale_bits; P.c.sb_out = out_scale_bits;
  P.c.std_dev = std_dev; P.c.flush = flush_fp32_subnorms; P.c.variant = variant; P.c.rmode = rmode; P.c.mode = 0;
  finish_cfg(P.c);
  P.in = in; P.deq = deq; P.codes = codes; P.e_in = e_in; P.e_out = e_out; P.bitmap = bitmap; P.hi = hi;
  P.seg_cnt = seg_cnt; P.blk_cnt = blk_cnt; P.mask = mask; P.status = status; P.lo_ext = lo_ext; P.hi_ext = hi_ext;
  P.hi_w = msq_packed_hi_width(in_ebits, in_mbits, out_ebits, out_mbits);
  P.hi_seg_stride = msq_packed_hi_seg_stride(block_size, P.hi_w);
  if (hi && P.hi_w == 0) P.hi = nullptr;
  return launch_quant(P, true, (cudaStream_t)stream, path);
}

// quantize_mx_outlier_v1 (utils/quant.py:147-266, axes=[0]) for up to 16 weight matrices [rows[i], cols[i]] in ONE
// launch -- what the reference's RTN driver does linear by linear for a decoder layer (llm/llama.py:238-253).
// in / out: HOST arrays of n DEVICE pointers (fp32, contiguous, 16-byte aligned); results are bit-identical to n calls
// of msq_quantize_mx_outlier_f32(pre = 1, L = rows, post = cols).  block_size 8, 16 or 32; each matrix < 2^31 elements.
int msq_quantize_mx_outlier_multi_f32(int n, const float* const* in, float* const* out, const long long* rows,
                                      const long long* cols, int block_size, int in_scale_bits, int out_scale_bits,
                                      int in_ebits, int in_mbits, int in_emax, float in_max_norm, int out_ebits,
                                      int out_mbits, int out_emax, float out_max_norm, float std_dev,
                                      int flush_fp32_subnorms, int rmode, int* status, void* stream) {
  if (n <= 0 || n > MSQ_MULTI_MAX || !in || !out || !rows || !cols) return MSQ_ERR_BAD_ARG;
  msq::DevGuard dev_guard_(in[0]);
  if (in_scale_bits <= 0 || out_scale_bits <= 0 || in_scale_bits > 8 || out_scale_bits > 8 || rmode < 0 || rmode > 2)
    return MSQ_ERR_BAD_ARG;
  if (block_size != 8 && block_size != 16 && block_size != 32) return MSQ_ERR_UNSUPPORTED;
  MultiQParams MP;
  memset(&MP, 0, sizeof(MP));
  MP.n = n; MP.status = status;
  MP.c.fin = make_fmt(in_ebits, in_mbits, in_emax, in_max_norm);
  MP.c.fout = make_fmt(out_ebits, out_mbits, out_emax, out_max_norm);
  MP.c.sb_in = in_scale_bits; MP.c.sb_out = out_scale_bits;
  MP.c.std_dev = std_dev; MP.c.flush = flush_fp32_subnorms; MP.c.variant = 0; MP.c.rmode = rmode; MP.c.mode = 0;
  finish_cfg(MP.c);
  {
    // persistent TMA-staged kernel when every matrix can be addressed by a tensor map (cols % 4 == 0);
    // MSQ_QUANT_TMA=0 keeps the direct-load kernel (A/B measurements)
    const char* e = getenv("MSQ_QUANT_TMA");
    if (!(e && e[0] == '0')) {
      bool ok = true;
      for (int i = 0; i < n && ok; ++i) ok = in[i] && out[i] && rows[i] > 0 && cols[i] > 0 && rows[i] * cols[i] < 0x7fffffffLL;
      if (!ok) return MSQ_ERR_BAD_ARG;
      const int rc = launch_strided_tma(n, in, out, rows, cols, block_size, MP.c, status, (cudaStream_t)stream);
      if (rc != MSQ_ERR_UNSUPPORTED) return rc;
    }
  }
  unsigned long long total = 0;
  for (int i = 0; i < n; ++i) {
    if (!in[i] || !out[i] || rows[i] <= 0 || cols[i] <= 0 || rows[i] * cols[i] >= 0x7fffffffLL) return MSQ_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(in[i]) & 15) || (reinterpret_cast<uintptr_t>(out[i]) & 15)) return MSQ_ERR_ALIGN;
    MP.in[i] = in[i]; MP.deq[i] = out[i];
    MP.L[i] = (int)rows[i]; MP.post[i] = (int)cols[i];
    MP.gx[i] = (int)((cols[i] + 255) / 256);
    MP.cta_start[i] = (unsigned)total;
    total += (unsigned long long)MP.gx[i] * (unsigned long long)((rows[i] + block_size - 1) / block_size);
    if (total >= 0x7fffffffULL) return MSQ_ERR_UNSUPPORTED;
  }
  MP.cta_start[n] = (unsigned)total;
  return launch_strided_multi(MP, block_size, (unsigned)total, (cudaStream_t)stream) == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_quantize_mx_f32(
    const float* in, float* deq, long long pre, long long L, long long post, int block_size,
    int scale_bits, int ebits, int mbits, int emax, float max_norm,
    int flush_fp32_subnorms, int rmode, int mode, const float* ext_max,
    unsigned char* codes, signed char* e_shared, int* status, int* path, void* stream) {
  msq::DevGuard dev_guard_(in);
  if (!in || scale_bits <= 0 || scale_bits > 8 || rmode < 0 || rmode > 2 || (mode != 0 && mode != 1)) return MSQ_ERR_BAD_ARG;
  QParams P;
  memset(&P, 0, sizeof(P));
  int rc = fill_geometry(P, pre, L, post, block_size);
  if (rc) return rc;
  P.c.fin = make_fmt(ebits, mbits, emax, max_norm);
  P.c.fout = P.c.fin;
  P.c.sb_in = P.c.sb_out = scale_bits;
  P.c.flush = flush_fp32_subnorms; P.c.rmode = rmode; P.c.mode = mode;
  finish_cfg(P.c);
  P.in = in; P.deq = deq; P.codes = codes; P.e_in = e_shared; P.status = status; P.ext_max = ext_max;
  return launch_quant(P, false, (cudaStream_t)stream, path);
}

int msq_outlier_bounds_across_blocks_f32(const float* in, float* lo, float* hi, long long pre, long long L,
                                         long long post, int block_size, float std_dev, void* stream) {
  msq::DevGuard dev_guard_(in);
  if (!in || !lo || !hi || block_size <= 0) return MSQ_ERR_BAD_ARG;
  const int nblk = (int)((L + block_size - 1) / block_size);
  if (nblk < 2) return MSQ_ERR_BAD_ARG;    // reference asserts: std of a single sample is NaN (mx_ops.py:66)
  const long long n = pre * block_size * post;
  k_bounds_across_blocks<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, lo, hi, pre, (int)L, post,
                                                                                         block_size, nblk, std_dev);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

static int elemwise_grid(long long n, int vec) {
  long long b = (n + 256LL * vec - 1) / (256LL * vec);
  long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

int msq_quantize_elemwise(const void* in, void* out, long long n, int dtype, int bits, int exp_bits, float max_norm,
                          int rmode, int saturate_normals, int allow_denorm, int native, void* stream) {
  msq::DevGuard dev_guard_(in);
  if (!in || !out || n < 0 || bits < 2 || bits > 25 || rmode < 0 || rmode > 2) return MSQ_ERR_BAD_ARG;
  if (n == 0) return MSQ_OK;
  Fmt f = make_fmt(exp_bits, bits, 0, max_norm);
  cudaStream_t s = (cudaStream_t)stream;
  const bool al = aligned16(in) && aligned16(out);
  if (dtype == 0) {
    if (al) k_elemwise<float, 4><<<elemwise_grid(n, 4), 256, 0, s>>>((const float*)in, (float*)out, n, f, rmode, saturate_normals, allow_denorm, native);
    else k_elemwise<float, 1><<<elemwise_grid(n, 1), 256, 0, s>>>((const float*)in, (float*)out, n, f, rmode, saturate_normals, allow_denorm, native);
  } else if (dtype == 1) {
    if (al) k_elemwise<__half, 8><<<elemwise_grid(n, 8), 256, 0, s>>>((const __half*)in, (__half*)out, n, f, rmode, saturate_normals, allow_denorm, native);
    else k_elemwise<__half, 1><<<elemwise_grid(n, 1), 256, 0, s>>>((const __half*)in, (__half*)out, n, f, rmode, saturate_normals, allow_denorm, native);
  } else if (dtype == 2) {
    if (al) k_elemwise<__nv_bfloat16, 8><<<elemwise_grid(n, 8), 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, f, rmode, saturate_normals, allow_denorm, native);
    else k_elemwise<__nv_bfloat16, 1><<<elemwise_grid(n, 1), 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, f, rmode, saturate_normals, allow_denorm, native);
  } else {
    return MSQ_ERR_BAD_ARG;
  }
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"
