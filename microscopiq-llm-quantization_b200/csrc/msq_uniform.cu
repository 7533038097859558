// msq_uniform.cu -- GPTQ uniform affine quantizer (utils/quant.py:268-390) and the
// shared-exponent helper (utils/quant.py:498-541).
#include "msq_common.cuh"
#include "msq_api_internal.h"

namespace msq {

// quantize(x, scale, zero, maxq)  utils/quant.py:268-272.  x is [pre, C, post]; scale/zero have C entries.
__global__ void __launch_bounds__(256) k_uniform_quantize(const float* __restrict__ x, float* __restrict__ out,
                                                          const float* __restrict__ scale, const float* __restrict__ zero,
                                                          int maxq, long long Cn, long long post, long long n) {
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    const long long c = (i / post) % Cn;
    const float s = __ldg(scale + c), z = __ldg(zero + c), v = __ldg(x + i);
    float r;
    if (maxq < 0) {
      const float gt = v > __fmul_rn(s, 0.5f) ? 1.0f : 0.0f;
      const float lt = v < __fmul_rn(z, 0.5f) ? 1.0f : 0.0f;
      r = __fadd_rn(__fmul_rn(gt, s), __fmul_rn(lt, z));
    } else {
      float q = __fadd_rn(rintf(__fdiv_rn(v, s)), z);
      q = q < 0.0f ? 0.0f : q;
      q = q > (float)maxq ? (float)maxq : q;
      r = __fmul_rn(s, __fsub_rn(q, z));
    }
    out[i] = r;
  }
}

__device__ __forceinline__ float blk_min(float v, float* sm) {
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sm[0];
  for (int i = 1; i < 8; ++i) r = fminf(r, sm[i]);
  return r;
}
__device__ __forceinline__ float blk_max(float v, float* sm) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sm[0];
  for (int i = 1; i < 8; ++i) r = fmaxf(r, sm[i]);
  return r;
}
__device__ __forceinline__ double blk_sum(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < 8; ++i) r += sm[i];
  return r;
}

// Quantizer.find_params  utils/quant.py:298-357, one CTA per row of x [R, n].
__global__ void __launch_bounds__(256) k_uniform_find_params(const float* __restrict__ x, float* __restrict__ scale,
                                                             float* __restrict__ zero, long long n, int maxq, int sym,
                                                             int mse, float norm, int grid, int ncand) {
  __shared__ float sm[8];
  __shared__ double smd[8];
  const float* row = x + (long long)blockIdx.x * n;
  float mn = 0.0f, mx = 0.0f;                 // torch.minimum(x.min, 0) / torch.maximum(x.max, 0)
  int nan = 0;
  for (long long i = threadIdx.x; i < n; i += 256) { float v = __ldg(row + i); nan |= (v != v); mn = fminf(mn, v); mx = fmaxf(mx, v); }
  float xmin = blk_min(mn, sm), xmax = blk_max(mx, sm);
  if (__syncthreads_or(nan)) { xmin = qnan(); xmax = qnan(); }
  if (sym) {
    xmax = fmaxf(fabsf(xmin), xmax);
    if (xmin < 0.0f) xmin = -xmax;
  }
  if (xmin == 0.0f && xmax == 0.0f) { xmin = -1.0f; xmax = 1.0f; }
  float s, z;
  if (maxq < 0) { s = xmax; z = xmin; }
  else {
    s = __fdiv_rn(__fsub_rn(xmax, xmin), (float)maxq);
    z = sym ? (float)((maxq + 1) / 2.0) : rintf(__fdiv_rn(-xmin, s));
  }
  if (mse && maxq >= 0) {
    float best = __int_as_float(0x7F800000);
    for (int c = 0; c < ncand; ++c) {
      const float p = (float)(1.0 - (double)c / (double)grid);
      const float xmin1 = __fmul_rn(p, xmin), xmax1 = __fmul_rn(p, xmax);
      const float s1 = __fdiv_rn(__fsub_rn(xmax1, xmin1), (float)maxq);
      const float z1 = sym ? z : rintf(__fdiv_rn(-xmin1, s1));
      double acc = 0.0;
      for (long long i = threadIdx.x; i < n; i += 256) {
        const float v = __ldg(row + i);
        float q = __fadd_rn(rintf(__fdiv_rn(v, s1)), z1);
        q = q < 0.0f ? 0.0f : q;
        q = q > (float)maxq ? (float)maxq : q;
        const float d = fabsf(__fsub_rn(__fmul_rn(s1, __fsub_rn(q, z1)), v));
        acc += (double)powf(d, norm);
      }
      const float err = (float)blk_sum(acc, smd);
      if (err < best) { best = err; s = s1; z = z1; }
    }
  }
  if (threadIdx.x == 0) { scale[blockIdx.x] = s; zero[blockIdx.x] = z; }
}

// _shared_exponents(method="max")  utils/quant.py:512-541: x [pre, L, post] -> out [pre, post]
__global__ void __launch_bounds__(256) k_shared_exponents(const float* __restrict__ x, float* __restrict__ out,
                                                          long long pre, long long L, long long post, int ebits) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= pre * post) return;
  const long long p = t / post, j = t - p * post;
  float m = 0.0f;
  int nan = 0;
  for (long long l = 0; l < L; ++l) { float a = fabsf(__ldg(x + (p * L + l) * post + j)); nan |= (a != a); m = fmaxf(m, a); }
  float e;
  if (nan) e = qnan();
  else if (m == __int_as_float(0x7F800000)) e = m;
  else e = (float)((m == 0.0f) ? -126 : floor_log2f(m));
  if (ebits > 0 && e == e) {
    const float emax = (float)((1 << (ebits - 1)) - 1);
    if (e > emax) e = qnan();
    else if (e < -emax) e = -emax;
  }
  out[t] = e;
}

}  // namespace msq

using namespace msq;

extern "C" {

int msq_uniform_quantize_f32(const float* x, float* out, const float* scale, const float* zero, int maxq,
                             long long pre, long long Cn, long long post, void* stream) {
  msq::DevGuard dev_guard_(x);
  if (!x || !out || !scale || !zero || pre <= 0 || Cn <= 0 || post <= 0) return MSQ_ERR_BAD_ARG;
  const long long n = pre * Cn * post;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  k_uniform_quantize<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, out, scale, zero, maxq, Cn, post, n);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_uniform_find_params_f32(const float* x, float* scale, float* zero, long long rows, long long n, int maxq,
                                int sym, int mse, float norm, int grid, float maxshrink, void* stream) {
  msq::DevGuard dev_guard_(x);
  if (!x || !scale || !zero || rows <= 0 || n <= 0 || rows > 0x7fffffffLL) return MSQ_ERR_BAD_ARG;
  const int ncand = (int)(maxshrink * grid);             // range(int(maxshrink * grid)), quant.py:342
  k_uniform_find_params<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, scale, zero, n, maxq, sym, mse, norm, grid, ncand);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_shared_exponents_f32(const float* x, float* out, long long pre, long long L, long long post, int ebits,
                             void* stream) {
  msq::DevGuard dev_guard_(x);
  if (!x || !out || pre <= 0 || L <= 0 || post <= 0) return MSQ_ERR_BAD_ARG;
  const long long n = pre * post;
  k_shared_exponents<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, out, pre, L, post, ebits);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"
