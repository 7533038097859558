// msq_split_a16.cu -- 16-bit activations as a sum of block-scaled E4M3 terms, so that a W4A16 linear can run on the
// block-scaled tensor-core path (tcgen05.mma kind::mxf8f6f4, the packed 4-bit weight codes as the other operand) instead of
// dequantizing the weight to bf16 in producer warps:
//     x[m, k] = sum_{t < T} a_t[m, k] * 2^(e_t[m, k / 32]),   a_t in E4M3 (4 significant bits each), e_t = e_1 - 3 (t - 1), T = 2 or 3
// Per block of 32 along K: e_1 = (exponent of the block's largest magnitude) - 7; a_1 = x * 2^-e_1 truncated onto the
// E4M3 grid, a_2 = the remainder * 8 truncated, a_3 likewise -- every step exact in fp32.  The three terms hold 12
// significant bits in a window that reaches 22 binades below the block maximum: a bf16 element (8 bits) is reproduced
// EXACTLY when it is within 2^15 of its block's maximum, an fp16 element (11 bits) within 2^11; smaller elements are
// truncated at 2^-22 of the block maximum (below the fp32 accumulation error of the contraction itself).
// With T = 2 (8 significant bits: enough for bf16) the window is 2^12 / truncation at 2^-19 of the block maximum.
// The terms are stacked along the token dimension ([T M, K] codes, [T M, K/32] exponents): one GEMM with T M "tokens"
// reads the weight once; the caller adds the T M-row slabs of the result.
#include "msq_api_internal.h"
#include "msq_common.cuh"

namespace msq {
namespace {

// v >= 0, v < 2^8: largest E4M3 value <= v (code without sign) and that value
__device__ __forceinline__ uint32_t trunc_e4m3(float v, float& a) {
  const uint32_t u = __float_as_uint(v);
  if (u >= 0x3C800000u) {                       // >= 2^-6: normal, keep three mantissa bits
    a = __uint_as_float(u & 0xFFF00000u);
    return (u >> 20) - 960u;                    // ((E + 7) << 3) | m3
  }
  const uint32_t c = (uint32_t)__float2int_rz(__fmul_rn(v, 512.0f));      // subnormals: multiples of 2^-9
  a = __fmul_rn((float)c, 1.0f / 512.0f);
  return c;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// thread = 8 consecutive elements, 4 lanes = one block of 32
template <typename T, int NT>
__global__ void __launch_bounds__(256) k_split_a16(const T* __restrict__ x, long long ldx, int M, int K, uint8_t* __restrict__ codes,
                                                   int8_t* __restrict__ e_out) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const int per_row = K >> 3;
  const long long total = (long long)M * per_row;
  const bool active = t < total;
  const int m = active ? (int)(t / per_row) : 0;
  const int c8 = active ? (int)(t - (long long)m * per_row) : 0;
  float v[8];
  if (active) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + (long long)m * ldx + c8 * 8));
    const T* p = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = to_f32<T>(p[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
  }
  float amax = 0.0f;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) { amax = fmaxf(amax, fabsf(v[i])); bad |= !(fabsf(v[i]) <= 3.402823466e38f); }
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  bad = __shfl_xor_sync(0xffffffffu, (int)bad, 1) | (int)bad;
  bad = __shfl_xor_sync(0xffffffffu, (int)bad, 2) | (int)bad;
  // e_1 = E(max) - 7, kept inside the UE8M0 range of all three terms (an all-zero block gets the smallest scale)
  int e1 = (int)((__float_as_uint(amax) >> 23) & 0xFFu) - 127 - 7;
  e1 = max(e1, -127 + 6);
  const float inv = pow2i(-e1);
  uint32_t c[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) c[j][0] = c[j][1] = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t sign = (__float_as_uint(v[i]) >> 24) & 0x80u;
    float r = __fmul_rn(fabsf(v[i]), inv), a;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint32_t code = trunc_e4m3(r, a) | sign;
      c[j][i >> 2] |= code << (8 * (i & 3));
      r = __fmul_rn(__fsub_rn(r, a), 8.0f);
    }
  }
  if (!active) return;
  const int kb32 = K >> 5;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const long long row = (long long)j * M + m;
    *reinterpret_cast<uint2*>(codes + row * K + c8 * 8) = make_uint2(c[j][0], c[j][1]);
    if ((c8 & 3) == 0) e_out[row * kb32 + (c8 >> 2)] = bad ? (int8_t)MSQ_EXP_NAN : (int8_t)(e1 - 3 * j);
  }
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" {

// x [M, K] bf16 (dtype 2) or fp16 (dtype 1), row pitch ldx elements -> codes [terms * M, K] E4M3 bytes, e [terms * M, K/32]
// int8 (term t of token m in row t * M + m); terms = 2 or 3; K % 32 == 0, x 16-byte aligned, ldx % 8 == 0.
// Two terms hold 8 significant bits: bf16 exactly within 2^12 of the block maximum, truncation at 2^-19 of it below; three
// terms 12 bits: bf16 exactly within 2^15, fp16 within 2^11, truncation at 2^-22.  A block that holds an Inf / NaN gets the
// NaN scale (MSQ_EXP_NAN) in every term.
int msq_split_a16_e4m3(const void* x, long long ldx, int dtype, int M, int K, int terms, unsigned char* codes, signed char* e,
                       void* stream) {
  msq::DevGuard dev_guard_(x);
  if (!x || !codes || !e || M <= 0 || K <= 0 || (terms != 2 && terms != 3)) return MSQ_ERR_BAD_ARG;
  if (dtype != 1 && dtype != 2) return MSQ_ERR_UNSUPPORTED;
  if (K % 32 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (ldx % 8) || (reinterpret_cast<uintptr_t>(codes) & 7)) return MSQ_ERR_ALIGN;
  const long long total = (long long)M * (K >> 3);
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  int8_t* e8 = reinterpret_cast<int8_t*>(e);
  if (dtype == 2) {
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
    if (terms == 2) k_split_a16<__nv_bfloat16, 2><<<grid, 256, 0, s>>>(xb, ldx, M, K, codes, e8);
    else k_split_a16<__nv_bfloat16, 3><<<grid, 256, 0, s>>>(xb, ldx, M, K, codes, e8);
  } else {
    const __half* xh = reinterpret_cast<const __half*>(x);
    if (terms == 2) k_split_a16<__half, 2><<<grid, 256, 0, s>>>(xh, ldx, M, K, codes, e8);
    else k_split_a16<__half, 3><<<grid, 256, 0, s>>>(xh, ldx, M, K, codes, e8);
  }
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"
