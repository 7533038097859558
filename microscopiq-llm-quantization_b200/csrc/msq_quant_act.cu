// msq_quant_act.cu -- the ACTIVATION side of the block-scaled linears: plain MX quantization of [rows, K] activations
// along K (blocks of 32, one shared power-of-two exponent each) straight to the operands of the tcgen05 block-scaled GEMMs:
// packed FP4 (E2M1) or FP8 (E4M3) codes + int8 shared exponents.  Same arithmetic as msq_quantize_mx_f32(mode = 1)
// (cpp/mx.cuh:37-52 + cpp/shared_exp.cuh:14-53 + cpp/quantize.cuh:88-149 of the reference, restated in core_plain) and
// bit-identical planes (tests/test_gpu_mxops.py), but written for the place it runs: in front of EVERY linear call.
//
//  * fp32, bf16 or fp16 input is read as it is (the generic entry point needs an fp32 copy first: two more passes over
//    the activations); a thread owns a whole block of 32 values (no shuffles, the block-level work once per 32 elements,
//    16- / 32-byte code stores, coalesced exponent bytes);
//  * element rounding + encoding by cvt.rn.satfinite.e2m1x2.f32 / e4m3x2.f32.  The reference rounds the exact value
//    half AWAY from zero on the integer mantissa (quantize.cuh:120-131) and saturates to max_norm; the hardware rounds
//    half to EVEN and saturates to the same max_norm.  Setting the last mantissa bit of x before the exact power-of-two
//    scaling moves every tie just above itself and nothing across a tie (ties of the 4- and 8-bit grids have that bit
//    clear), so the two agree on every input; a zero result carries no sign (quantize.cuh returns +0).
//  * one launch, no workspace: measured against the generic path 4096 x 4096: fp32 -> FP4 38.8 us, bf16 -> FP4 83.9 us.
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_quant_kernels.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace msq {
namespace {

struct ActParams {
  const void* x; long long ldx;          // elements
  unsigned char* codes; long long ldc;   // bytes per row of codes
  signed char* e; long long lde;         // exponents per row
  int* status;
  long long rows; int K;
  int tasks, cpr; unsigned int cpr_m; int cpr_k;      // blocks, blocks per row and its magic divider
  int S;                                 // 2^(scale_bits - 1) - 1
  int flush;
};

// 32 values of one block from 128 (fp32) or 64 (16-bit) contiguous bytes
// 32 bytes per load instruction (one whole sector per lane: with 16-byte loads at a 128-byte lane stride every sector is
// requested twice and the fp32 path ran at 0.6-0.7 of the copy roofline)
__device__ __forceinline__ void ldg32(const char* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
// 32 values of one block from 128 (fp32) or 64 (16-bit) contiguous bytes
template <int DT>
__device__ __forceinline__ void load_block(const char* p, float (&v)[32]) {
  if (DT == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t r[8];
      ldg32(p + 32 * j, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * j + i] = __uint_as_float(r[i]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t w[8];
      ldg32(p + 32 * j, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (DT == 1) {
          v[16 * j + 2 * i] = __uint_as_float(w[i] << 16);
          v[16 * j + 2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        } else {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
          v[16 * j + 2 * i] = f.x; v[16 * j + 2 * i + 1] = f.y;
        }
      }
    }
  }
}

// eight fp32 -> eight E2M1 codes (element j in nibble j) / four fp32 -> four E4M3 codes (element j in byte j)
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

// thread = one block of 32 values: no shuffles, the block-level work once per 32 elements, 16- / 32-byte code stores and
// one exponent byte per lane (consecutive lanes = consecutive blocks of a row).  A lane's loads are 16 bytes at a 128- /
// 64-byte stride across the warp; the other half of every sector is the next load of the same lane (an L1 hit)
template <int DT, bool FP4>
__global__ void __launch_bounds__(256) k_quant_act(const ActParams P) {
  constexpr int EMAX = FP4 ? 2 : 8;
  int st = 0;
  const int stride = (int)(gridDim.x * blockDim.x);
  for (int b = (int)(blockIdx.x * blockDim.x + threadIdx.x); b < P.tasks; b += stride) {
    const int row = fast_div(b, P.cpr_m, P.cpr_k);
    const int kb = b - row * P.cpr;
    float v[32];
    load_block<DT>(reinterpret_cast<const char*>(P.x) + ((long long)row * P.ldx + (long long)kb * 32) * (DT == 0 ? 4 : 2), v);
    // exponent field of the block maximum: maximum of the magnitudes' bit patterns (an Inf or NaN gives 255)
    uint32_t mb = 0;
#pragma unroll
    for (int i = 0; i < 32; i += 2)
      mb = max(mb, max((uint32_t)__float_as_int(v[i]) & 0x7FFFFFFFu, (uint32_t)__float_as_int(v[i + 1]) & 0x7FFFFFFFu));
    // shared exponent exactly as core_plain mode 1 (shared_exp.cuh:14-53)
    const int be = (int)(mb >> 23);
    const bool bad = be == 255;
    const bool flush = (be == 0) && P.flush;
    int sh = bad ? 255 : be - EMAX;
    const int ub = sh - 127;
    sh = ub > P.S ? 255 : sh;
    sh = ub < -P.S ? 127 - P.S : sh;              // >= 0 for S <= 127
    const bool dead = sh == 255;
    if (dead) st |= bad ? MSQ_ST_NONFINITE_INPUT : MSQ_ST_EXP_IN_OVERFLOW;
    const float inv = (flush || dead) ? 0.0f : pow2i(127 - sh);
    // ties away from zero (quantize.cuh:120-131) = ties to even on x | 1; the power-of-two scaling is exact
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __fmul_rn(__int_as_float(__float_as_int(v[i]) | 1), inv);
    unsigned char* dst = P.codes + (long long)row * P.ldc + (long long)kb * (FP4 ? 16 : 32);
    if (FP4) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t a = cvt8_e2m1(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3], v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]);
        a &= (((a & 0x77777777u) + 0x77777777u) & 0x88888888u) | 0x77777777u;        // a zero magnitude carries no sign
        w[j] = dead ? 0u : a;
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t a = cvt4_e4m3(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        a &= (((a & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) & 0x80808080u) | 0x7F7F7F7Fu;
        w[j] = dead ? 0u : a;
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(dst + 16) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    P.e[(long long)row * P.lde + kb] = (signed char)(dead ? MSQ_EXP_NAN : sh - 127);
  }
  if (P.status) {
    st = __reduce_or_sync(0xffffffffu, st);
    if ((threadIdx.x & 31) == 0 && st) atomicOr(P.status, st);
  }
}

template <int DT>
static int launch(ActParams P, bool fp4, cudaStream_t s) {
  P.cpr = P.K / 32;                                     // blocks per row
  const long long tasks = P.rows * P.cpr;
  if (tasks >= 0x7fffffffLL) return MSQ_ERR_UNSUPPORTED;
  P.tasks = (int)tasks;
  magic_div(P.cpr, &P.cpr_m, &P.cpr_k);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (tasks + 255) / 256;
  // exactly one wave of resident CTAs once the tensor is large (a fixed 8 per SM left a partial second wave)
  void (*k)(const ActParams) = fp4 ? k_quant_act<DT, true> : k_quant_act<DT, false>;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 4; }
  const long long cap = (long long)sms * occ;
  if (blocks > cap) blocks = cap;
  k<<<(unsigned)blocks, 256, 0, s>>>(P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" int msq_quantize_mx_act(const void* x, int x_dtype, long long ldx, long long rows, int K, int elem_is_fp4, int scale_bits,
                                   int flush_fp32_subnorms, unsigned char* codes, long long ldc_bytes, signed char* e_shared,
                                   long long lde, int* status, void* stream) {
  msq::DevGuard dev_guard_(x);
  if (!x || !codes || !e_shared || rows <= 0 || K <= 0 || x_dtype < 0 || x_dtype > 2) return MSQ_ERR_BAD_ARG;
  if (K % 32 != 0) return MSQ_ERR_UNSUPPORTED;
  if (scale_bits < 1 || scale_bits > 8) return MSQ_ERR_UNSUPPORTED;
  const int esz = x_dtype == 0 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(x) & 31) || ((ldx * esz) & 31) || (reinterpret_cast<uintptr_t>(codes) & 15) ||
      (ldc_bytes & 15))
    return MSQ_ERR_ALIGN;
  if (ldx < K || ldc_bytes < (elem_is_fp4 ? K / 2 : K) || lde < K / 32) return MSQ_ERR_BAD_ARG;
  ActParams P;
  P.x = x; P.ldx = ldx; P.codes = codes; P.ldc = ldc_bytes; P.e = e_shared; P.lde = lde; P.status = status;
  P.rows = rows; P.K = K; P.S = (1 << (scale_bits - 1)) - 1; P.flush = flush_fp32_subnorms ? 1 : 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == 0) return launch<0>(P, elem_is_fp4 != 0, s);
  if (x_dtype == 1) return launch<1>(P, elem_is_fp4 != 0, s);
  return launch<2>(P, elem_is_fp4 != 0, s);
}
