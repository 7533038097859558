// msq_quant_act.cu -- the ACTIVATION side of the block-scaled linears: plain MX quantization of [rows, K] activations
// along K (blocks of 32, one shared power-of-two exponent each) straight to the operands of the tcgen05 block-scaled GEMMs:
// packed FP4 (E2M1) or FP8 (E4M3) codes + int8 shared exponents.  Same arithmetic as msq_quantize_mx_f32(mode = 1)
// (cpp/mx.cuh:37-52 + cpp/shared_exp.cuh:14-53 + cpp/quantize.cuh:88-149 of the reference, restated in core_plain) and
// bit-identical planes (tests/test_gpu_mxops.py), but written for the place it runs: in front of EVERY linear call.
//
//  * fp32, bf16 or fp16 input is read as it is (the generic entry point needs an fp32 copy first: two more passes over
//    the activations), 16 bytes per lane, a block of 32 values = 8 lanes (fp32) or 4 lanes (16-bit) of one warp: every
//    load instruction of a warp covers 512 contiguous bytes, the block maximum is 3 (2) shuffles;
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
  int tasks, cpr; unsigned int cpr_m; int cpr_k;      // warp tasks, tasks per row and its magic divider
  int S;                                 // 2^(scale_bits - 1) - 1
  int flush;
};

template <int DT> struct Elt;            // 0 = fp32, 1 = bf16, 2 = fp16
template <> struct Elt<0> { static constexpr int EPL = 4; };
template <> struct Elt<1> { static constexpr int EPL = 8; };
template <> struct Elt<2> { static constexpr int EPL = 8; };

template <int DT>
__device__ __forceinline__ void load16(const void* p, float (&v)[Elt<DT>::EPL]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  if (DT == 0) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  } else {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (DT == 1) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      }
    }
  }
}

__device__ __forceinline__ uint32_t cvt_e2m1x2(float lo, float hi) {      // byte: lo in the low nibble
  uint32_t d;
  asm("{\n.reg .b8 b;\ncvt.rn.satfinite.e2m1x2.f32 b, %2, %1;\ncvt.u32.u8 %0, b;\n}\n" : "=r"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ uint32_t cvt_e4m3x2(float lo, float hi) {      // 16 bits: lo in the low byte
  uint32_t d;
  asm("{\n.reg .b16 h;\ncvt.rn.satfinite.e4m3x2.f32 h, %2, %1;\ncvt.u32.u16 %0, h;\n}\n" : "=r"(d) : "f"(lo), "f"(hi));
  return d;
}

// one 16-byte piece of a block per lane; LPB = lanes per block of 32 values
template <int DT, bool FP4>
__global__ void __launch_bounds__(256) k_quant_act(const ActParams P) {
  constexpr int EPL = Elt<DT>::EPL, LPB = 32 / EPL;
  constexpr int EMAX = FP4 ? 2 : 8;
  const int pieces_per_row = P.K / EPL;
  int st = 0;
  // a warp task = 32 consecutive pieces (512 bytes) of one row; lanes past the end of the row idle as whole blocks
  // (pieces_per_row is a multiple of LPB and LPB divides 32), the loop condition is warp-uniform (shuffles inside)
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  for (int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < P.tasks; w += nwarps) {
    const int row = fast_div(w, P.cpr_m, P.cpr_k);
    const int pc = (w - row * P.cpr) * 32 + lane;
    const bool active = pc < pieces_per_row;
    float v[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = 0.0f;
    if (active) load16<DT>(reinterpret_cast<const char*>(P.x) + ((long long)row * P.ldx + (long long)pc * EPL) * (DT == 0 ? 4 : 2), v);
    // exponent field of the block maximum = maximum of the exponent fields (255 as soon as one value is Inf / NaN)
    uint32_t eb = 0;
#pragma unroll
    for (int i = 0; i < EPL; ++i) eb = max(eb, ((uint32_t)__float_as_int(v[i]) >> 23) & 0xFFu);
#pragma unroll
    for (int o = LPB / 2; o > 0; o >>= 1) eb = max(eb, __shfl_xor_sync(0xffffffffu, eb, o));
    // shared exponent exactly as core_plain mode 1 (shared_exp.cuh:14-53)
    const int be = (int)eb;                                        // = exponent field of the block maximum (255: Inf / NaN)
    const bool bad = be == 255;
    const bool flush = (be == 0) && P.flush;
    int sh = bad ? 255 : be - EMAX;
    const int ub = sh - 127;
    sh = ub > P.S ? 255 : sh;
    sh = ub < -P.S ? 127 - P.S : sh;
    const bool dead = sh == 255;
    if (dead) st |= bad ? MSQ_ST_NONFINITE_INPUT : MSQ_ST_EXP_IN_OVERFLOW;
    // sh < 0 cannot happen for S <= 127 (sh >= 127 - S >= 0)
    const float inv = (flush || dead) ? 0.0f : pow2i(127 - sh);
    if (FP4) {
      // EPL values -> EPL / 2 bytes
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < EPL; i += 2) {
        const float a = __fmul_rn(__int_as_float(__float_as_int(v[i]) | 1), inv);
        const float b = __fmul_rn(__int_as_float(__float_as_int(v[i + 1]) | 1), inv);
        acc |= cvt_e2m1x2(a, b) << (4 * i);
      }
      // a zero magnitude carries no sign
      const uint32_t nz = ((acc & 0x77777777u) + 0x77777777u) & 0x88888888u;
      acc &= nz | 0x77777777u;
      if (dead) acc = 0u;
      unsigned char* dst = P.codes + (long long)row * P.ldc + (long long)pc * (EPL / 2);
      if (active) {
        if (EPL == 8) *reinterpret_cast<uint32_t*>(dst) = acc;
        else *reinterpret_cast<uint16_t*>(dst) = (uint16_t)acc;
      }
    } else {
      uint32_t wd[EPL / 4];
#pragma unroll
      for (int j = 0; j < EPL / 4; ++j) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
          const float a = __fmul_rn(__int_as_float(__float_as_int(v[4 * j + i]) | 1), inv);
          const float b = __fmul_rn(__int_as_float(__float_as_int(v[4 * j + i + 1]) | 1), inv);
          acc |= cvt_e4m3x2(a, b) << (8 * i);
        }
        const uint32_t nz = ((acc & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) & 0x80808080u;
        acc &= nz | 0x7F7F7F7Fu;
        wd[j] = dead ? 0u : acc;
      }
      unsigned char* dst = P.codes + (long long)row * P.ldc + (long long)pc * EPL;
      if (active) {
        if (EPL == 8) *reinterpret_cast<uint2*>(dst) = make_uint2(wd[0], wd[EPL / 4 - 1]);
        else *reinterpret_cast<uint32_t*>(dst) = wd[0];
      }
    }
    if (active && (pc & (LPB - 1)) == 0) P.e[(long long)row * P.lde + pc / LPB] = (signed char)(dead ? MSQ_EXP_NAN : sh - 127);
  }
  if (P.status) {
    st = __reduce_or_sync(0xffffffffu, st);
    if ((threadIdx.x & 31) == 0 && st) atomicOr(P.status, st);
  }
}

template <int DT>
static int launch(ActParams P, bool fp4, cudaStream_t s) {
  constexpr int EPL = Elt<DT>::EPL;
  P.cpr = (P.K / EPL + 31) / 32;
  const long long tasks = P.rows * P.cpr;
  if (tasks >= 0x7fffffffLL) return MSQ_ERR_UNSUPPORTED;
  P.tasks = (int)tasks;
  magic_div(P.cpr, &P.cpr_m, &P.cpr_k);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (tasks + 7) / 8;                   // 8 warps per CTA
  const long long cap = (long long)sms * 16;            // a few tasks per warp once the tensor is large
  if (blocks > cap) blocks = cap;
  if (fp4) k_quant_act<DT, true><<<(unsigned)blocks, 256, 0, s>>>(P);
  else k_quant_act<DT, false><<<(unsigned)blocks, 256, 0, s>>>(P);
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
  if ((reinterpret_cast<uintptr_t>(x) & 15) || ((ldx * esz) & 15) || (reinterpret_cast<uintptr_t>(codes) & 7) ||
      (ldc_bytes & 7))
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
