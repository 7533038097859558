// msq_reduce.cu -- sum / max over the innermost dimension: the two reduction entry points of the reference's
// pybind module (reduce_sum_inner_dim / reduce_max_inner_dim, number_system/mx/cpp/funcs.cpp:197-211, kernels
// cpp/reduce.cuh:150-210).  Contract of the reference: 32 <= inner <= 1024, inner a multiple of 32.
//
// One warp per row: every 32-element chunk of the row is one coalesced 128-byte load reduced by a shuffle-down tree
// (the reference's first level, reduce.cuh:38-43), the per-chunk results -- padded with the start value -- go through
// the same tree once more (second level, reduce.cuh:118-129).  Same association order as the reference, so the fp32
// sum is bit-identical to it; no shared memory, no block-level barrier, no atomics.
#include "msq_api_internal.h"
#include "msq_common.cuh"

namespace msq {

template <int OP>   // 0 sum, 1 max
__device__ __forceinline__ float red2(float a, float b) {
  if (OP == 0) return __fadd_rn(a, b);
  return b > a ? b : a;
}

template <int OP>
__device__ __forceinline__ float warp_tree(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red2<OP>(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

template <int OP>
__global__ void __launch_bounds__(256) k_reduce_inner(const float* __restrict__ in, float* __restrict__ out, long long rows,
                                                      int inner) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  const int chunks = inner >> 5;
  const float start = OP == 0 ? 0.0f : -65504.0f;          // reduce.cuh:86-87
  for (long long r = warp; r < rows; r += nwarps) {
    const float* src = in + r * inner;
    float part = start;
    for (int c = 0; c < chunks; ++c) {
      float v = warp_tree<OP>(__ldg(src + c * 32 + lane));
      v = __shfl_sync(0xffffffffu, v, 0);
      if (lane == c) part = v;
    }
    // lanes >= chunks hold the start value, as in the reference (a max therefore floors at -65504 unless inner == 1024)
    const float tot = warp_tree<OP>(part);
    if (lane == 0) out[r] = tot;
  }
}

}  // namespace msq

extern "C" int msq_reduce_inner_dim_f32(const float* in, float* out, long long rows, int inner, int op, void* stream) {
  msq::DevGuard dev_guard_(in);
  if (!in || !out || rows < 0 || (op != 0 && op != 1)) return MSQ_ERR_BAD_ARG;
  if (inner < 32 || inner > 1024 || (inner % 32)) return MSQ_ERR_UNSUPPORTED;     // AT_ASSERTM, reduce.cu:34-35
  if (rows == 0) return MSQ_OK;
  long long blocks = (rows + 7) / 8;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (op == 0) msq::k_reduce_inner<0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, rows, inner);
  else msq::k_reduce_inner<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, rows, inner);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}
