// msq_quant_kernels.cuh -- parameter block and launchers shared by the quantizer kernels.
//
// Geometry: the tensor is viewed as [pre, L, post]; blocks of `bs` run along L
// (zero-padded to a multiple of bs, utils/quant.py:563-583).  Block-linear index
// t = (p*nblk + b)*post + j.  A packed *segment* is MSQ_SEG_BLOCKS consecutive t.
#pragma once
#include "msq_quant_core.cuh"
#include "msq_api_internal.h"

extern "C" {
int msq_packed_hi_width(int in_ebits, int in_mbits, int out_ebits, int out_mbits);
long long msq_packed_hi_seg_stride(int block_size, int hi_w);
}

namespace msq {

struct QParams {
  long long pre, post, nblocks;   // nblocks = pre*nblk*post
  int L, bs, nblk;
  QCfg c;
  const float* in;
  float* deq;          // optional
  uint8_t* codes;      // optional (packed, inlier container width)
  int8_t* e_in;        // optional [nblocks]
  int8_t* e_out;       // optional [nblocks]
  uint32_t* bitmap;    // optional, 1 bit / element, linear element order
  uint8_t* hi;         // optional, segment-compacted high bits of outlier codes
  int* seg_cnt;        // optional [nseg]
  uint8_t* blk_cnt;    // optional [nblocks] outliers per block
  uint8_t* mask;       // optional, 1 byte / element
  int* status;         // optional, OR of MSQ_ST_*
  const float* lo_ext; // variant 1: [pre, bs, post]
  const float* hi_ext;
  const float* ext_max;// plain: optional external block max [nblocks]
  long long hi_seg_stride;  // bytes
  int hi_w;            // container width of one hi field (0 = no hi stream)
  // exact division of a 31-bit index by post / nblk: q = (n * m) >> k   (fill_geometry)
  unsigned int post_m, nblk_m;
  int post_k, nblk_k;
};

// floor(n / d) for 0 <= n < 2^31 with the (m, k) pair of magic_div()
__device__ __forceinline__ int fast_div(int n, unsigned int m, int k) {
  return (int)(((unsigned long long)(unsigned int)n * m) >> k);
}
__host__ inline void magic_div(long long d, unsigned int* m, int* k) {
  // k = 31 + ceil(log2 d), m = floor(2^k / d) + 1: the error term n*e/2^k stays below 1/d for n < 2^31
  int l = 0;
  while ((1LL << l) < d) ++l;
  *k = 31 + l;
  *m = (unsigned int)((((unsigned __int128)1) << *k) / (unsigned __int128)d + 1);
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// exclusive scan over the 256 threads of the CTA; returns offset, *sum = CTA total
__device__ __forceinline__ int cta_excl_scan(int v, int* s_warp, int* sum) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int wt;
  int ex = warp_excl_scan(v, lane, wt);
  __syncthreads();
  if (lane == 0) s_warp[w] = wt;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { int t = s_warp[i]; if (i < w) base += t; tot += t; }
  *sum = tot;
  return ex + base;
}


// several [rows, cols] matrices quantized along the rows by one launch (k_strided_multi)
constexpr int MSQ_MULTI_MAX = 16;
struct MultiQParams {
  int n;
  QCfg c;
  int* status;
  const float* in[MSQ_MULTI_MAX];
  float* deq[MSQ_MULTI_MAX];
  int L[MSQ_MULTI_MAX], post[MSQ_MULTI_MAX], gx[MSQ_MULTI_MAX];
  unsigned int cta_start[MSQ_MULTI_MAX + 1];
};
cudaError_t launch_strided_multi(const MultiQParams& MP, int bs, unsigned total_ctas, cudaStream_t stream);
// persistent TMA-staged variant (msq_quant_tma.cu); MSQ_ERR_UNSUPPORTED = geometry not addressable by a tensor map
int launch_strided_tma(int n, const float* const* in, float* const* out, const long long* rows, const long long* cols,
                       int bs, const QCfg& c, int* status, cudaStream_t stream);

// quantize + pack of one 2-D weight through the TMA-staged kernel; MSQ_ERR_UNSUPPORTED = not its geometry / formats
int launch_pack_tma(const QParams& Q, cudaStream_t stream);

// launchers (one translation unit each); return false if the geometry is not theirs
cudaError_t launch_strided(const QParams& P, bool outlier, size_t smem, cudaStream_t stream);
cudaError_t launch_inner(const QParams& P, bool outlier, size_t smem, cudaStream_t stream);
cudaError_t launch_generic(const QParams& P, bool outlier, cudaStream_t stream);
// long blocks (64 / 128) along a strided axis, values (+ exponents / counts) only: msq_quant_strided_big.cu, path 6
cudaError_t launch_strided_big(const QParams& P, bool outlier, cudaStream_t stream);

}  // namespace msq
