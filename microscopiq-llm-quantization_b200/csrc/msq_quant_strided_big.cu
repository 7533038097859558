// msq_quant_strided_big.cu -- MX(+outlier) fake-quant with LONG blocks (64 / 128 values: "group 128",
// llm/opt_direct.py:98-105 style block sizes) that run along a STRIDED axis (post > 1; for a weight [rows, cols]
// quantized with axes=[0]: along the rows).  Path 6 of msq_quantize_mx_outlier_f32 / msq_quantize_mx_f32.
//
// A thread cannot hold such a block (the 8/16/32 mappings of msq_quant_strided.cu keep one block per thread), and lanes
// that share a block must read DIFFERENT rows, i.e. addresses `post` floats apart.  So a CTA stages a tile of BS rows x 64
// columns in shared memory with coalesced 16-byte loads, and LPB = BS/16 adjacent lanes own one column: lane l holds rows
// 16 l .. 16 l + 15 in registers and the block statistics continue across the lanes with the same xor-butterflies as the
// contiguous-axis kernel (k_inner) -- the canonical adjacent pairwise tree, so results are bit-identical to every other
// mapping and to the oracle.  Shared-memory rows are rotated by (32 / LPB) * (row / 16) floats: a warp = 32/LPB columns x
// LPB lanes then reads 32 different banks in every step, and the rotation (a multiple of 4) keeps the 16-byte chunks of
// the load / store phases intact.  Results go back to the same shared-memory positions and leave with 16-byte stores.
#include "msq_quant_kernels.cuh"

namespace msq {

constexpr int BIG_TW = 64;                     // tile width (columns)

template <int BS, bool OUTLIER>
__global__ void __launch_bounds__(256, 4) k_strided_big(const QParams P) {
  constexpr int N = 16, LPB = BS / N, CPW = 32 / LPB;          // values per lane, lanes per block, columns per warp
  constexpr int PASSES = BIG_TW / (8 * CPW);
  extern __shared__ float4 s_tile4[];
  float* s_tile = reinterpret_cast<float*>(s_tile4);           // [BS][64], row r rotated by CPW * (r / 16)
  const int post = (int)P.post;
  const int ctiles = (post + BIG_TW - 1) / BIG_TW;
  // blockIdx.x -> (p, b, column tile)
  const int ct = blockIdx.x % ctiles;
  const int pb = blockIdx.x / ctiles;
  const int p = pb / P.nblk, b = pb - p * P.nblk;
  const int c0 = ct * BIG_TW;
  const int rows_valid = min(BS, P.L - b * BS);                // zero padding below (utils/quant.py:563-583)
  const long long base = ((long long)p * P.L + (long long)b * BS) * post + c0;
  const int cols_valid = min(BIG_TW, post - c0);               // post % 4 == 0: whole 16-byte chunks
  // ---- load: 16 chunks of 16 bytes per row
  for (int q = threadIdx.x; q < BS * (BIG_TW / 4); q += 256) {
    const int r = q >> 4, c4 = (q & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid && c4 < cols_valid) v = __ldg(reinterpret_cast<const float4*>(P.in + base + (long long)r * post + c4));
    *reinterpret_cast<float4*>(s_tile + r * BIG_TW + ((c4 + CPW * (r >> 4)) & (BIG_TW - 1))) = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = lane % LPB, j = lane / LPB;
  int st = 0;
#pragma unroll 1
  for (int pass = 0; pass < PASSES; ++pass) {
    const int col = (pass * 8 + warp) * CPW + j;
    float x[N];
    uint32_t code[N];
    float* src = s_tile + (N * l) * BIG_TW + ((col + CPW * l) & (BIG_TW - 1));
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = src[i * BIG_TW];
    int e_in, e_out = 0, cnt = 0, bst;
    if (OUTLIER) {
      const BlockOut bo = core_outlier<N, LPB, false, false>(x, code, P.c, nullptr, nullptr);
      e_in = bo.e_in; e_out = bo.e_out; cnt = bo.cnt; bst = bo.st;
    } else {
      const long long t = ((long long)p * P.nblk + b) * post + c0 + col;
      const PlainOut po = core_plain<N, LPB>(x, code, false, P.c, (P.ext_max && col < cols_valid) ? P.ext_max + t : nullptr);
      e_in = po.e; bst = po.st;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) src[i * BIG_TW] = x[i];
    if (col < cols_valid) {
      st |= bst;
      if (l == 0) {
        const long long t = ((long long)p * P.nblk + b) * post + c0 + col;
        if (P.e_in) P.e_in[t] = (int8_t)e_in;
        if (OUTLIER && P.e_out) P.e_out[t] = (int8_t)e_out;
        if (OUTLIER && P.blk_cnt) P.blk_cnt[t] = (uint8_t)cnt;
      }
    }
  }
  __syncthreads();
  if (P.deq) {
    for (int q = threadIdx.x; q < BS * (BIG_TW / 4); q += 256) {
      const int r = q >> 4, c4 = (q & 15) * 4;
      if (r < rows_valid && c4 < cols_valid)
        *reinterpret_cast<float4*>(P.deq + base + (long long)r * post + c4) =
            *reinterpret_cast<const float4*>(s_tile + r * BIG_TW + ((c4 + CPW * (r >> 4)) & (BIG_TW - 1)));
    }
  }
  if (P.status) {
    st = __reduce_or_sync(0xffffffffu, st);
    if (lane == 0 && st) atomicOr(P.status, st);
  }
}

// bs in {64, 128}; post % 4 == 0; 16-byte aligned in / deq; no mask / codes / external bounds
cudaError_t launch_strided_big(const QParams& P, bool outlier, cudaStream_t stream) {
  void (*k)(const QParams) = nullptr;
  if (P.bs == 64) k = outlier ? k_strided_big<64, true> : k_strided_big<64, false>;
  else k = outlier ? k_strided_big<128, true> : k_strided_big<128, false>;
  const size_t smem = (size_t)P.bs * BIG_TW * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long ctiles = (P.post + BIG_TW - 1) / BIG_TW;
  const long long grid = P.pre * P.nblk * ctiles;
  k<<<(unsigned)grid, 256, smem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace msq
