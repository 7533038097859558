// msq_quant_inner.cu -- blocks contiguous along the innermost axis (post == 1, L % bs == 0):
// the K-blocked layout of MXLinear weights / activations (number_system/mx/linear.py:66-85).
// bs/8 lanes share a block, two 128-bit loads each (8 elements per lane halve the per-lane replicated
// block tail and remove one butterfly stage); block reductions are xor-butterflies.
#include "msq_quant_kernels.cuh"

namespace msq {

// --------------------------------------------------------------------------
// innermost mapping: post == 1, L % bs == 0, blocks tile the tensor contiguously
// --------------------------------------------------------------------------
template <int LPB, bool OUTLIER, bool PACK, bool EXT>
__global__ void __launch_bounds__(256) k_inner(const QParams P) {
  constexpr int N = 8;                        // elements per lane
  constexpr int BS = LPB * N;
  extern __shared__ uint32_t s_hi[];
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const long long seg = blockIdx.x;
  const bool pack_hi = PACK && OUTLIER && P.hi_w > 0;
  constexpr bool want_codes = PACK;
  const int wi = P.c.fin.width;
  if (pack_hi) {
    const int words = (int)(P.hi_seg_stride >> 2);
    for (int i = threadIdx.x; i < words; i += 256) s_hi[i] = 0;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
  }
  int st = 0;
  const long long nelem = P.nblocks * BS;
  const long long seg_e0 = seg * MSQ_SEG_BLOCKS * BS;
#pragma unroll ((OUTLIER && !PACK) ? 2 : 1)
  for (int it = 0; it < (MSQ_SEG_BLOCKS * BS) / (256 * N); ++it) {
    const long long e0 = seg_e0 + (long long)it * (256 * N) + threadIdx.x * N;
    const bool active = e0 < nelem;
    float x[N];
    uint32_t code[N];
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (active) v = __ldg(reinterpret_cast<const float4*>(P.in + e0) + q);
      x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    const long long t = e0 / BS;             // block index (== e0 >> log2)
    BlockOut bo;
    if (OUTLIER) {
      float lo[N], hi[N];
#pragma unroll
      for (int i = 0; i < N; ++i) lo[i] = hi[i] = 0.0f;
      if (EXT && active) {
        const long long nb_row = P.nblk;      // blocks per row
        const long long p = t / nb_row;
        const int i0 = (int)(e0 - t * BS);
#pragma unroll
        for (int i = 0; i < N; ++i) { lo[i] = P.lo_ext[p * BS + i0 + i]; hi[i] = P.hi_ext[p * BS + i0 + i]; }
      }
      bo = core_outlier<N, LPB, PACK, EXT>(x, code, P.c, lo, hi);
    } else {
      PlainOut po = core_plain<N, LPB>(x, code, want_codes, P.c, (P.ext_max && active) ? P.ext_max + t : nullptr);
      bo.mask = 0; bo.e_in = po.e; bo.e_out = 0; bo.cnt = 0; bo.st = po.st;
    }
    if (active) {
      st |= bo.st;
      if (!PACK || P.deq) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q)
          reinterpret_cast<float4*>(P.deq + e0)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
      }
      if (PACK) {
        uint64_t w = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) w |= (uint64_t)(code[i] & ((1u << wi) - 1u)) << (i * wi);
        const long long byte = (e0 * wi) >> 3;
        if (wi == 2) *reinterpret_cast<uint16_t*>(P.codes + byte) = (uint16_t)w;
        else if (wi == 4) *reinterpret_cast<uint32_t*>(P.codes + byte) = (uint32_t)w;
        else *reinterpret_cast<uint2*>(P.codes + byte) = make_uint2((uint32_t)w, (uint32_t)(w >> 32));
      }
      if ((threadIdx.x % LPB) == 0) {
        if (P.e_in) P.e_in[t] = (int8_t)bo.e_in;
        if (OUTLIER && P.e_out) P.e_out[t] = (int8_t)bo.e_out;
        if (OUTLIER && P.blk_cnt) P.blk_cnt[t] = (uint8_t)bo.cnt;
      }
    }
    constexpr uint32_t NM = (1u << N) - 1u;   // this lane's bits of the block mask
    if (PACK && OUTLIER) {
      constexpr int LPW = 32 / N;             // lanes per bitmap word
      uint32_t w = (bo.mask & NM) << ((threadIdx.x & (LPW - 1)) * N);
#pragma unroll
      for (int o = 1; o < LPW; o <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, o);
      if (active && (threadIdx.x & (LPW - 1)) == 0) P.bitmap[e0 >> 5] = w;
    }
    if (pack_hi) {
      int tot;
      const int mine = active ? __popc(bo.mask & NM) : 0;
      int off = cta_excl_scan(mine, s_warp, &tot);
      if (active && mine) {
        const int bit0 = (s_base + off) * P.hi_w;
        int w = bit0 >> 5, b = bit0 & 31;
        uint32_t cur = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          if ((bo.mask >> i) & 1u) {
            cur |= (code[i] >> wi) << b;
            b += P.hi_w;
            if (b == 32) { atomicOr(&s_hi[w], cur); cur = 0; b = 0; ++w; }
          }
        }
        if (b) atomicOr(&s_hi[w], cur);
      }
      __syncthreads();
      if (threadIdx.x == 0) s_base += tot;
      __syncthreads();
    }
  }
  if (pack_hi) {
    const int total = s_base;
    const int words = (total * P.hi_w + 31) >> 5;
    uint32_t* dst = reinterpret_cast<uint32_t*>(P.hi + seg * P.hi_seg_stride);
    for (int i = threadIdx.x; i < words; i += 256) dst[i] = s_hi[i];
    if (threadIdx.x == 0 && P.seg_cnt) P.seg_cnt[seg] = total;
  } else if (PACK && OUTLIER && P.seg_cnt && threadIdx.x == 0) {
    P.seg_cnt[seg] = 0;
  }
  if (P.status) {
    st = __reduce_or_sync(0xffffffffu, st);
    if ((threadIdx.x & 31) == 0 && st) atomicOr(P.status, st);
  }
}


cudaError_t launch_inner(const QParams& P, bool outlier, size_t smem, cudaStream_t stream) {
  const long long nseg = (P.nblocks + MSQ_SEG_BLOCKS - 1) / MSQ_SEG_BLOCKS;
  void (*k)(const QParams) = nullptr;
  const bool codes = P.codes != nullptr;
  const bool ext = outlier && P.c.variant != 0;
#define MSQ_PICK(LPB_)                                                                            \
  if (!outlier) k = codes ? k_inner<LPB_, false, true, false> : k_inner<LPB_, false, false, false>; \
  else if (ext) k = codes ? k_inner<LPB_, true, true, true> : k_inner<LPB_, true, false, true>;     \
  else k = codes ? k_inner<LPB_, true, true, false> : k_inner<LPB_, true, false, false>;
  switch (P.bs) {
    case 8: MSQ_PICK(1) break;
    case 16: MSQ_PICK(2) break;
    case 32: MSQ_PICK(4) break;
    case 64: MSQ_PICK(8) break;
    case 128: MSQ_PICK(16) break;
    default: return cudaErrorInvalidValue;
  }
#undef MSQ_PICK
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<(unsigned)nseg, 256, smem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace msq
