// msq_quant_strided.cu -- blocks strided by `post` (e.g. 16 output rows of one weight column,
// the layout the reference's GPTQ / RTN drivers use: axes=[0] on [rows, cols]).
//
// Lane = column: a warp reads 32 consecutive columns of one row per load (128 B, fully
// coalesced), a thread keeps its whole block (BS rows) in registers.  256 threads = 256
// consecutive block-linear indices; 4 sub-iterations cover one packed segment.
//
// PACK = false : fake-quant (fp32 out, what the reference API returns) + optional per-block
//                exponents / outlier counts.
// PACK = true  : packed planes (codes, bitmap, hi stream, exponents); fp32 out optional.
// Element offsets are 32-bit (host guarantees numel < 2^31 on this path).
#include "msq_quant_kernels.cuh"

#ifndef MSQ_FQ_THREADS
#define MSQ_FQ_THREADS 256     // CTA size of the fake-quant (non-packing) launches
#endif
#ifndef MSQ_STRIDED_MINB
#define MSQ_STRIDED_MINB 4
#endif
namespace msq {

// one row of 32 codes (a warp's 32 columns) -> 32*WI bits: lanes combine into 32-bit words
template <int BS, int WI>
__device__ __forceinline__ void store_codes(uint8_t* codes, const uint32_t (&code)[BS], int base, int post, int nv, int lane) {
  constexpr int G = 32 / WI;                 // lanes per 32-bit word
#pragma unroll
  for (int i = 0; i < BS; ++i) {
    uint32_t w = (code[i] & ((1u << WI) - 1u)) << ((lane & (G - 1)) * WI);
#pragma unroll
    for (int o = 1; o < G; o <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, o);
    if (i < nv && (lane & (G - 1)) == 0)
      *reinterpret_cast<uint32_t*>(codes + (((long long)(base + i * post) * WI) >> 3)) = w;
  }
}

// 4-bit slots, 16 rows: an 8x8 byte transpose across each group of 8 lanes (3 shuffle + PRMT stages instead of
// 3 shuffles per row) leaves lane g of a group with rows 2g and 2g+1 of the group's 8 columns; two word stores
// per lane.  c[] = the 16 codes of this lane's column, rows r0 .. r0+15 of the block.
__device__ __forceinline__ void store_codes4_16(uint8_t* codes, const uint32_t* c, int rowbase, int r0, int post, int nv,
                                                int lane) {
  const int g = lane & 7;
  uint32_t W0, W1;
  {
    uint32_t B[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) B[k] = (c[2 * k] & 0xFu) | ((c[2 * k + 1] << 4) & 0xF0u);
    W0 = __byte_perm(__byte_perm(B[0], B[1], 0x0040), __byte_perm(B[2], B[3], 0x0040), 0x5410);
    W1 = __byte_perm(__byte_perm(B[4], B[5], 0x0040), __byte_perm(B[6], B[7], 0x0040), 0x5410);
  }
  // stage A: swap the off-diagonal 4x4 blocks (lanes g and g^4 exchange one word)
  {
    const uint32_t r = __shfl_xor_sync(0xffffffffu, (g & 4) ? W0 : W1, 4);
    if (g & 4) W0 = r; else W1 = r;
  }
  // stage B: swap 2x2 sub-blocks (16-bit halves of both words travel in one shuffle)
  {
    const uint32_t r = __shfl_xor_sync(0xffffffffu, __byte_perm(W0, W1, (g & 2) ? 0x5410 : 0x7632), 2);
    W0 = __byte_perm(W0, r, (g & 2) ? 0x3254 : 0x5410);
    W1 = __byte_perm(W1, r, (g & 2) ? 0x3276 : 0x7610);
  }
  // stage C: swap single bytes
  {
    const uint32_t r = __shfl_xor_sync(0xffffffffu, __byte_perm(W0, W1, (g & 1) ? 0x6420 : 0x7531), 1);
    W0 = __byte_perm(W0, r, (g & 1) ? 0x3514 : 0x5240);
    W1 = __byte_perm(W1, r, (g & 1) ? 0x3716 : 0x7260);
  }
  // byte k of (W0, W1) = column k of the group; low nibble = row 2g, high nibble = row 2g+1
  const uint32_t a0 = W0 & 0x0F0F0F0Fu, a1 = W1 & 0x0F0F0F0Fu;
  const uint32_t b0 = (W0 >> 4) & 0x0F0F0F0Fu, b1 = (W1 >> 4) & 0x0F0F0F0Fu;
  const uint32_t row_lo = __byte_perm(a0 | (a0 >> 4), a1 | (a1 >> 4), 0x6420);
  const uint32_t row_hi = __byte_perm(b0 | (b0 >> 4), b1 | (b1 >> 4), 0x6420);
  uint32_t* dst = reinterpret_cast<uint32_t*>(codes) + (lane >> 3);
  const int r = r0 + 2 * g;
  if (r < nv) dst[(rowbase + r * post) >> 3] = row_lo;
  if (r + 1 < nv) dst[(rowbase + (r + 1) * post) >> 3] = row_hi;
}

template <int BS, bool OUTLIER, bool PACK, bool EXT, bool CNT>
__global__ void __launch_bounds__(256, (BS <= 16 ? (PACK ? 4 : MSQ_STRIDED_MINB) : 2)) k_strided(const QParams P) {
  extern __shared__ uint32_t s_hi[];
  __shared__ int s_warp[8];
  __shared__ int s_base;
  // deferred hi placement (blocks whose hi fields fit 64 bits): every block stashes its compacted fields and its count,
  // ONE scan over the segment's 1024 counts places them afterwards -- instead of a CTA scan + 3 barriers per 256 blocks
  __shared__ unsigned long long s_acc[PACK && OUTLIER ? MSQ_SEG_BLOCKS : 1];
  __shared__ unsigned char s_cnt[PACK && OUTLIER ? MSQ_SEG_BLOCKS : 1];
  const long long seg = blockIdx.x;
  const bool pack_hi = PACK && OUTLIER && P.hi_w > 0;
  const bool deferred = pack_hi && BS * P.hi_w <= 64;
  const int wi = P.c.fin.width;
  const int lane = threadIdx.x & 31;
  const int post = (int)P.post;
  if (pack_hi) {
    const int words = (int)(P.hi_seg_stride >> 2);
    for (int i = threadIdx.x; i < words; i += 256) s_hi[i] = 0;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
  }
  int st = 0;
  // PACK: one CTA = one packed segment (4 sub-iterations of 256 consecutive block-linear indices).
  // fake-quant: 3-D grid, blockIdx.z = p, blockIdx.y = b, blockIdx.x*256 + tid = column: no divisions.
  constexpr int NSUB = PACK ? MSQ_SEG_BLOCKS / 256 : 1;
#pragma unroll 1
  for (int sub = 0; sub < NSUB; ++sub) {
    // 32-bit index math: the host routes tensors with >= 2^31 elements to the generic kernel
    int t0, base = 0, p = 0, j = 0, b = 0;
    bool active;
    if (PACK) {
      t0 = (int)seg * MSQ_SEG_BLOCKS + sub * 256 + (int)threadIdx.x;
      active = t0 < (int)P.nblocks;
      if (active) {
        const int pb = fast_div(t0, P.post_m, P.post_k);
        j = t0 - pb * post;
        p = fast_div(pb, P.nblk_m, P.nblk_k); b = pb - p * P.nblk;
      }
    } else {
      b = (int)blockIdx.y; p = (int)blockIdx.z;   // uniform per CTA, no divisions
      j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
      active = j < post;
      t0 = (p * P.nblk + b) * post + j;
    }
    if (active) base = (p * P.L + b * BS) * post + j;
    float x[BS];
    uint32_t code[BS];
    const int nv = active ? min(BS, P.L - b * BS) : 0;       // valid rows, the rest is zero padding
    // row addresses as base + (int64)(stride_bytes * i): one IMAD.WIDE per row instead of a 4-instruction
    // multiply / sign-extend / 64-bit add sequence
    const int sb = post * 4;
    const char* pin = reinterpret_cast<const char*>(P.in + base);
    auto row_in = [&](int i) { return reinterpret_cast<const float*>(pin + (long long)sb * i); };
    if (nv == BS) {
#pragma unroll
      for (int i = 0; i < BS; ++i) x[i] = __ldg(row_in(i));
    } else {
#pragma unroll
      for (int i = 0; i < BS; ++i) x[i] = (i < nv) ? __ldg(row_in(i)) : 0.0f;
    }

    BlockOut bo;
    if (OUTLIER) {
      float lo[BS], hi[BS];
      if (EXT) {
#pragma unroll
        for (int i = 0; i < BS; ++i) {
          const int o = active ? (p * BS + i) * post + j : 0;
          lo[i] = P.lo_ext[o]; hi[i] = P.hi_ext[o];
        }
      }
      bo = core_outlier<BS, 1, PACK, EXT, CNT>(x, code, P.c, lo, hi);
    } else {
      PlainOut po = core_plain<BS, 1>(x, code, PACK, P.c, (P.ext_max && active) ? P.ext_max + t0 : nullptr);
      bo.mask = 0; bo.e_in = po.e; bo.e_out = 0; bo.cnt = 0; bo.st = po.st;
    }
    if (active) st |= bo.st;

    // ---- stores: lane = column, so every row is one coalesced warp store ----
    if (!PACK || P.deq) {
      char* pout = reinterpret_cast<char*>(P.deq + base);
      auto row_out = [&](int i) { return reinterpret_cast<float*>(pout + (long long)sb * i); };
      if (nv == BS) {
#pragma unroll
        for (int i = 0; i < BS; ++i) *row_out(i) = x[i];
      } else {
#pragma unroll
        for (int i = 0; i < BS; ++i)
          if (i < nv) *row_out(i) = x[i];
      }
    }
    if (active) {
      if (P.e_in) P.e_in[t0] = (int8_t)bo.e_in;
      if (OUTLIER && P.e_out) P.e_out[t0] = (int8_t)bo.e_out;
      if (OUTLIER && CNT && P.blk_cnt) P.blk_cnt[t0] = (uint8_t)bo.cnt;
    }
    if (PACK) {
      // host guarantees post % 32 == 0, so a warp's 32 columns share (p, b) and whole words
      if (wi == 4 && BS % 16 == 0) {
#pragma unroll
        for (int h = 0; h < BS / 16; ++h) store_codes4_16(P.codes, code + 16 * h, base - lane, 16 * h, post, nv, lane);
      } else if (wi == 4) store_codes<BS, 4>(P.codes, code, base, post, nv, lane);
      else if (wi == 2) store_codes<BS, 2>(P.codes, code, base, post, nv, lane);
      else store_codes<BS, 8>(P.codes, code, base, post, nv, lane);
      if (OUTLIER) {
        // one ballot per row; lane 0 stores the row's word (address = first word + row * words per row)
        char* bp = reinterpret_cast<char*>(P.bitmap + ((base - lane) >> 5));
        const int bstride = (post >> 5) * 4;
        const bool l0 = lane == 0;
        if (nv == BS) {
#pragma unroll
          for (int i = 0; i < BS; ++i) {
            const uint32_t w = __ballot_sync(0xffffffffu, bo.mask & (1u << i));
            if (l0) *reinterpret_cast<uint32_t*>(bp + (long long)bstride * i) = w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < BS; ++i) {
            const uint32_t w = __ballot_sync(0xffffffffu, bo.mask & (1u << i));
            if (i < nv && l0) *reinterpret_cast<uint32_t*>(bp + (long long)bstride * i) = w;
          }
        }
      }
    }
    if (deferred) {
      unsigned long long acc = 0;
      if (active && bo.cnt) {
        // walk the block backwards, shifting the accumulator by hi_w for every outlier (an inlier's code has no bits
        // above the slot width, so its field is zero and needs no masking): first outlier ends in the lowest bits
#pragma unroll
        for (int i = BS - 1; i >= 0; --i) {
          const int sh = ((bo.mask >> i) & 1u) ? P.hi_w : 0;
          acc = (acc << sh) | (unsigned long long)(code[i] >> wi);
        }
      }
      s_acc[sub * 256 + threadIdx.x] = acc;
      s_cnt[sub * 256 + threadIdx.x] = (unsigned char)(active ? bo.cnt : 0);
    } else if (pack_hi) {
      int tot;
      const int off = cta_excl_scan(active ? bo.cnt : 0, s_warp, &tot);
      if (active && bo.cnt) {
        // this block's hi fields are contiguous in the stream, first outlier in the lowest bits
        const int bit0 = (s_base + off) * P.hi_w;
        if (BS * P.hi_w <= 64) {
          // branch-free: walk the block backwards, shifting the accumulator by hi_w for every outlier
          // (an inlier's code has no bits above the slot width, so its field is zero and needs no masking)
          unsigned long long acc = 0;
#pragma unroll
          for (int i = BS - 1; i >= 0; --i) {
            const int sh = ((bo.mask >> i) & 1u) ? P.hi_w : 0;
            acc = (acc << sh) | (unsigned long long)(code[i] >> wi);
          }
          const int w = bit0 >> 5, b = bit0 & 31;
          const unsigned long long lo = acc << b;
          const uint32_t top = b ? (uint32_t)(acc >> (64 - b)) : 0u;
          atomicOr(&s_hi[w], (uint32_t)lo);
          if ((uint32_t)(lo >> 32)) atomicOr(&s_hi[w + 1], (uint32_t)(lo >> 32));
          if (top) atomicOr(&s_hi[w + 2], top);
        } else {
          int w = bit0 >> 5, b = bit0 & 31;
          uint32_t cur = 0;
#pragma unroll
          for (int i = 0; i < BS; ++i) {
            if ((bo.mask >> i) & 1u) {
              cur |= (code[i] >> wi) << b;
              b += P.hi_w;
              if (b == 32) { atomicOr(&s_hi[w], cur); cur = 0; b = 0; ++w; }
            }
          }
          if (b) atomicOr(&s_hi[w], cur);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) s_base += tot;
      __syncthreads();
    }
  }
  if (deferred) {
    __syncthreads();
    // thread t places the four consecutive blocks 4t .. 4t+3 of the segment
    int c4[4], mine = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { c4[u] = s_cnt[4 * threadIdx.x + u]; mine += c4[u]; }
    int tot;
    int off = cta_excl_scan(mine, s_warp, &tot);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c4[u]) {
        const unsigned long long acc = s_acc[4 * threadIdx.x + u];
        const int bit0 = off * P.hi_w;
        const int w = bit0 >> 5, b = bit0 & 31;
        const unsigned long long lo = acc << b;
        const uint32_t top = b ? (uint32_t)(acc >> (64 - b)) : 0u;
        atomicOr(&s_hi[w], (uint32_t)lo);
        if ((uint32_t)(lo >> 32)) atomicOr(&s_hi[w + 1], (uint32_t)(lo >> 32));
        if (top) atomicOr(&s_hi[w + 2], top);
        off += c4[u];
      }
    }
    if (threadIdx.x == 0) s_base = tot;
    __syncthreads();
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
    if (lane == 0 && st) atomicOr(P.status, st);
  }
}

// ---- several weight matrices in ONE launch (fake-quant, blocks along the rows) -----------------------------------
// The drivers quantize the linears of a decoder layer back to back (llm/llama.py:238-253); one launch over all of
// them removes the launch gaps and the partially filled last wave of every matrix.  Same per-block arithmetic as
// k_strided (core_outlier), CTA -> (matrix, row block, 256 columns) through a table in the parameter block.
template <int BS>
__global__ void __launch_bounds__(256, (BS <= 16 ? MSQ_STRIDED_MINB : 2)) k_strided_multi(const __grid_constant__ MultiQParams MP) {
  int t = 0;
#pragma unroll 1
  while (t + 1 < MP.n && blockIdx.x >= MP.cta_start[t + 1]) ++t;
  const int local = (int)(blockIdx.x - MP.cta_start[t]);
  const int gx = MP.gx[t];
  const int b = local / gx;
  const int post = MP.post[t], L = MP.L[t];
  const int j = (local - b * gx) * 256 + (int)threadIdx.x;
  const bool active = j < post;
  const int base = active ? b * BS * post + j : 0;
  const int nv = active ? min(BS, L - b * BS) : 0;
  float x[BS];
  uint32_t code[BS];
  const int sb = post * 4;
  const char* pin = reinterpret_cast<const char*>(MP.in[t] + base);
  if (nv == BS) {
#pragma unroll
    for (int i = 0; i < BS; ++i) x[i] = __ldg(reinterpret_cast<const float*>(pin + (long long)sb * i));
  } else {
#pragma unroll
    for (int i = 0; i < BS; ++i) x[i] = (i < nv) ? __ldg(reinterpret_cast<const float*>(pin + (long long)sb * i)) : 0.0f;
  }
  const BlockOut bo = core_outlier<BS, 1, false, false, false>(x, code, MP.c, nullptr, nullptr);
  char* pout = reinterpret_cast<char*>(MP.deq[t] + base);
  if (nv == BS) {
#pragma unroll
    for (int i = 0; i < BS; ++i) *reinterpret_cast<float*>(pout + (long long)sb * i) = x[i];
  } else {
#pragma unroll
    for (int i = 0; i < BS; ++i)
      if (i < nv) *reinterpret_cast<float*>(pout + (long long)sb * i) = x[i];
  }
  if (MP.status) {
    int st = active ? bo.st : 0;
    st = __reduce_or_sync(0xffffffffu, st);
    if ((threadIdx.x & 31) == 0 && st) atomicOr(MP.status, st);
  }
}

cudaError_t launch_strided_multi(const MultiQParams& MP, int bs, unsigned total_ctas, cudaStream_t stream) {
  if (bs == 16) k_strided_multi<16><<<total_ctas, 256, 0, stream>>>(MP);
  else if (bs == 32) k_strided_multi<32><<<total_ctas, 256, 0, stream>>>(MP);
  else if (bs == 8) k_strided_multi<8><<<total_ctas, 256, 0, stream>>>(MP);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_strided(const QParams& P, bool outlier, size_t smem, cudaStream_t stream) {
  const long long nseg = (P.nblocks + MSQ_SEG_BLOCKS - 1) / MSQ_SEG_BLOCKS;
  void (*k)(const QParams) = nullptr;
  const bool pack = P.codes != nullptr;
  const bool ext = outlier && P.c.variant != 0;
  const bool cnt = pack || P.blk_cnt != nullptr;
#define MSQ_PICK(BS_)                                                                                             \
  if (!outlier) k = pack ? k_strided<BS_, false, true, false, false> : k_strided<BS_, false, false, false, false>; \
  else if (ext) k = pack ? k_strided<BS_, true, true, true, true> : k_strided<BS_, true, false, true, true>;       \
  else if (pack) k = k_strided<BS_, true, true, false, true>;                                                      \
  else k = cnt ? k_strided<BS_, true, false, false, true> : k_strided<BS_, true, false, false, false>;
  if (P.bs == 16) { MSQ_PICK(16) }
  else if (P.bs == 32) { MSQ_PICK(32) }
  else if (P.bs == 8) { MSQ_PICK(8) }
  else return cudaErrorInvalidValue;
#undef MSQ_PICK
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (pack) {
    k<<<(unsigned)nseg, 256, smem, stream>>>(P);
  } else {
    if (P.nblk > 65535 || P.pre > 65535) {     // gridDim.y/z limits (65535 * bs rows); the caller falls back
      return cudaErrorInvalidConfiguration;    // to k_generic
    }
    dim3 grid((unsigned)((P.post + MSQ_FQ_THREADS - 1) / MSQ_FQ_THREADS), (unsigned)P.nblk, (unsigned)P.pre);
    k<<<grid, MSQ_FQ_THREADS, 0, stream>>>(P);
  }
  return cudaGetLastError();
}

}  // namespace msq
