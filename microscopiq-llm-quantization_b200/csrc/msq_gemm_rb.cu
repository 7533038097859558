// msq_gemm_rb.cu -- packed-weight linear on tcgen05 for the artefact the quantizers EMIT (the reference
// quantizes with axes=[0], block 16, llm/llama.py:130-137):   C[m, n] (+)= sum_k X[m, k] * W[n, k]
//
// W [N, K] is packed in ROW blocks: `bs` (16 or 32) consecutive output rows of one input column k share the two
// exponents e_in / e_out [N/bs, K]; every element has an inlier-width slot (2 or 4 bits, row-major [N, K]) and a
// bitmap bit; an outlier's code is wider than its slot and its high bits live in the segment-compacted `hi`
// stream (DESIGN.md "packed layout").  The scale of a weight changes with k inside a row, so the hardware block
// scaling (one scale per 32 k) cannot apply: the producers dequantize to bf16 (code * 2^e is exact in bf16) and
// feed tcgen05.mma kind::f16, like msq_gemm_w4.cu.
//
// Random access into the compacted hi stream uses a rank index built once per weight from the bitmap
// (msq_packed_build_hi_index: the stream offset of every 8th block): a group of `bs` lanes owns the bs x 8 patch
// (one row block, 8 consecutive k), whose outliers occupy ONE contiguous range of the stream -- the offset of an
// element inside it is a prefix over 8 column ballots plus a popcount.
//   warp 0      : TMA loads the X tile [BNX x 64] (128B swizzle)
//   warps 8..15 : dequantize the W tile [128 x 64] into the 128B-swizzled K-major layout; fence.proxy.async; arrive
//   warp 1      : tcgen05.mma.kind::f16 128 x BNX x 16, fp32 accumulators in TMEM (double buffered)
//   warps 4..7  : epilogue, TMEM -> registers -> C[m, n]
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_tc.cuh"

namespace msq {

constexpr int WM = 128, WBK = 64, WSTAGES = 4, WMAXN = 256;
constexpr int WA_BYTES = WM * WBK * 2, WB_BYTES = WMAXN * WBK * 2;
constexpr int RB_SMEM = WSTAGES * (WA_BYTES + WB_BYTES) + 1024 + 256 + (16 + 256) * 4;

struct RBParams {
  const uint8_t* codes; const uint32_t* bitmap; const int8_t* e_in; const int8_t* e_out;
  float* C; long long ldc;
  int M, N, K;              // C is [M, N]; W is [N, K]; X is [M, K]
  int bnx;                  // activation tile (UMMA N), multiple of 16
  int tiles_n, tiles_m;
  const uint8_t* hi; const uint32_t* hi_off8; long long hi_seg_stride;
  int bs;                   // rows per block: 16 or 32
  int wi, hi_w;             // slot width (2 or 4), extra outlier bits kept in the hi stream (0, 2 or 4)
  int in_ebits, in_bits, out_ebits, out_bits;
  int accumulate;           // C += instead of C =
  int ksplit;               // K is cut into ksplit slices (atomicAdd epilogue when > 1)
  int kb_per_split;         // 64-wide K blocks per slice
};

__global__ void __launch_bounds__(512, 1) k_gemm_rb(const __grid_constant__ CUtensorMap tmX, const RBParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // W tiles
  uint8_t* sB = smem + WSTAGES * WA_BYTES;              // X tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WSTAGES * (WA_BYTES + WB_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + WSTAGES;
  uint64_t* acc_full = bars + 2 * WSTAGES;
  uint64_t* acc_empty = bars + 2 * WSTAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WSTAGES + 4);
  float* lut_in = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);      // [16]  slot -> value
  float* lut_out = lut_in + 16;                                                           // [256] outlier code -> value

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = P.tiles_n * P.tiles_m * P.ksplit;
  const int kblocks_total = (P.K + WBK - 1) / WBK;

  if (warp == 0 && lane == 0) tc::tma_prefetch_desc(&tmX);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WSTAGES; ++i) { tc::mbar_init(&full[i], 1 + 8); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  for (int c = threadIdx.x; c < 256; c += blockDim.x) {
    if (c < 16) lut_in[c] = c < (1 << P.wi) ? decode_elem((uint32_t)c, P.in_ebits, P.in_bits) : 0.0f;
    lut_out[c] = c < (1 << (P.wi + P.hi_w)) ? decode_elem((uint32_t)c, P.out_ebits, P.out_bits) : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile index -> (n tile, m tile, first K block, number of K blocks)
  auto tile_coords = [&](int t, int& tn, int& tm, int& kb0, int& nkb) {
    const int ks = t % P.ksplit;
    const int r = t / P.ksplit;
    tm = r / P.tiles_n; tn = r - tm * P.tiles_n;
    kb0 = ks * P.kb_per_split;
    nkb = min(P.kb_per_split, kblocks_total - kb0);
  };

  // role loops on the whole warp, TMA / tcgen05 instructions under tc::elect_one() (see msq_tc.cuh)
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tn, tm, kb0, nkb; tile_coords(t, tn, tm, kb0, nkb);
      for (int kb = kb0; kb < kb0 + nkb; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&full[stage], (uint32_t)(P.bnx * WBK * 2));
          tc::tma_load_2d(sB + stage * WB_BYTES, &tmX, &full[stage], kb * WBK, tm * P.bnx);
        }
        __syncwarp();
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = tc::make_idesc(1 /*bf16*/, WM, P.bnx);
    const uint64_t ad0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sA)), bd0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sB));
    int stage = 0; uint32_t phase = 0;
    int buf = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tn, tm, kb0, nkb; tile_coords(t, tn, tm, kb0, nkb);
      tc::mbar_wait(&acc_empty[buf], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(buf * WMAXN);
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)stage * (uint32_t)(WA_BYTES >> 4));
        const uint64_t bd = bd0 + (uint64_t)((uint32_t)stage * (uint32_t)(WB_BYTES >> 4));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < WBK / 16; ++k)
            tc::umma_f16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          tc::umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&acc_full[buf]);
      __syncwarp();
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 8) {
    // ---- weight dequant producers: 256 threads = 256 / bs lane groups; a group owns bs rows x 8 k per task
    const int ptid = threadIdx.x - 256;
    const int bs = P.bs, wi = P.wi, hi_w = P.hi_w;
    const int li = ptid & (bs - 1);                       // row inside the block
    const int gid = ptid / bs, ngrp = 256 / bs;
    const unsigned gmask = bs == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
    const int gshift = bs == 32 ? 0 : (lane & 16);
    const uint32_t smask = (1u << wi) - 1u;
    const int tasks = (WM / bs) * 8;                      // per stage
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tn, tm, kb0, nkb; tile_coords(t, tn, tm, kb0, nkb);
      const int n0 = tn * WM;
      for (int kb = kb0; kb < kb0 + nkb; ++kb) {
        // 1. raw loads of every task of this stage (independent of each other: all in flight together)
        uint32_t slots[4], mb[4], off8[4];
        uint2 ei[4], eo[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int tt = gid + ngrp * it;
          slots[it] = mb[it] = off8[it] = 0; ei[it] = eo[it] = make_uint2(0, 0);
          if (tt < tasks) {
            const int kc = tt & 7, rb = tt >> 3;
            const int n = n0 + rb * bs + li, k0 = kb * WBK + kc * 8;
            if (n < P.N && k0 < P.K) {
              const long long e = (long long)n * P.K + k0;
              const uint8_t* cp = P.codes + ((e * wi) >> 3);
              slots[it] = wi == 4 ? __ldg(reinterpret_cast<const uint32_t*>(cp)) : (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(cp));
              mb[it] = (__ldg(P.bitmap + (e >> 5)) >> (k0 & 31)) & 0xFFu;
              const long long tb = (long long)(n / bs) * P.K + k0;          // block-linear index of (row block, k0)
              ei[it] = __ldg(reinterpret_cast<const uint2*>(P.e_in + tb));
              eo[it] = __ldg(reinterpret_cast<const uint2*>(P.e_out + tb));
              if (hi_w) off8[it] = __ldg(P.hi_off8 + (tb >> 3));
            }
          }
        }
        tc::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* dst = sA + stage * WA_BYTES;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int tt = gid + ngrp * it;
          if (tt >= tasks) continue;                      // group-uniform
          const int kc = tt & 7, rb = tt >> 3;
          const int r = rb * bs + li;
          const int n = n0 + r, k0 = kb * WBK + kc * 8;
          const uint32_t m = mb[it];
          // 2. stream offsets of this lane's outliers: prefix over the 8 columns + rank inside the column
          uint32_t hib = 0;                               // 8 x hi_w bits, one field per k
          if (hi_w) {                                     // uniform
            const long long tb = (long long)(n / bs) * P.K + k0;
            const uint8_t* hbase = P.hi + (tb >> 10) * P.hi_seg_stride;     // MSQ_SEG_BLOCKS = 1024 blocks per segment
            uint32_t run = off8[it];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const unsigned bal = (__ballot_sync(0xffffffffu, (m >> j) & 1u) & gmask) >> gshift;
              if ((m >> j) & 1u) {
                const uint32_t e = run + __popc(bal & ((1u << li) - 1u));
                const uint32_t bitpos = e * hi_w;
                const uint32_t v = (__ldg(hbase + (bitpos >> 3)) >> (bitpos & 7)) & ((1u << hi_w) - 1u);
                hib |= v << (hi_w * j);
              }
              run += __popc(bal);
            }
          }
          // 3. dequantize the 8 elements: value = lut[code] * 2^e, exact in bf16
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t slot = (slots[it] >> (wi * j)) & smask;
            const bool isout = (m >> j) & 1u;
            const uint32_t hv = hi_w ? ((hib >> (hi_w * j)) & ((1u << hi_w) - 1u)) : 0u;
            const float q = isout ? lut_out[(hv << wi) | slot] : lut_in[slot];
            // inlier: q * 2^e_in ; outlier: (q * 2^e_out) * 2^-e_in  (utils/quant.py:136: the outlier was scaled by 2^e_in first)
            const int e_i = (int)(signed char)(((j < 4 ? ei[it].x : ei[it].y) >> (8 * (j & 3))) & 0xFFu);
            const int e_o = (int)(signed char)(((j < 4 ? eo[it].x : eo[it].y) >> (8 * (j & 3))) & 0xFFu);
            const bool nan_blk = e_i == MSQ_EXP_NAN || e_o == MSQ_EXP_NAN;
            const int e = isout ? max(-126, min(127, e_o - e_i)) : e_i;
            const float sc = nan_blk ? qnan() : pow2i(e);
            const uint32_t bits = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(q * sc));
            if (j & 1) w[j >> 1] |= bits << 16; else w[j >> 1] = bits;
          }
          const bool valid = n < P.N && k0 < P.K;
          *reinterpret_cast<uint4*>(dst + r * 128 + ((kc ^ (r & 7)) << 4)) =
              valid ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(0, 0, 0, 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&full[stage]);      // one arrival per producer warp
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int buf = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tn, tm, kb0, nkb; tile_coords(t, tn, tm, kb0, nkb);
      tc::mbar_wait(&acc_full[buf], acc_phase);
      tc::tc_fence_after();
      const int n = tn * WM + q * 32 + lane;
      for (int c = 0; c < P.bnx / 16; ++c) {
        uint32_t r[32];
        // 16 columns per step keeps the tail exact for BNX that are multiples of 16
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * WMAXN + c * 16))
            : "memory");
        tc::tmem_ld_wait();
        if (n < P.N) {
#pragma unroll
          for (int v = 0; v < 16; ++v) {
            const int m = tm * P.bnx + c * 16 + v;
            if (m < P.M) {
              float* p = P.C + (long long)m * P.ldc + n;
              const float val = __uint_as_float(r[v]);
              if (P.ksplit > 1) atomicAdd(p, val);
              else *p = P.accumulate ? (*p + val) : val;
            }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[buf]);
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}


}  // namespace msq

using namespace msq;

namespace {

// stream offset (in entries, relative to its segment) of the first outlier of every 8th block
__global__ void __launch_bounds__(256) k_build_hi_index(const uint32_t* __restrict__ bitmap, long long pre, long long L,
                                                        long long post, int bs, long long nb, uint32_t* __restrict__ off8) {
  __shared__ uint32_t wsum[8];
  const long long nblk = (L + bs - 1) / bs;
  const long long t0 = (long long)blockIdx.x * MSQ_SEG_BLOCKS + threadIdx.x * 4;       // 4 consecutive blocks per thread
  uint32_t c[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long t = t0 + u;
    c[u] = 0;
    if (t < nb) {
      const long long j = t % post, pb = t / post, b = pb % nblk, p = pb / nblk;
      for (int i = 0; i < bs; ++i) {
        const long long l = b * bs + i;
        if (l >= L) break;
        const long long e = (p * L + l) * post + j;
        c[u] += (__ldg(bitmap + (e >> 5)) >> (e & 31)) & 1u;
      }
    }
  }
  const uint32_t mine = c[0] + c[1] + c[2] + c[3];
  uint32_t incl = mine;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; ++w) base += wsum[w];
  uint32_t ex = base + incl - mine;                      // exclusive prefix of this thread's first block
  // threads own blocks t0 .. t0+3: every 8th block starts at an even thread
  if ((threadIdx.x & 1) == 0 && t0 < nb) off8[t0 >> 3] = ex;
}

}  // namespace

extern "C" int msq_packed_hi_width(int in_ebits, int in_mbits, int out_ebits, int out_mbits);

extern "C" {

// hi_off8[t / 8] = number of hi-stream entries that precede block t (t a multiple of 8) inside its segment of
// MSQ_SEG_BLOCKS blocks: the rank index that gives the packed-weight GEMM random access into the compacted stream.
// bitmap / geometry as in msq_quantize_mx_outlier_f32; hi_off8 holds ceil(nblocks / 8) uint32.
int msq_packed_build_hi_index(const unsigned int* bitmap, long long pre, long long L, long long post, int block_size,
                              unsigned int* hi_off8, void* stream) {
  msq::DevGuard dev_guard_(bitmap);
  if (!bitmap || !hi_off8 || pre <= 0 || L <= 0 || post <= 0 || block_size <= 0) return MSQ_ERR_BAD_ARG;
  const long long nblk = (L + block_size - 1) / block_size, nb = pre * nblk * post;
  const long long nseg = (nb + MSQ_SEG_BLOCKS - 1) / MSQ_SEG_BLOCKS;
  k_build_hi_index<<<(unsigned)nseg, 256, 0, (cudaStream_t)stream>>>(bitmap, pre, L, post, block_size, nb, hi_off8);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

// C[M,N] fp32 (+)= X[M,K] (bf16, K contiguous, ld = ldx) * W[N,K]^T with W packed in row blocks of block_size (16 or 32)
// output rows (axes=[0] of utils/quant.py:147-266 / what GPTQ.fasterquant emits), hi stream included.
// replaces the fake-quant + F.linear of number_system/mx/linear.py:78-91 for the quantizer's own artefact.
// Needs N % block_size == 0 and K % 64 == 0 (else MSQ_ERR_UNSUPPORTED: dequantize + msq_gemm_nt_f32out).
int msq_gemm_packed_rowblock_bf16(const void* X, long long ldx, const unsigned char* codes, const unsigned int* bitmap,
                                  const signed char* e_in, const signed char* e_out, const unsigned char* hi,
                                  const unsigned int* hi_off8, long long hi_seg_stride, float* C, long long ldc, int M, int N,
                                  int K, int block_size, int in_ebits, int in_mbits, int out_ebits, int out_mbits,
                                  int accumulate, void* stream) {
  msq::DevGuard dev_guard_(X);
  if (!X || !codes || !bitmap || !e_in || !e_out || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if ((block_size != 16 && block_size != 32) || (N % block_size) || (K % 64)) return MSQ_ERR_UNSUPPORTED;
  Fmt fi = make_fmt(in_ebits, in_mbits, 0, 1.0f), fo = make_fmt(out_ebits, out_mbits, 0, 1.0f);
  const int wi = fi.width, hi_w = msq_packed_hi_width(in_ebits, in_mbits, out_ebits, out_mbits);
  if ((wi != 2 && wi != 4) || (hi_w != 0 && hi_w != 2 && hi_w != 4) || wi + hi_w > 8) return MSQ_ERR_UNSUPPORTED;
  if (hi_w && (!hi || !hi_off8)) return MSQ_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(X) & 15) || (ldx % 8) || (reinterpret_cast<uintptr_t>(codes) & 15) ||
      (reinterpret_cast<uintptr_t>(e_in) & 7) || (reinterpret_cast<uintptr_t>(e_out) & 7))
    return MSQ_ERR_ALIGN;
  RBParams P;
  memset(&P, 0, sizeof(P));
  P.codes = codes; P.bitmap = bitmap; P.e_in = e_in; P.e_out = e_out; P.hi = hi; P.hi_off8 = hi_off8;
  P.hi_seg_stride = hi_seg_stride; P.C = C; P.ldc = ldc; P.M = M; P.N = N; P.K = K; P.bs = block_size;
  int bnx = M >= WMAXN ? WMAXN : ((M + 15) / 16) * 16;
  P.bnx = bnx;
  P.tiles_n = (N + WM - 1) / WM; P.tiles_m = (M + bnx - 1) / bnx;
  P.wi = wi; P.hi_w = hi_w; P.in_ebits = in_ebits; P.in_bits = in_mbits; P.out_ebits = out_ebits; P.out_bits = out_mbits;
  P.accumulate = accumulate;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int kblocks_total = (K + WBK - 1) / WBK;
  int ksplit = 1;
  {
    const int tiles = P.tiles_n * P.tiles_m;
    double best = 1e30;
    for (int ks = 1; ks <= (tiles < sms ? 16 : 1); ks *= 2) {
      const int per = (kblocks_total + ks - 1) / ks;
      if (ks > 1 && (per < 4 || (ks - 1) * per >= kblocks_total)) break;
      const int rounds = (tiles * ks + sms - 1) / sms;
      const double cost = (double)rounds * (per + 4.0) + (ks > 1 ? 2.0 : 0.0);
      if (cost < best - 1e-9) { best = cost; ksplit = ks; }
    }
  }
  P.ksplit = ksplit;
  P.kb_per_split = (kblocks_total + ksplit - 1) / ksplit;
  if (ksplit > 1 && !accumulate) {
    if (ldc == N) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), (cudaStream_t)stream);
    else cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, (cudaStream_t)stream);
  }
  CUtensorMap tmX;
  int rc = make_tmap_2d(&tmX, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, ldx, bnx, WBK);
  if (rc) return rc;
  if (cudaFuncSetAttribute(k_gemm_rb, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
  int grid = P.tiles_n * P.tiles_m * P.ksplit;
  if (grid > sms) grid = sms;
  k_gemm_rb<<<grid, 512, RB_SMEM, (cudaStream_t)stream>>>(tmX, P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"
