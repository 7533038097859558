// msq_gemm_mx8.cu -- packed-weight linear on Blackwell's native block-scaled tensor-core path:
//     C[m, n] (+)= sum_k  dq(X)[m, k] * dq(W)[n, k]
// with  tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale  (E4M3 x E4M3 elements, one UE8M0 scale per
// 32 elements along K for both operands, fp32 accumulation in TMEM).
//
//   X : plain MX-FP8 (E4M3) activations, block 32 along K: codes [M, K] bytes + int8 exponents [M, K/32]
//       (what pack_mx(x, 8, "fp8_e4m3", axes=[-1], block_size=32) produces) -- TMA-loaded as bytes.
//   W : packed MicroScopiQ weight [N, K], FP4 (E2M1) inliers AND outliers, block 32 along K: 4-bit codes,
//       outlier bitmap, e_in / e_out.  Producer warps widen the 4-bit codes to E4M3 bytes (every E2M1 value
//       is an E4M3 value) straight into the 128B-swizzled K-major tile; the block scale 2^e_in goes to the
//       scale-factor operand.  An outlier's value is q * 2^(e_out - e_in) (utils/quant.py:136), i.e.
//       q * 2^(e_out - 2 e_in) relative to the block scale the hardware applies, so the producer stores
//       that product as the element's E4M3 value -- exact while -6 <= e_out - 2 e_in <= 6, which the
//       host checks before choosing this path (otherwise the bf16 kernel of msq_gemm_w4.cu is used).
//
// Scale factors: per stage (128 K = 4 blocks) the 128 x 4 UE8M0 bytes of each operand form one 512-byte
// chunk laid out  byte(r, kb) = (r % 32) * 16 + (r / 32) * 4 + kb,  copied shared -> tensor memory with
// tcgen05.cp.32x128b.warpx4 (one 32-bit TMEM column per 32 rows, broadcast to the four lane quarters);
// the MMA for K-block kb selects byte kb through the a_sf_id / b_sf_id fields of the instruction
// descriptor.  The weight tile is the UMMA M operand (SFA = weight scales), the activations the N operand.
#include <stdlib.h>
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_tc.cuh"

namespace msq {

constexpr int XM = 128, XBK = 128, XSTAGES = 4, XMAXST = 8, XMAXN = 256;   // XSTAGES: ring at full activation tiles; XMAXST: at small ones
constexpr int XA_BYTES = XM * XBK, XB_BYTES = XMAXN * XBK;          // 1 byte per element
constexpr int XSF_BYTES = 512 + 1024;                                // SFA chunk + two SFB chunks
constexpr int XRAW_BYTES = XM * XBK / 2;                             // packed FP4 codes of one (128-row, 128-k) chunk
constexpr int MX8_SMEM = 225 * 1024;   // one CTA per SM: XSTAGES full stages, or XMAXST small ones + the staged activation scales
static_assert(XSTAGES * (XA_BYTES + XB_BYTES + XSF_BYTES) + 1024 + 256 <= MX8_SMEM, "ring");
static_assert(XMAXST * (XA_BYTES + 4096 + XSF_BYTES) + 1024 + 256 <= MX8_SMEM, "ring");
constexpr uint32_t TM_ACC = 0, TM_SFA = 256, TM_SFB = 260;           // TMEM columns

struct Mx8Params {
  const uint8_t* wcodes; const uint32_t* wbitmap; const int8_t* we_in; const int8_t* we_out;
  const int8_t* xe;           // activation shared exponents [M, K/32]
  float* C; long long ldc;
  int M, N, K;
  int bnx, tiles_n, tiles_m, ksplit, kb_per_split;
  int plain;                  // weight has no outliers
  int accumulate;
  int wdense;                 // 1: the weight tile comes pre-widened (E4M3 bytes) through TMA, scales from wsf
  const uint8_t* wsf;         // dense mode: UE8M0 weight scales [N, K/32]
  // FP4-by-TMA mode: the packed 4-bit codes are the A operand as they are (TMA type 16U4_ALIGN16B expands 16 codes
  // into a 16-byte chunk, the layout kind::mxf8f6f4 reads E2M1 from); scales = we_in; no thread touches a weight.
  // Outliers (value q * 2^(e_out - e_in) instead of the q * 2^e_in the hardware computes) are a sparse CSR
  // correction applied by the epilogue warps:  C[m, n] += sum_e x[m, o_k[e]] * o_dw[e]  for e in row n.
  int nst;                    // ring depth (XSTAGES, or up to XMAXST when the activation tile is small: decode sizes are latency-bound)
  int xb_bytes;               // bytes of one activation stage
  int sfb_pre;                // tiled mode, one tile per CTA: the activation scale chunks of the CTA's whole K slice are staged
                              // once before the main loop (the per-stage generic-proxy store + proxy fence -- which also waits
                              // for the prefetch loads in flight: ~1000 cycles per stage -- leaves the pipeline)
  int wtiled;                 // FP4-by-TMA with the weight re-laid out tile by tile (msq_tile_fp4_weight): every stage is ONE
  const uint8_t* wsf_tiled;   // contiguous 8 KB code chunk + ONE 512-byte scale chunk already in the tcgen05.cp layout
  int wfp4;
  int fold_m;                 // > 0: the M rows are fold slabs of fold_m tokens each (stacked activation terms of a 16-bit
                              // operand, msq_split_a16_e4m3): row m is ADDED into output row m % fold_m (atomics, C zero-filled)
  int pdl;                    // launched as a programmatic dependent of the zero-fill kernel: C may be touched only after
                              // griddepcontrol.wait (the main loop overlaps the zero-fill and the launch gap behind it)
  int w_fmt;                  // tensor-core format code of the weight operand (0 = E4M3, 5 = E2M1)
  int w_tx_bytes;             // bytes the weight TMA load reports to the mbarrier
  const int* o_ptr; const int* o_k; const float* o_dw;      // CSR over output features (NULL: no outliers)
  const uint8_t* xcodes; long long ldx;                     // activation codes as plain memory (for the correction)
};

__device__ __forceinline__ uint64_t make_sf_desc(uint32_t smem_addr) {
  // no swizzle, 8-row x 16-byte core matrices 128 bytes apart (SBO), single 16-byte column (LBO unused)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void utccp_32x128b(uint32_t tmem_addr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_addr), "l"(desc) : "memory");
}
__device__ __forceinline__ void umma_mxf8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}
// block-scaled instruction descriptor: E4M3 x E4M3 -> F32, UE8M0 scales, both K-major
//   [4,6) b_sf_id  [7,10) A fmt  [10,13) B fmt  [17,23) N>>3  [23] scale fmt (1 = E8M0)  [24,29) M>>4  [29,31) a_sf_id
__device__ __forceinline__ uint32_t make_idesc_mx8(int M, int N, int sf_id) {
  return ((uint32_t)sf_id << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24) |
         ((uint32_t)sf_id << 29);
}

__global__ void __launch_bounds__(512, 1) k_gemm_mx8(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                      const Mx8Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                      // W tiles (E4M3 bytes)
  const int NST = P.nst, XBB = P.xb_bytes;
  uint8_t* sB = smem + NST * XA_BYTES;                     // X tiles
  uint8_t* sSF = smem + NST * (XA_BYTES + XBB);            // per stage: SFA[512] | SFB[1024]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSF + NST * XSF_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + XMAXST;
  uint64_t* acc_full = bars + 2 * XMAXST;
  uint64_t* acc_empty = bars + 2 * XMAXST + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * XMAXST + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = P.tiles_n * P.tiles_m * P.ksplit;
  const int kblocks_total = (P.K + XBK - 1) / XBK;

  uint8_t* presf = reinterpret_cast<uint8_t*>(bars) + 256;   // sfb_pre: one 512-byte chunk per K block of this CTA's slice
  if (warp == 0 && lane == 0) tc::tma_prefetch_desc(&tmX);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&full[i], P.sfb_pre ? 1 : 1 + 8); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(acc_empty, 128);
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  auto tile_coords = [&](int t, int& tn, int& tm, int& kb0, int& nkb) {
    const int ks = t % P.ksplit;
    const int r = t / P.ksplit;
    tm = r / P.tiles_n; tn = r - tm * P.tiles_n;
    kb0 = ks * P.kb_per_split;
    nkb = min(P.kb_per_split, kblocks_total - kb0);
  };
  if (P.sfb_pre && (int)blockIdx.x < num_tiles) {
    // activation scales of this CTA's (single) tile: word (m, kb) = the four UE8M0 bytes of K block kb, in the chunk layout
    int tn, tm, kb0, nkb; tile_coords(blockIdx.x, tn, tm, kb0, nkb);
    const int kblk32 = P.K >> 5;
    for (int i = threadIdx.x; i < nkb * 128; i += blockDim.x) {
      const int kbl = i >> 7, rr = i & 127;
      const int m = tm * P.bnx + rr;
      uint32_t w = 0x7F7F7F7Fu;
      if (rr < P.bnx && m < P.M)
        w = __vadd4(__ldg(reinterpret_cast<const uint32_t*>(P.xe + (long long)m * kblk32) + kb0 + kbl), 0x7F7F7F7Fu);
      *reinterpret_cast<uint32_t*>(presf + kbl * 512 + (rr & 31) * 16 + (rr >> 5) * 4) = w;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);     // warp-uniform for the compiler, too

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tn, tm, kb0, nkb; tile_coords(t, tn, tm, kb0, nkb);
      for (int kb = kb0; kb < kb0 + nkb; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&full[stage], (uint32_t)(P.bnx * XBK + P.w_tx_bytes + (P.wtiled ? 512 : 0)));
          if (P.wtiled) {
            const long long chunk = (long long)tn * kblocks_total + kb;
            tc::tma_load_2d(sA + stage * XA_BYTES, &tmW, &full[stage], 0, (int)(chunk * XM));
            bulk_load_1d(sSF + stage * XSF_BYTES, P.wsf_tiled + chunk * 512, 512, &full[stage]);
          } else if (P.wdense || P.wfp4) tc::tma_load_2d(sA + stage * XA_BYTES, &tmW, &full[stage], kb * XBK, tn * XM);
          tc::tma_load_2d(sB + stage * XBB, &tmX, &full[stage], kb * XBK, tm * P.bnx);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue.  The WHOLE warp runs the loop and only the tcgen05 instructions are predicated on lane 0: everything
    // around them is then warp-uniform, lives in uniform registers and needs no R2UR transfers -- written under
    // `if (lane == 0)` the body of this loop was 278 SASS instructions per stage, which at decode sizes (four 128x16x32
    // MMAs per stage, ~30 cycles of tensor work) WAS the per-stage time of the kernel (~800 cycles, whatever the ring depth,
    // the load path or the number of CTAs).  Descriptors are base + stage * constant (the address field cannot carry).
    int stage = 0; uint32_t phase = 0;
    uint32_t acc_phase = 0;
    const int sfb_chunks = (P.bnx + 127) / 128;
    const uint64_t ad0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sA)), bd0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sB));
    const uint64_t sfd0 = make_sf_desc(tc::smem_u32(sSF)), psd0 = make_sf_desc(tc::smem_u32(presf));
    const uint32_t a_step = (uint32_t)(XA_BYTES >> 4), b_step = (uint32_t)(XBB >> 4), sf_step = (uint32_t)(XSF_BYTES >> 4);
    const uint32_t idesc0 = make_idesc_mx8(XM, P.bnx, 0) | ((uint32_t)P.w_fmt << 7);      // A format: 5 = E2M1
    const uint32_t idesc_k = (1u << 4) | (1u << 29);                                       // next scale-factor id (both operands)
    const uint32_t tacc = tmem_base + TM_ACC, tsfa = tmem_base + TM_SFA, tsfb = tmem_base + TM_SFB;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tn, tm, kb0, nkb; tile_coords(t, tn, tm, kb0, nkb);
      tc::mbar_wait(acc_empty, acc_phase ^ 1);
      tc::tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)stage * a_step), bd = bd0 + (uint64_t)((uint32_t)stage * b_step);
        const uint64_t sfd = sfd0 + (uint64_t)((uint32_t)stage * sf_step);
        if (tc::elect_one()) {
          utccp_32x128b(tsfa, sfd);
          if (P.sfb_pre) utccp_32x128b(tsfb, psd0 + (uint64_t)((uint32_t)kb * 32u));
          else
            for (int c = 0; c < sfb_chunks; ++c) utccp_32x128b(tsfb + 4 * c, sfd + (uint64_t)(32u + 32u * (uint32_t)c));
#pragma unroll
          for (int k = 0; k < XBK / 32; ++k)      // 32 one-byte elements = 32 bytes = 2 descriptor units per UMMA_K
            umma_mxf8(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc0 + (uint32_t)k * idesc_k, (kb | k) != 0, tsfa, tsfb);
          tc::umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(acc_full);
      __syncwarp();
      acc_phase ^= 1;
    }
  } else if (warp >= 8) {
    // ---- producers: 256 threads, per stage 128 rows x 128 K of weights = 1024 chunks of 16 elements
    constexpr int PF = 2;
    const int ptid = threadIdx.x - 256;
    const long long code_row_bytes = (long long)P.K >> 1;
    const int kblk32 = P.K >> 5;
    struct Raw { uint2 c; uint32_t bm; int ei, eo; };
    auto fetch = [&](int n0, int kb, Raw (&raw)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = ptid + 256 * i;
        const int r = c >> 3, kc = c & 7;
        const int n = n0 + r, k = kb * XBK + kc * 16;
        raw[i].c = make_uint2(0, 0); raw[i].bm = 0; raw[i].ei = 0; raw[i].eo = 0;
        if (n < P.N && k < P.K) {
          raw[i].c = __ldg(reinterpret_cast<const uint2*>(P.wcodes + (long long)n * code_row_bytes + (k >> 1)));
          const int blk = n * kblk32 + (k >> 5);
          raw[i].ei = __ldg(P.we_in + blk);
          if (!P.plain) {
            raw[i].bm = __ldg(P.wbitmap + (((long long)n * P.K + k) >> 5));
            raw[i].eo = __ldg(P.we_out + blk);
          }
        }
      }
    };
    // E2M1 magnitude (3 bits) -> E4M3 byte: 0, 0.5, 1, 1.5, 2, 3, 4, 6
    const uint32_t t0 = 0x3C383000u, t1 = 0x4C484440u;     // bytes {00,30,38,3C} {40,44,48,4C}
    const EncC enc8 = make_enc(4, 5);
    auto widen4 = [&](uint32_t c4) {     // 4 nibbles -> 4 E4M3 bytes
      const uint32_t mag = __byte_perm(t0, t1, c4 & 0x7777u);
      const uint32_t s = c4 & 0x8888u;
      return mag | ((s << 4) & 0x80u) | ((s << 8) & 0x8000u) | ((s << 12) & 0x800000u) | ((s << 16) & 0x80000000u);
    };
    auto convert = [&](int stage, const Raw (&raw)[4], int n0, int kb) {
      uint8_t* dst = sA + stage * XA_BYTES;
      uint8_t* sfa = sSF + stage * XSF_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = ptid + 256 * i;
        const int r = c >> 3, kc = c & 7;
        const int k = kb * XBK + kc * 16;
        uint4 outv = make_uint4(0, 0, 0, 0);
        uint32_t sf_byte = 0x7Fu;                    // 2^0 for rows / K-blocks outside the matrix
        uint32_t ob = 0;
        const bool inside = n0 + r < P.N && k < P.K;
        if (inside) {
          sf_byte = (uint32_t)(raw[i].ei + 127) & 0xFFu;    // UE8M0; MSQ_EXP_NAN (-128) maps to 0xFF = NaN
          outv = make_uint4(widen4(raw[i].c.x & 0xFFFFu), widen4(raw[i].c.x >> 16), widen4(raw[i].c.y & 0xFFFFu),
                            widen4(raw[i].c.y >> 16));
          ob = (raw[i].bm >> (k & 31)) & 0xFFFFu;
        }
        uint8_t* piece = dst + r * 128 + ((kc ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(piece) = outv;
        // outliers (rare): q * 2^(e_out - 2 e_in) relative to the block scale, patched byte by byte in shared memory
        // (no register-array indexing, so nothing spills to local memory and the code stays small)
        while (ob) {
          const int e = __ffs(ob) - 1;
          ob &= ob - 1;
          const int ei = raw[i].ei, eo = raw[i].eo;
          const uint32_t code = ((e < 8 ? raw[i].c.x : raw[i].c.y) >> (4 * (e & 7))) & 0xFu;
          const float v = __fmul_rn(decode_elem(code, 2, 3), pow2i(max(-60, min(60, eo - 2 * ei))));
          piece[e] = (uint8_t)((eo == MSQ_EXP_NAN) ? 0x7Fu : (encode_fast(v, enc8) & 0xFFu));
        }
        if ((kc & 1) == 0) sfa[(r & 31) * 16 + (r >> 5) * 4 + (kc >> 1)] = (uint8_t)sf_byte;
      }
    };
    // activation scale chunks: thread t < bnx handles activation row t of the tile (4 K-blocks = one 32-bit word)
    const bool sf_words = (kblk32 & 3) == 0 && (reinterpret_cast<uintptr_t>(P.xe) & 3) == 0;
    auto load_sfb = [&](int m0, int kb) -> uint32_t {
      const int m = m0 + ptid;
      if (ptid >= P.bnx || m >= P.M) return 0x7F7F7F7Fu;
      if (sf_words) return __vadd4(__ldg(reinterpret_cast<const uint32_t*>(P.xe + (long long)m * kblk32) + kb), 0x7F7F7F7Fu);
      uint32_t raw4 = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kbq = kb * 4 + q;
        const uint32_t b = (kbq < kblk32) ? ((uint32_t)(__ldg(P.xe + (long long)m * kblk32 + kbq) + 127) & 0xFFu) : 0x7Fu;
        raw4 |= b << (8 * q);
      }
      return raw4;
    };
    auto store_sfb = [&](int stage, uint32_t v) {
      if (ptid < P.bnx) {
        const int rr = ptid & 127;
        *reinterpret_cast<uint32_t*>(sSF + stage * XSF_BYTES + 512 + (ptid >> 7) * 512 + (rr & 31) * 16 + (rr >> 5) * 4) = v;
      }
    };
    struct Cursor { int t, tn, tm, kb0, nkb, kb; bool ok; };
    auto cur_init = [&](Cursor& c) {
      c.t = blockIdx.x; c.kb = 0; c.ok = c.t < num_tiles;
      if (c.ok) tile_coords(c.t, c.tn, c.tm, c.kb0, c.nkb);
    };
    auto cur_next = [&](Cursor& c) {
      if (++c.kb >= c.nkb) {
        c.t += gridDim.x; c.kb = 0; c.ok = c.t < num_tiles;
        if (c.ok) tile_coords(c.t, c.tn, c.tm, c.kb0, c.nkb);
      }
    };
    Cursor fc, cc;
    cur_init(fc); cur_init(cc);
    int stage = 0; uint32_t phase = 0;
    const bool wtma = P.wdense || P.wfp4;
    if (wtma && !P.sfb_pre) {
      // the weight bytes arrive by TMA; only the two scale-factor chunks are assembled here.  K % 128 == 0 in this
      // mode, so one K-block's four scales are one aligned 32-bit word; the next stage's words are fetched before
      // waiting for the current stage to drain.
      auto ld_sf = [&](const Cursor& c, uint32_t& va, uint32_t& vb) {
        va = 0x7F7F7F7Fu; vb = 0x7F7F7F7Fu;
        if (!c.ok) return;
        const int kb = c.kb0 + c.kb;
        const int n = c.tn * XM + ptid;
        if (ptid < XM && n < P.N && !P.wtiled)
          va = P.wfp4 ? __vadd4(__ldg(reinterpret_cast<const uint32_t*>(P.we_in + (long long)n * kblk32) + kb), 0x7F7F7F7Fu)
                      : __ldg(reinterpret_cast<const uint32_t*>(P.wsf + (long long)n * kblk32) + kb);
        vb = load_sfb(c.tm * P.bnx, kb);
      };
      // scale words of the next SFD stages are kept in registers: their global-load latency is off the per-stage chain
      constexpr int SFD = 4;
      uint32_t ra[SFD], rb[SFD];
#pragma unroll
      for (int j = 0; j < SFD; ++j) {
        ld_sf(fc, ra[j], rb[j]);
        if (j + 1 < SFD) cur_next(fc);
      }
      while (cc.ok) {
#pragma unroll
        for (int j = 0; j < SFD; ++j) {
          if (cc.ok) {
            const uint32_t va = ra[j], vb = rb[j];
            cur_next(fc);
            ld_sf(fc, ra[j], rb[j]);                 // the stage SFD ahead takes this slot
            tc::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sf = sSF + stage * XSF_BYTES;
            if (ptid < XM && !P.wtiled) *reinterpret_cast<uint32_t*>(sf + (ptid & 31) * 16 + (ptid >> 5) * 4) = va;
            store_sfb(stage, vb);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[stage]);
            if (++stage == NST) { stage = 0; phase ^= 1; }
            cur_next(cc);
          }
        }
      }
    }
    Raw ring[PF][4];
    uint32_t sfb_ring[PF];
#pragma unroll
    for (int j = 0; j < PF - 1; ++j) {
      if (!wtma && fc.ok) {
        fetch(fc.tn * XM, fc.kb0 + fc.kb, ring[j]);
        sfb_ring[j] = load_sfb(fc.tm * P.bnx, fc.kb0 + fc.kb);
        cur_next(fc);
      }
    }
    while (!wtma && cc.ok) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        if (cc.ok) {
          if (fc.ok) {
            fetch(fc.tn * XM, fc.kb0 + fc.kb, ring[(j + PF - 1) % PF]);
            sfb_ring[(j + PF - 1) % PF] = load_sfb(fc.tm * P.bnx, fc.kb0 + fc.kb);
            cur_next(fc);
          }
          tc::mbar_wait(&empty[stage], phase ^ 1);
          convert(stage, ring[j], cc.tn * XM, cc.kb0 + cc.kb);
          store_sfb(stage, sfb_ring[j]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&full[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
          cur_next(cc);
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tn, tm, kb0, nkb; tile_coords(t, tn, tm, kb0, nkb);
      const int n = tn * XM + q * 32 + lane;
      // sparse outlier correction: this feature's CSR range is fetched before the accumulator is waited for
      int oe0 = 0, oe1 = 0;
      if (P.o_ptr && n < P.N) { oe0 = __ldg(P.o_ptr + n); oe1 = __ldg(P.o_ptr + n + 1); }
      // decode sizes: the whole correction of this feature (<= 32 tokens) is summed into registers WHILE the main loop
      // runs -- the epilogue warps have nothing else to do until the accumulator is complete -- and added afterwards
      constexpr int CORR_MAX = 32;
      float corr[CORR_MAX];
      // C is accumulated by atomics (split K / folded slabs): the correction does not have to pass through the thread that
      // owns the accumulator row.  The whole warp works on one outlier at a time -- lane = token, one gather per lane
      // instead of one lane gathering every token -- and adds x * dw straight into C, WHILE the main loop runs.
      const bool atomic_corr = P.o_ptr != nullptr && (P.ksplit > 1 || P.fold_m) && !P.accumulate;
      if (atomic_corr) {
        if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");       // C has been zero-filled from here on
        const int klo = kb0 * XBK, khi = (kb0 + nkb) * XBK;
        const int kblk32 = P.K >> 5;
        uint32_t have = __ballot_sync(0xffffffffu, oe1 > oe0);
        while (have) {
          const int L = __ffs(have) - 1;
          have &= have - 1;
          const int b0 = __shfl_sync(0xffffffffu, oe0, L), b1 = __shfl_sync(0xffffffffu, oe1, L);
          float* ccol = P.C + (tn * XM + q * 32 + L);
          for (int eb = b0; eb < b1; eb += 32) {
            // this batch's (k, dw) pairs: one coalesced load, then broadcast by shuffles
            int kk = -1; float dd = 0.0f;
            if (eb + lane < b1) { kk = __ldg(P.o_k + eb + lane); dd = __ldg(P.o_dw + eb + lane); }
            if (kk < klo || kk >= khi) kk = -1;                       // another K slice's outlier
            uint32_t todo = __ballot_sync(0xffffffffu, kk >= 0);
            while (todo) {
              // up to four outliers in flight: all gathers are issued before the first atomic
              int ks[4]; float ds[4]; uint32_t cb[4]; int xb[4];
              int cnt = 0;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                ks[u] = -1; ds[u] = 0.0f;
                if (todo) {
                  const int j = __ffs(todo) - 1;
                  todo &= todo - 1;
                  ks[u] = __shfl_sync(0xffffffffu, kk, j); ds[u] = __shfl_sync(0xffffffffu, dd, j);
                  ++cnt;
                }
              }
              for (int v = lane; v < P.bnx; v += 32) {
                const int m = tm * P.bnx + v;
                const long long mc = min(m, P.M - 1);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int k = max(ks[u], 0);
                  cb[u] = __ldg(P.xcodes + mc * P.ldx + k);
                  xb[u] = __ldg(P.xe + mc * kblk32 + (k >> 5));
                }
                if (m < P.M) {
                  float* p = ccol + (long long)(P.fold_m ? m % P.fold_m : m) * P.ldc;
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    if (u < cnt) {
                      const float xs = xb[u] == MSQ_EXP_NAN ? qnan() : pow2i(xb[u]);
                      atomicAdd(p, __fmul_rn(__fmul_rn(decode_elem(cb[u], 4, 5), xs), ds[u]));
                    }
                  }
                }
              }
            }
          }
        }
        oe1 = oe0;                         // nothing left for the accumulator-side corrections below
      }
      const bool pre_corr = P.bnx <= CORR_MAX && !atomic_corr;
      if (pre_corr) {
#pragma unroll
        for (int v = 0; v < CORR_MAX; ++v) corr[v] = 0.0f;
        const int klo = kb0 * XBK, khi = (kb0 + nkb) * XBK;
        const int kblk32 = P.K >> 5;
        for (int e = oe0; e < oe1; ++e) {
          const int k = __ldg(P.o_k + e);
          if (k < klo || k >= khi) continue;
          const float dw = __ldg(P.o_dw + e);
#pragma unroll
          for (int v = 0; v < CORR_MAX; ++v) {
            const int m = tm * P.bnx + v;
            if (v < P.bnx && m < P.M) {
              const int xe = P.xe[(long long)m * kblk32 + (k >> 5)];
              const float xs = xe == MSQ_EXP_NAN ? qnan() : pow2i(xe);
              const float xv = __fmul_rn(decode_elem((uint32_t)P.xcodes[(long long)m * P.ldx + k], 4, 5), xs);
              corr[v] = fmaf(xv, dw, corr[v]);
            }
          }
        }
        oe1 = oe0;                       // nothing left for the in-loop correction below
      }
      tc::mbar_wait(acc_full, acc_phase);
      tc::tc_fence_after();
      if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int c = 0; c < P.bnx / 16; ++c) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + TM_ACC + (uint32_t)(c * 16))
            : "memory");
        tc::tmem_ld_wait();
        if (n < P.N) {
          if (oe1 > oe0) {    // sparse outlier correction for this (feature, 16 tokens, K slice)
            const int klo = kb0 * XBK, khi = (kb0 + nkb) * XBK;
            const int kblk32 = P.K >> 5;
            for (int e = oe0; e < oe1; ++e) {
              const int k = __ldg(P.o_k + e);
              if (k < klo || k >= khi) continue;
              const float dw = __ldg(P.o_dw + e);
#pragma unroll
              for (int v = 0; v < 16; ++v) {
                const int m = tm * P.bnx + c * 16 + v;
                if (m < P.M) {
                  const int xe = P.xe[(long long)m * kblk32 + (k >> 5)];
                  const float xs = xe == MSQ_EXP_NAN ? qnan() : pow2i(xe);
                  const float xv = __fmul_rn(decode_elem((uint32_t)P.xcodes[(long long)m * P.ldx + k], 4, 5), xs);
                  r[v] = __float_as_uint(fmaf(xv, dw, __uint_as_float(r[v])));
                }
              }
            }
          }
#pragma unroll
          for (int v = 0; v < 16; ++v) {
            const int m = tm * P.bnx + c * 16 + v;
            if (m < P.M) {
              float* p = P.C + (long long)(P.fold_m ? m % P.fold_m : m) * P.ldc + n;
              float val = __uint_as_float(r[v]);
              if (pre_corr) val += (c == 0 ? corr[v] : corr[16 + v]);
              if (P.ksplit > 1 || P.fold_m) atomicAdd(p, val);
              else *p = P.accumulate ? (*p + val) : val;
            }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(acc_empty);
      acc_phase ^= 1;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// zero-fill of C in front of a split-K launch.  It releases its dependent (the GEMM, launched with programmatic stream
// serialization) at once: the GEMM's prologue and main loop run while C is being cleared, its epilogue waits
// (griddepcontrol.wait) for this grid to have completed before the first atomicAdd.
__global__ void __launch_bounds__(256) k_zero_c(float* __restrict__ C, long long ldc, int M, int N) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long n4 = N >> 2;                       // N % 4 == 0 and ldc % 4 == 0 (checked by the launcher)
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < (long long)M * n4; i += (long long)gridDim.x * 256) {
    const long long m = i / n4, c = i - m * n4;
    *reinterpret_cast<float4*>(C + m * ldc + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// FP4 weight [N, K] (codes row-major, K/2 bytes per row; e_in [N, K/32]) -> GEMM-ready tiles: for every (128-row tile tn,
// 128-wide K block kb) one contiguous 8 KB chunk of codes (row r at r * 64 bytes) and one 512-byte scale chunk in the
// tcgen05.cp layout (UE8M0 = e + 127).  A stage of k_gemm_mx8 then reads two contiguous pieces of DRAM instead of 128
// 64-byte row segments 4 KB apart and 128 scattered scale words (measured: 1.3 TB/s of DRAM reads with the row-major planes)
__global__ void __launch_bounds__(256) k_tile_fp4_weight(const uint8_t* __restrict__ codes, const int8_t* __restrict__ e_in,
                                                         uint8_t* __restrict__ codes_t, uint8_t* __restrict__ sf_t, int N, int K,
                                                         int kblocks) {
  const int kb = blockIdx.x, tn = blockIdx.y;
  const long long chunk = (long long)tn * kblocks + kb;
  const int t = threadIdx.x;
  {
    const int r = t >> 1, h = t & 1;                       // row, 32-byte half of its 64 bytes
    const int n = tn * XM + r;
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (n < N) {
      const uint4* src = reinterpret_cast<const uint4*>(codes + (long long)n * (K >> 1) + kb * 64 + h * 32);
      a = __ldg(src); b = __ldg(src + 1);
    }
    uint4* dst = reinterpret_cast<uint4*>(codes_t + chunk * XRAW_BYTES + r * 64 + h * 32);
    dst[0] = a; dst[1] = b;
  }
  if (t < XM) {
    const int n = tn * XM + t;
    uint32_t w = 0x7F7F7F7Fu;
    if (n < N) w = __vadd4(__ldg(reinterpret_cast<const uint32_t*>(e_in + (long long)n * (K >> 5)) + kb), 0x7F7F7F7Fu);
    *reinterpret_cast<uint32_t*>(sf_t + chunk * 512 + (t & 31) * 16 + (t >> 5) * 4) = w;
  }
}

// packed FP4 weight -> dense E4M3 bytes [N, K] + UE8M0 scales [N, K/32] (outliers folded as in k_gemm_mx8)
__global__ void __launch_bounds__(256) k_widen_fp4_e4m3(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ bitmap,
                                                        const int8_t* __restrict__ e_in, const int8_t* __restrict__ e_out,
                                                        uint8_t* __restrict__ out, uint8_t* __restrict__ sf, long long nchunks,
                                                        int K, int plain) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;      // one chunk = 16 consecutive K elements
  if (c >= nchunks) return;
  const long long e0 = c * 16;
  const long long blk = e0 >> 5;
  const int k = (int)(e0 % K);
  const uint2 cw = __ldg(reinterpret_cast<const uint2*>(codes + (e0 >> 1)));
  const int ei = e_in[blk], eo = plain ? 0 : e_out[blk];
  const uint32_t t0 = 0x3C383000u, t1 = 0x4C484440u;
  auto widen4 = [&](uint32_t c4) {
    const uint32_t mag = __byte_perm(t0, t1, c4 & 0x7777u);
    const uint32_t s = c4 & 0x8888u;
    return mag | ((s << 4) & 0x80u) | ((s << 8) & 0x8000u) | ((s << 12) & 0x800000u) | ((s << 16) & 0x80000000u);
  };
  uint32_t w[4] = {widen4(cw.x & 0xFFFFu), widen4(cw.x >> 16), widen4(cw.y & 0xFFFFu), widen4(cw.y >> 16)};
  *reinterpret_cast<uint4*>(out + e0) = make_uint4(w[0], w[1], w[2], w[3]);
  if (!plain) {
    uint32_t ob = (__ldg(bitmap + (e0 >> 5)) >> (k & 31)) & 0xFFFFu;
    const EncC enc8 = make_enc(4, 5);
    while (ob) {                                   // rare: patch the byte in place (same thread, program order)
      const int e = __ffs(ob) - 1;
      ob &= ob - 1;
      const uint32_t code = ((e < 8 ? cw.x : cw.y) >> (4 * (e & 7))) & 0xFu;
      const float v = __fmul_rn(decode_elem(code, 2, 3), pow2i(max(-60, min(60, eo - 2 * ei))));
      out[e0 + e] = (uint8_t)((eo == MSQ_EXP_NAN) ? 0x7Fu : (encode_fast(v, enc8) & 0xFFu));
    }
  }
  if ((e0 & 31) == 0) sf[blk] = (uint8_t)((ei + 127) & 0xFF);
}

}  // namespace msq

using namespace msq;

extern "C" {

// C[M,N] fp32 (+)= dq(X)[M,K] * dq(W)[N,K]^T on the block-scaled tensor-core path.
//   xcodes [M, K] E4M3 bytes (ld = ldx bytes), xe int8 [M, K/32]         (plain MX-FP8 activations)
//   wcodes / wbitmap / we_in / we_out: packed FP4 weight, block 32 along K (plain != 0: no outliers)
// The caller guarantees that every outlier block satisfies -6 <= e_out - 2 e_in <= 6 (exact E4M3 folding).
struct Mx4Sparse { const int* o_ptr; const int* o_k; const float* o_dw; const unsigned char* codes_tiled; const unsigned char* sf_tiled; };

static int launch_mx8(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* wcodes,
                      const unsigned int* wbitmap, const signed char* we_in, const signed char* we_out,
                      const unsigned char* wdense, long long ldw, const unsigned char* wsf, float* C, long long ldc, int M,
                      int N, int K, int plain, int accumulate, cudaStream_t s, const Mx4Sparse* fp4 = nullptr, int fold_m = 0) {
  Mx8Params P;
  memset(&P, 0, sizeof(P));
  P.w_tx_bytes = wdense ? XA_BYTES : 0;
  // a 16U4_ALIGN16B load reports the bytes it READ (8 per 16-byte chunk written) to the mbarrier [measured: expecting
  // the 16 KB it writes never completes; 8 KB does]
  if (fp4) {
    P.w_tx_bytes = XA_BYTES / 2; P.wfp4 = 1; P.w_fmt = 5; P.o_ptr = fp4->o_ptr; P.o_k = fp4->o_k; P.o_dw = fp4->o_dw; P.xcodes = xcodes; P.ldx = ldx;
    if (fp4->codes_tiled) { P.wtiled = 1; P.wsf_tiled = fp4->sf_tiled; }
  }
  P.wcodes = wcodes; P.wbitmap = wbitmap; P.we_in = we_in; P.we_out = we_out; P.xe = xe; P.C = C; P.ldc = ldc;
  P.M = M; P.N = N; P.K = K; P.plain = plain; P.accumulate = accumulate;
  P.wdense = wdense != nullptr; P.wsf = wsf;
  const int bnx = M >= XMAXN ? XMAXN : ((M + 15) / 16) * 16;
  P.bnx = bnx;
  P.tiles_n = (N + XM - 1) / XM; P.tiles_m = (M + bnx - 1) / bnx;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int kblocks_total = (K + XBK - 1) / XBK;
  // split K so that the persistent grid stays busy: cost = rounds over the SMs x (K-blocks per item + a fixed
  // per-item charge for the TMEM drain and the atomics)
  int ksplit = 1;
  {
    const int tiles = P.tiles_n * P.tiles_m;
    double best = 1e30;
    // only when the un-split grid cannot fill the machine: split-K pays a memset and fp32 atomics on the whole output
    for (int ks = 1; ks <= (tiles < sms ? 16 : 1); ks *= 2) {
      const int per = (kblocks_total + ks - 1) / ks;
      if (ks > 1 && (per < 2 || (ks - 1) * per >= kblocks_total)) break;      // every split must own K-blocks
      const int rounds = (tiles * ks + sms - 1) / sms;
      const double cost = (double)rounds * (per + 2.0) + (ks > 1 ? 1.0 : 0.0);
      if (cost < best - 1e-9) { best = cost; ksplit = ks; }
    }
  }
  P.ksplit = ksplit;
  P.kb_per_split = (kblocks_total + ksplit - 1) / ksplit;
  // ring: at decode sizes a stage is one 8-16 KB weight tile and a 2-4 KB activation tile, and the kernel is bound by the
  // latency of the per-stage hand-offs: twice the depth keeps twice the bytes in flight
  P.xb_bytes = bnx <= 32 ? ((bnx * XBK + 1023) / 1024) * 1024 : XB_BYTES;
  P.nst = bnx <= 32 ? XMAXST : XSTAGES;
  {
    const bool one_tile = P.tiles_n * P.tiles_m * ksplit <= sms;
    const long long ring = (long long)P.nst * (XA_BYTES + P.xb_bytes + XSF_BYTES);
    P.sfb_pre = (P.wtiled && bnx <= 128 && one_tile && (long long)P.kb_per_split * 512 <= (long long)MX8_SMEM - 1024 - 256 - ring) ? 1 : 0;
  }
  P.fold_m = fold_m;
  const int M_out = fold_m ? fold_m : M;
  if ((ksplit > 1 || fold_m) && !accumulate) {
    const bool no_pdl = getenv("MSQ_MX8_NO_PDL") != nullptr;
    if (!no_pdl && (N % 4) == 0 && (ldc % 4) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
      const long long n4 = (long long)M_out * (N >> 2);
      int zb = (int)((n4 + 255) / 256);
      if (zb > 4 * sms) zb = 4 * sms;
      k_zero_c<<<zb, 256, 0, s>>>(C, ldc, M_out, N);
      P.pdl = 1;
    } else if (ldc == N) cudaMemsetAsync(C, 0, (size_t)M_out * N * sizeof(float), s);
    else cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M_out, s);
  }
  CUtensorMap tmX, tmW;
  int rc = make_tmap_2d(&tmX, xcodes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, K, ldx, bnx, XBK);
  if (rc) return rc;
  if (wdense) {
    rc = make_tmap_2d(&tmW, wdense, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, ldw, XM, XBK);
    if (rc) return rc;
  } else if (fp4 && fp4->codes_tiled) {
    rc = make_tmap_2d_u4(&tmW, fp4->codes_tiled, (long long)P.tiles_n * kblocks_total * XM, XBK, XBK / 2, XM);
    if (rc) return rc;
  } else if (fp4) {
    rc = make_tmap_2d_u4(&tmW, wcodes, N, K, (long long)K / 2, XM);
    if (rc) return rc;
  } else {
    tmW = tmX;
  }
  // the opt-in is per device: set it on every launch (microseconds) instead of once per process
  if (cudaFuncSetAttribute(k_gemm_mx8, cudaFuncAttributeMaxDynamicSharedMemorySize, MX8_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
  int grid = P.tiles_n * P.tiles_m * P.ksplit;
  if (grid > sms) grid = sms;
  if (P.pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = MX8_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_gemm_mx8, tmX, tmW, P) == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
  }
  k_gemm_mx8<<<grid, 512, MX8_SMEM, s>>>(tmX, tmW, P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_gemm_mxfp8_blockscaled(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* wcodes,
                               const unsigned int* wbitmap, const signed char* we_in, const signed char* we_out, float* C,
                               long long ldc, int M, int N, int K, int plain, int accumulate, void* stream) {
  msq::DevGuard dev_guard_(xcodes);
  if (!xcodes || !xe || !wcodes || !we_in || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (!plain && (!wbitmap || !we_out)) return MSQ_ERR_BAD_ARG;
  if (K % 32 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xcodes) & 15) || (ldx % 16) || (reinterpret_cast<uintptr_t>(wcodes) & 15)) return MSQ_ERR_ALIGN;
  return launch_mx8(xcodes, ldx, xe, wcodes, wbitmap, we_in, we_out, nullptr, 0, nullptr, C, ldc, M, N, K, plain, accumulate,
                    (cudaStream_t)stream);
}

// packed FP4 weight -> dense MX-FP8 operand: E4M3 bytes [N, K] + UE8M0 scales [N, K/32]
int msq_widen_fp4_to_mxfp8(const unsigned char* wcodes, const unsigned int* wbitmap, const signed char* we_in,
                           const signed char* we_out, unsigned char* out_bytes, unsigned char* out_sf, int N, int K, int plain,
                           void* stream) {
  msq::DevGuard dev_guard_(wcodes);
  if (!wcodes || !we_in || !out_bytes || !out_sf || N <= 0 || K <= 0 || (K % 32)) return MSQ_ERR_BAD_ARG;
  if (!plain && (!wbitmap || !we_out)) return MSQ_ERR_BAD_ARG;
  const long long nchunks = (long long)N * K / 16;
  k_widen_fp4_e4m3<<<(unsigned)((nchunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wcodes, wbitmap, we_in, we_out, out_bytes,
                                                                                      out_sf, nchunks, K, plain);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

// same contraction with BOTH operands dense MX-FP8 (bytes + UE8M0 scales): a pure TMA-fed block-scaled GEMM
int msq_gemm_mxfp8_dense(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* wbytes,
                         long long ldw, const unsigned char* wsf, float* C, long long ldc, int M, int N, int K, int accumulate,
                         void* stream) {
  msq::DevGuard dev_guard_(xcodes);
  if (!xcodes || !xe || !wbytes || !wsf || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (K % 128 != 0) return MSQ_ERR_UNSUPPORTED;          // scale words are read four K-blocks at a time
  if ((reinterpret_cast<uintptr_t>(xe) & 3) || (reinterpret_cast<uintptr_t>(wsf) & 3)) return MSQ_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(xcodes) & 15) || (ldx % 16) || (reinterpret_cast<uintptr_t>(wbytes) & 15) || (ldw % 16)) return MSQ_ERR_ALIGN;
  return launch_mx8(xcodes, ldx, xe, nullptr, nullptr, nullptr, nullptr, wbytes, ldw, wsf, C, ldc, M, N, K, 1, accumulate,
                    (cudaStream_t)stream);
}

// FP4-by-TMA mode: the packed 4-bit codes of the weight (block 32 along K, E2M1, equal inlier / outlier width) feed
// tcgen05.mma kind::mxf8f6f4 directly -- TMA expands 16 codes into a 16-byte chunk (CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B),
// the block scale is we_in, no thread widens anything -- and the outliers are a sparse correction in the epilogue:
// CSR over output features, o_dw = q * (2^(e_out - e_in) - 2^e_in) (what the true weight differs from q * 2^e_in by).
// o_ptr may be NULL (no outliers).  Needs K % 128 == 0 and wcodes 32-byte aligned.
int msq_gemm_mxfp4w_blockscaled(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* wcodes,
                                const signed char* we_in, const int* o_ptr, const int* o_k, const float* o_dw, float* C,
                                long long ldc, int M, int N, int K, int accumulate, void* stream) {
  msq::DevGuard dev_guard_(xcodes);
  if (!xcodes || !xe || !wcodes || !we_in || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (o_ptr && (!o_k || !o_dw)) return MSQ_ERR_BAD_ARG;
  if (K % 128 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xe) & 3) || (reinterpret_cast<uintptr_t>(we_in) & 3)) return MSQ_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(xcodes) & 15) || (ldx % 16) || (reinterpret_cast<uintptr_t>(wcodes) & 31)) return MSQ_ERR_ALIGN;
  Mx4Sparse sp{o_ptr, o_k, o_dw, nullptr, nullptr};
  return launch_mx8(xcodes, ldx, xe, wcodes, nullptr, we_in, nullptr, nullptr, 0, nullptr, C, ldc, M, N, K, 1, accumulate,
                    (cudaStream_t)stream, &sp);
}

// GEMM-ready layout of a K-blocked FP4 weight (see k_tile_fp4_weight): codes_tiled [ceil(N/128) * (K/128) * 8192] bytes,
// sf_tiled [ceil(N/128) * (K/128) * 512] bytes.  K % 128 == 0; codes 16-byte aligned; outputs 32- / 16-byte aligned.
int msq_tile_fp4_weight(const unsigned char* wcodes, const signed char* we_in, unsigned char* codes_tiled,
                        unsigned char* sf_tiled, int N, int K, void* stream) {
  msq::DevGuard dev_guard_(wcodes);
  if (!wcodes || !we_in || !codes_tiled || !sf_tiled || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (K % 128 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(wcodes) & 15) || (reinterpret_cast<uintptr_t>(we_in) & 3) ||
      (reinterpret_cast<uintptr_t>(codes_tiled) & 31) || (reinterpret_cast<uintptr_t>(sf_tiled) & 15))
    return MSQ_ERR_ALIGN;
  dim3 grid((unsigned)(K / XBK), (unsigned)((N + XM - 1) / XM));
  k_tile_fp4_weight<<<grid, 256, 0, (cudaStream_t)stream>>>(wcodes, we_in, codes_tiled, sf_tiled, N, K, K / XBK);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

// msq_gemm_mxfp4w_blockscaled on the tiled layout: every pipeline stage is one contiguous 8 KB + one 512 B read
int msq_gemm_mxfp4w_tiled_blockscaled(const unsigned char* xcodes, long long ldx, const signed char* xe,
                                      const unsigned char* codes_tiled, const unsigned char* sf_tiled, const int* o_ptr,
                                      const int* o_k, const float* o_dw, float* C, long long ldc, int M, int N, int K,
                                      int accumulate, void* stream) {
  msq::DevGuard dev_guard_(xcodes);
  if (!xcodes || !xe || !codes_tiled || !sf_tiled || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (o_ptr && (!o_k || !o_dw)) return MSQ_ERR_BAD_ARG;
  if (K % 128 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xe) & 3) || (reinterpret_cast<uintptr_t>(xcodes) & 15) || (ldx % 16) ||
      (reinterpret_cast<uintptr_t>(codes_tiled) & 31) || (reinterpret_cast<uintptr_t>(sf_tiled) & 15))
    return MSQ_ERR_ALIGN;
  Mx4Sparse sp{o_ptr, o_k, o_dw, codes_tiled, sf_tiled};
  return launch_mx8(xcodes, ldx, xe, codes_tiled, nullptr, reinterpret_cast<const signed char*>(sf_tiled), nullptr, nullptr, 0,
                    nullptr, C, ldc, M, N, K, 1, accumulate, (cudaStream_t)stream, &sp);
}

// the same GEMM over STACKED activation terms (msq_split_a16_e4m3: M = terms * fold_m rows): the slabs are summed into
// C [fold_m, N] by the epilogue's atomics instead of by a pass over a [M, N] result
int msq_gemm_mxfp4w_tiled_folded(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* codes_tiled,
                                 const unsigned char* sf_tiled, const int* o_ptr, const int* o_k, const float* o_dw, float* C,
                                 long long ldc, int M, int fold_m, int N, int K, void* stream) {
  msq::DevGuard dev_guard_(xcodes);
  if (!xcodes || !xe || !codes_tiled || !sf_tiled || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (fold_m <= 0 || M % fold_m != 0) return MSQ_ERR_BAD_ARG;
  if (o_ptr && (!o_k || !o_dw)) return MSQ_ERR_BAD_ARG;
  if (K % 128 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xe) & 3) || (reinterpret_cast<uintptr_t>(xcodes) & 15) || (ldx % 16) ||
      (reinterpret_cast<uintptr_t>(codes_tiled) & 31) || (reinterpret_cast<uintptr_t>(sf_tiled) & 15))
    return MSQ_ERR_ALIGN;
  Mx4Sparse sp{o_ptr, o_k, o_dw, codes_tiled, sf_tiled};
  return launch_mx8(xcodes, ldx, xe, codes_tiled, nullptr, reinterpret_cast<const signed char*>(sf_tiled), nullptr, nullptr, 0,
                    nullptr, C, ldc, M, N, K, 1, 0, (cudaStream_t)stream, &sp, fold_m == M ? 0 : fold_m);
}

}  // extern "C"
