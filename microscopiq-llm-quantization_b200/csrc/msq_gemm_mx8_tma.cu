// msq_gemm_mx8_tma.cu -- the compute-bound (large-M) shape of the block-scaled linear
//     C[m, n] (+)= sum_k dq(X)[m, k] * dq(W)[n, k]          tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale
// in which NO thread touches an operand: the element tiles arrive by tensor loads, the UE8M0 scale factors of BOTH operands
// arrive by linear bulk copies from copies that are already in the tcgen05.cp chunk layout (msq_tile_scales_ue8m0: the
// weight's once per weight, the activations' by one small launch in front of the GEMM).  k_gemm_mx8 (msq_gemm_mx8.cu)
// assembles the scale chunks with producer warps through the generic proxy -- a store, a proxy fence and an mbarrier
// arrive per stage, which at decode sizes hides behind the weight stream but at M = 4096 was the per-stage time of the
// kernel (~2700 cycles per 128-K stage against 512 cycles of tensor work).
//
// Tile = 128 weight rows (UMMA M operand, so C's features sit on the TMEM lanes and one register of a warp is 128
// contiguous bytes of C) x BN tokens (UMMA N operand), BN <= 240 with TWO accumulators in tensor memory (2 x 240 fp32
// columns + 4 + 8 scale-factor columns <= 512): the four epilogue warps drain tile i while the MMAs of tile i + 1 run.
// BN = 256 (one accumulator) is used only when the token count makes it cheaper.  Persistent CTAs, tiles rastered in
// bands of 8 token tiles so that the CTAs running at the same time share weight AND activation tiles in L2.
//
// Weight operand: dense E4M3 bytes [N, K] (the widened copy, outliers folded: msq_widen_fp4_to_mxfp8), or the packed 4-bit
// codes themselves (plain FP4 weight; CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B expands 16 codes into a 16-byte chunk).
#include <stdlib.h>
#include <cuda_bf16.h>
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_tc.cuh"

namespace msq {
namespace {

constexpr int YM = 128, YBK = 128, YST = 4, YMAXN = 256, YGROUP = 8;
constexpr int YA_BYTES = YM * YBK, YB_BYTES = YMAXN * YBK, YSF_BYTES = 1024 + 2048;     // scale chunks of a stage (F4: 256 values of K)
constexpr int Y_SMEM = YST * (YA_BYTES + YB_BYTES + YSF_BYTES) + 1024 + 256;
static_assert(Y_SMEM <= 227 * 1024, "shared memory");

struct YParams {
  const uint8_t* wsf_t;       // weight scales, tiled:     chunk (tn, kb)    at (tn * kblocks + kb) * 512
  const uint8_t* xsf_t;       // activation scales, tiled: chunks (tm, kb)   at (tm * kblocks + kb) * 512 * sfb_chunks
  float* C; long long ldc;
  int M, N, K, bn, nacc, tiles_n, tiles_m, kblocks, sfb_chunks, accumulate, w_fmt, w_tx_bytes, pdl;
  int epi_h;                  // two-CTA kernel: epilogue warps per TMEM lane quarter (1 or 2)
  int ksplit, kb_per;         // one-CTA kernel: K slices per tile and stages per slice (ksplit = 1: no split)
  int c_bf16;                 // C is bf16 [M, ldc] (plain store, no accumulate / split-K) instead of fp32
  int terms;                  // the activations are `terms` stacked slabs of M rows (msq_split_a16_e4m3) that are ACCUMULATED:
                              // the K loop runs terms * kblocks stages, stage (t, kb) = weight tile kb x slab t's tile kb
  // sparse outlier correction fused into the epilogue (FP4 operand whose outlier slots the tensor core reads as
  // q * 2^e_in):  C[m, n] += sum_e x[m, o_k[e]] * o_dw[e]  over the CSR row of feature n; o_ptr NULL = none
  const int* o_ptr; const int* o_k; const float* o_dw;
  const uint8_t* xcodes; long long ldx; const int8_t* xe;
};

__device__ __forceinline__ uint64_t sf_desc(uint32_t smem_addr) {     // see make_sf_desc in msq_gemm_mx8.cu
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void bulk_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_sf(uint32_t tmem_addr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_addr), "l"(desc) : "memory");
}
__device__ __forceinline__ void mma_bs(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
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

// both operands packed FP4 (E2M1, two codes per byte, K-major), one UE8M0 scale per 32 values: K = 64 per instruction, two
// scale bytes per row and instruction (scale_vec::2X) taken from bytes {0, 1} or {2, 3} of the row's 32-bit cell (sf id 0 / 2)
__device__ __forceinline__ void mma_f4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                       uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}

// F4 = true: W4A4 -- the activations are packed FP4 as well (kind::mxf4, twice the tensor rate of kind::mxf8f6f4).  A stage
// is the same 128 bytes of K per row, i.e. 256 values: two 128-wide scale chunks per operand tile, four K = 64 instructions.
template <bool F4>
__global__ void __launch_bounds__(256, 1) k_gemm_mx8_tma(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                          const YParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + YST * YA_BYTES;
  uint8_t* sSF = smem + YST * (YA_BYTES + YB_BYTES);       // per stage: SFA[512] | SFB[1024]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSF + YST * YSF_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + YST;
  uint64_t* acc_full = bars + 2 * YST;
  uint64_t* acc_empty = bars + 2 * YST + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * YST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = P.tiles_n * P.tiles_m;
  // split-K (token counts whose tiles do not fill the SMs): a work item = (tile, K slice); slices add into C by atomics
  const int num_items = num_tiles * P.ksplit;
  const int band_tiles = YGROUP * P.tiles_n;
  auto tile_coords = [&](int t, int& tn, int& tm) {
    const int band = t / band_tiles;
    const int first = band * YGROUP;
    const int rows = min(YGROUP, P.tiles_m - first);
    const int r = t - band * band_tiles;
    tn = r / rows;
    tm = first + (r - tn * rows);
  };

  if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmX); tc::tma_prefetch_desc(&tmW); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < YST; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ---- loads: per stage one weight tile, one activation tile, one + sfb_chunks scale chunks, all on one mbarrier
    int stage = 0; uint32_t phase = 0;
    constexpr int KC = F4 ? 2 : 1;                  // 128-wide scale chunks of K per stage
    const uint32_t tx = (uint32_t)(P.w_tx_bytes + P.bn * YBK + KC * (512 + 512 * P.sfb_chunks));
    // programmatic dependent of k_tile_scales (which releases us at once): everything up to here overlapped it
    if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
      const int t = w / P.ksplit, ks = w - t * P.ksplit;
      const int kb0 = ks * P.kb_per, kb1 = min(P.kblocks / KC, kb0 + P.kb_per);
      int tn, tm; tile_coords(t, tn, tm);
      const uint8_t* wsf = P.wsf_t + (long long)tn * P.kblocks * 512;
      for (int term = 0; term < P.terms; ++term) {
        const uint8_t* xsf = P.xsf_t + ((long long)term * P.tiles_m + tm) * P.kblocks * 512 * P.sfb_chunks;
        const int xrow = term * P.M + tm * P.bn;
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&full[stage], tx);
            tc::tma_load_2d(sA + stage * YA_BYTES, &tmW, &full[stage], kb * YBK, tn * YM);
            tc::tma_load_2d(sB + stage * YB_BYTES, &tmX, &full[stage], kb * YBK, xrow);
            bulk_1d(sSF + stage * YSF_BYTES, wsf + (long long)kb * (512 * KC), 512u * KC, &full[stage]);
            bulk_1d(sSF + stage * YSF_BYTES + 512 * KC, xsf + (long long)kb * (512 * KC) * P.sfb_chunks,
                    512u * KC * (uint32_t)P.sfb_chunks, &full[stage]);
          }
          __syncwarp();
          if (++stage == YST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue (whole warp in the loop, the tcgen05 instructions under elect.sync: see msq_tc.cuh)
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t acc_phase = 0;           // bit `as` = phase of accumulator `as`
    const uint64_t ad0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sA)), bd0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sB));
    const uint64_t sfd0 = sf_desc(tc::smem_u32(sSF));
    const uint32_t a_step = (uint32_t)(YA_BYTES >> 4), b_step = (uint32_t)(YB_BYTES >> 4), sf_step = (uint32_t)(YSF_BYTES >> 4);
    // [4,6) b_sf_id  [7,10) A fmt  [10,13) B fmt  [17,23) N>>3  [23] scale fmt (1 = E8M0)  [24,29) M>>4  [29,31) a_sf_id
    // (F4: both formats 1 = E2M1 of kind::mxf4)
    const uint32_t idesc0 = (F4 ? ((1u << 7) | (1u << 10)) : ((uint32_t)P.w_fmt << 7)) | ((uint32_t)(P.bn >> 3) << 17) | (1u << 23) |
                            ((uint32_t)(YM >> 4) << 24);
    const uint32_t idesc_k = (1u << 4) | (1u << 29);
    // scale-factor columns behind the accumulators: F4 keeps the two K chunks of a stage side by side (A: 4 + 4, B: 8 + 8)
    const uint32_t tsfa = tmem_base + (uint32_t)(P.nacc * P.bn), tsfb = tsfa + (F4 ? 8 : 4);
    constexpr int KC = F4 ? 2 : 1;
    for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
      const int ks = w % P.ksplit;
      const int nst = P.terms * (min(P.kblocks / KC, ks * P.kb_per + P.kb_per) - ks * P.kb_per);
      tc::mbar_wait(&acc_empty[as], ((acc_phase >> as) & 1u) ^ 1u);
      tc::tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(as * P.bn);
      for (int kb = 0; kb < nst; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)stage * a_step), bd = bd0 + (uint64_t)((uint32_t)stage * b_step);
        const uint64_t sfd = sfd0 + (uint64_t)((uint32_t)stage * sf_step);
        if (tc::elect_one()) {
          if (F4) {
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
              cp_sf(tsfa + 4u * kc, sfd + 32u * kc);
              cp_sf(tsfb + 8u * kc, sfd + 64u + 32u * (uint32_t)(kc * P.sfb_chunks));
              if (P.sfb_chunks > 1) cp_sf(tsfb + 8u * kc + 4u, sfd + 64u + 32u * (uint32_t)(kc * P.sfb_chunks + 1));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)          // K = 64 each: 32 bytes along the 128-byte row, scale bytes (k & 1) * 2 ..
              mma_f4(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc0 + (uint32_t)((k & 1) * 2) * idesc_k, (kb | k) != 0,
                     tsfa + 4u * (uint32_t)(k >> 1), tsfb + 8u * (uint32_t)(k >> 1));
          } else {
            cp_sf(tsfa, sfd);
            cp_sf(tsfb, sfd + 32u);
            if (P.sfb_chunks > 1) cp_sf(tsfb + 4, sfd + 64u);
#pragma unroll
            for (int k = 0; k < YBK / 32; ++k)
              mma_bs(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc0 + (uint32_t)k * idesc_k, (kb | k) != 0, tsfa, tsfb);
          }
          tc::umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == YST) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&acc_full[as]);
      __syncwarp();
      acc_phase ^= 1u << as;
      if (P.nacc == 2) as ^= 1;
    }
  } else if (warp >= 4) {
    // ---- epilogue: TMEM lane = output feature, register = token; one register of a warp = 128 contiguous bytes of C
    const int q = warp & 3;
    int as = 0; uint32_t acc_phase = 0;           // bit `as` = phase of accumulator `as`
    for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
      const int t = w / P.ksplit, ks = w - t * P.ksplit;
      int tn, tm; tile_coords(t, tn, tm);
      const int n = tn * YM + q * 32 + lane;
      const int m0 = tm * P.bn;
      int oe0 = 0, oe1 = 0;
      if (P.o_ptr && n < P.N && ks == 0) { oe0 = __ldg(P.o_ptr + n); oe1 = __ldg(P.o_ptr + n + 1); }
      tc::mbar_wait(&acc_full[as], (acc_phase >> as) & 1u);
      tc::tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * P.bn);
      float* crow = P.C + (long long)m0 * P.ldc + n;
      for (int c = 0; c < P.bn; c += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tacc + (uint32_t)c)
            : "memory");
        tc::tmem_ld_wait();
        if (n < P.N) {
          if (oe1 > oe0) {      // this feature's outliers against the 16 tokens in registers
            const int kblk32 = P.K >> 5;
            for (int e = oe0; e < oe1; ++e) {
              const int k = __ldg(P.o_k + e);
              const float dw = __ldg(P.o_dw + e);
              // all 32 gathers are issued before the first use (rows clamped instead of predicated): one load latency
              // per (outlier, 16 tokens) instead of sixteen
              uint32_t cb[16];
              int eb[16];
#pragma unroll
              for (int v = 0; v < 16; ++v) {
                const long long m = min(m0 + c + v, P.M - 1);
                cb[v] = __ldg(P.xcodes + m * P.ldx + k);
                eb[v] = __ldg(P.xe + m * kblk32 + (k >> 5));
              }
#pragma unroll
              for (int v = 0; v < 16; ++v) {
                const float xs = eb[v] == MSQ_EXP_NAN ? qnan() : pow2i(eb[v]);
                const float xv = __fmul_rn(decode_elem(cb[v], 4, 5), xs);
                r[v] = __float_as_uint(fmaf(xv, dw, __uint_as_float(r[v])));
              }
            }
          }
#pragma unroll
          for (int v = 0; v < 16; ++v) {
            if (m0 + c + v < P.M) {
              const float val = __uint_as_float(r[v]);
              if (P.c_bf16) {
                reinterpret_cast<__nv_bfloat16*>(P.C)[(long long)(m0 + c + v) * P.ldc + n] = __float2bfloat16_rn(val);
              } else {
                float* p = crow + (long long)(c + v) * P.ldc;
                if (P.ksplit > 1) atomicAdd(p, val);
                else *p = P.accumulate ? (*p + val) : val;
              }
            }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[as]);
      acc_phase ^= 1u << as;
      if (P.nacc == 2) as ^= 1;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Two-CTA form (cta_group::2): a cluster of two CTAs owns a tile of 256 weight rows x BN tokens.  Each CTA stages ITS 128
// weight rows and ITS half of the tokens (BN/2 rows), the MMA -- issued by the leader CTA for both -- reads the halves of the
// token tile from both CTAs' shared memory: per SM and stage 16 KB + 16 KB land by TMA and are read by the tensor core instead
// of 16 KB + 32 KB (the one-CTA kernel sits at 107 of the 128 B/clk of shared-memory traffic an SM moves), and the ring is six
// stages deep.  All loads -- including the scale chunks, through uint32 tensor maps over the tiled scale arrays -- are
// cp.async.bulk.tensor.cta_group::2 that signal the LEADER's mbarrier; tcgen05.commit multicasts the "stage free" /
// "accumulator full" arrivals to both CTAs; the epilogue warps of both CTAs arrive on the leader's "accumulator free".
constexpr int ZST = 6;
constexpr int ZA_BYTES = YM * YBK, ZB_BYTES = (YMAXN / 2) * YBK, ZSF_BYTES = 1024 + 2048;
constexpr int Z_SMEM = ZST * (ZA_BYTES + ZB_BYTES + ZSF_BYTES) + 1024 + 256;
static_assert(Z_SMEM <= 227 * 1024, "shared memory");

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
  const uint32_t a = tc::smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 6000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void cp_sf2(uint32_t tmem_addr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(tmem_addr), "l"(desc) : "memory");
}
__device__ __forceinline__ void mma_bs2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                        uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}
__device__ __forceinline__ void mma_f4_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                         uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}
__device__ __forceinline__ void commit2_mc(uint64_t* bar) {       // arrive on `bar` in BOTH CTAs once the MMAs issued so far are done
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <bool F4>      // F4: both operands packed FP4 on kind::mxf4 (see k_gemm_mx8_tma<true>)
__global__ void __launch_bounds__(384, 1) k_gemm_mx8_tma2(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                           const __grid_constant__ CUtensorMap tmXSF,
                                                           const __grid_constant__ CUtensorMap tmWSF, const YParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ZST * ZA_BYTES;
  uint8_t* sSF = smem + ZST * (ZA_BYTES + ZB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSF + ZST * ZSF_BYTES);
  uint64_t* full = bars;                    // used in the leader CTA only
  uint64_t* empty = bars + ZST;
  uint64_t* acc_full = bars + 2 * ZST;
  uint64_t* acc_empty = bars + 2 * ZST + 2;  // used in the leader CTA only (512 arrivals: the epilogue threads of both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ZST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int num_tiles = P.tiles_n * P.tiles_m;                 // tiles_n counts 256-row tiles here
  const int band_tiles = YGROUP * P.tiles_n;
  auto tile_coords = [&](int t, int& tn, int& tm) {
    const int band = t / band_tiles;
    const int first = band * YGROUP;
    const int rows = min(YGROUP, P.tiles_m - first);
    const int r = t - band * band_tiles;
    tn = r / rows;
    tm = first + (r - tn * rows);
  };

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmX); tc::tma_prefetch_desc(&tmW); tc::tma_prefetch_desc(&tmXSF); tc::tma_prefetch_desc(&tmWSF);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < ZST; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 256u * (uint32_t)P.epi_h); }
    tc::mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  cluster_sync_all();                       // barrier inits and the allocation are visible in both CTAs
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int half = P.bn >> 1;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    constexpr int KC = F4 ? 2 : 1;
    const uint32_t tx_cta = (uint32_t)(P.w_tx_bytes + half * YBK + KC * (512 + 512 * P.sfb_chunks));
    if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int t = pair; t < num_tiles; t += npairs) {
      int tn, tm; tile_coords(t, tn, tm);
      const int wrow = tn * 2 * YM + (int)rank * YM;
      const int wsf_row = (tn * 2 + (int)rank) * P.kblocks;
      for (int term = 0; term < P.terms; ++term) {
        const int xrow = term * P.M + tm * P.bn + (int)rank * half;
        const int xsf_row = ((term * P.tiles_m + tm) * P.kblocks) * P.sfb_chunks;
        for (int kb = 0; kb < P.kblocks / KC; ++kb) {
          mbar_wait_b(&empty[stage], phase ^ 1);
          if (tc::elect_one()) {
            const uint32_t lbar = mapa_u32(tc::smem_u32(&full[stage]), 0);
            if (rank == 0) tc::mbar_expect_tx(&full[stage], 2u * tx_cta);
            tma_load_2d_cg2(sA + stage * ZA_BYTES, &tmW, lbar, kb * YBK, wrow);
            tma_load_2d_cg2(sB + stage * ZB_BYTES, &tmX, lbar, kb * YBK, xrow);
            tma_load_2d_cg2(sSF + stage * ZSF_BYTES, &tmWSF, lbar, 0, wsf_row + kb * KC);
            tma_load_2d_cg2(sSF + stage * ZSF_BYTES + 512 * KC, &tmXSF, lbar, 0, xsf_row + kb * KC * P.sfb_chunks);
          }
          __syncwarp();
          if (++stage == ZST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t acc_phase = 0;
    const uint64_t ad0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sA)), bd0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sB));
    const uint64_t sfd0 = sf_desc(tc::smem_u32(sSF));
    const uint32_t a_step = (uint32_t)(ZA_BYTES >> 4), b_step = (uint32_t)(ZB_BYTES >> 4), sf_step = (uint32_t)(ZSF_BYTES >> 4);
    const uint32_t idesc0 = (F4 ? ((1u << 7) | (1u << 10)) : ((uint32_t)P.w_fmt << 7)) | ((uint32_t)(P.bn >> 3) << 17) | (1u << 23) |
                            ((uint32_t)((2 * YM) >> 4) << 24);
    const uint32_t idesc_k = (1u << 4) | (1u << 29);
    const uint32_t tsfa = tmem_base + (uint32_t)(P.nacc * P.bn), tsfb = tsfa + (F4 ? 8 : 4);
    constexpr int KC = F4 ? 2 : 1;
    for (int t = pair; t < num_tiles; t += npairs) {
      mbar_wait_b(&acc_empty[as], ((acc_phase >> as) & 1u) ^ 1u);
      tc::tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(as * P.bn);
      for (int kb = 0; kb < P.terms * (P.kblocks / KC); ++kb) {
        mbar_wait_b(&full[stage], phase);
        tc::tc_fence_after();
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)stage * a_step), bd = bd0 + (uint64_t)((uint32_t)stage * b_step);
        const uint64_t sfd = sfd0 + (uint64_t)((uint32_t)stage * sf_step);
        if (tc::elect_one()) {
          if (F4) {
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
              cp_sf2(tsfa + 4u * kc, sfd + 32u * kc);
              cp_sf2(tsfb + 8u * kc, sfd + 64u + 32u * (uint32_t)(kc * P.sfb_chunks));
              if (P.sfb_chunks > 1) cp_sf2(tsfb + 8u * kc + 4u, sfd + 64u + 32u * (uint32_t)(kc * P.sfb_chunks + 1));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f4_2(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc0 + (uint32_t)((k & 1) * 2) * idesc_k, (kb | k) != 0,
                       tsfa + 4u * (uint32_t)(k >> 1), tsfb + 8u * (uint32_t)(k >> 1));
          } else {
            cp_sf2(tsfa, sfd);
            cp_sf2(tsfb, sfd + 32u);
            if (P.sfb_chunks > 1) cp_sf2(tsfb + 4, sfd + 64u);
#pragma unroll
            for (int k = 0; k < YBK / 32; ++k)
              mma_bs2(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc0 + (uint32_t)k * idesc_k, (kb | k) != 0, tsfa, tsfb);
          }
          commit2_mc(&empty[stage]);
        }
        __syncwarp();
        if (++stage == ZST) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) commit2_mc(&acc_full[as]);
      __syncwarp();
      acc_phase ^= 1u << as;
      if (P.nacc == 2) as ^= 1;
    }
  } else if (warp >= 4) {
    // ---- epilogue: EIGHT warps, two per TMEM lane quarter taking alternate 16-column chunks; the tcgen05.ld of a warp's
    // next chunk is in flight while it stores the current one.  With BN = 256 there is one accumulator and the drain is
    // exposed once per tile (ncu, 8192^3, four warps: tensor pipe 80 % active, the rest is this drain)
    const int q = warp & 3, h = (warp - 4) >> 2;
    int as = 0; uint32_t acc_phase = 0;
    auto ld16 = [&](uint32_t (&r)[16], uint32_t taddr) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr)
          : "memory");
    };
    for (int t = pair; t < num_tiles && h < P.epi_h; t += npairs) {
      int tn, tm; tile_coords(t, tn, tm);
      const int n = tn * 2 * YM + (int)rank * YM + q * 32 + lane;
      const int m0 = tm * P.bn;
      mbar_wait_b(&acc_full[as], (acc_phase >> as) & 1u);
      tc::tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * P.bn);
      float* crow = P.C + (long long)m0 * P.ldc + n;
      auto st16 = [&](const uint32_t (&r)[16], int c) {
        if (n < P.N) {
#pragma unroll
          for (int v = 0; v < 16; ++v) {
            if (m0 + c + v < P.M) {
              const float val = __uint_as_float(r[v]);
              if (P.c_bf16) {
                reinterpret_cast<__nv_bfloat16*>(P.C)[(long long)(m0 + c + v) * P.ldc + n] = __float2bfloat16_rn(val);
              } else {
                float* p = crow + (long long)(c + v) * P.ldc;
                *p = P.accumulate ? (*p + val) : val;
              }
            }
          }
        }
      };
      uint32_t ra[16], rb[16];
      const int half = 16 * P.epi_h;
      int c = h * 16;
      if (c < P.bn) ld16(ra, tacc + (uint32_t)c);
      for (; c < P.bn; c += 2 * half) {
        tc::tmem_ld_wait();
        if (c + half < P.bn) ld16(rb, tacc + (uint32_t)(c + half));
        st16(ra, c);
        if (c + half < P.bn) {
          tc::tmem_ld_wait();
          if (c + 2 * half < P.bn) ld16(ra, tacc + (uint32_t)(c + 2 * half));
          st16(rb, c + half);
        }
      }
      tc::tc_fence_before();
      arrive_cluster(mapa_u32(tc::smem_u32(&acc_empty[as]), 0));
      acc_phase ^= 1u << as;
      if (P.nacc == 2) as ^= 1;
    }
  }
  tc::tc_fence_before();
  cluster_sync_all();                       // nobody frees tensor memory (or leaves) while the peer still uses the pair
  if (warp == 2) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// scales [rows, K/32] (int8 shared exponents, or UE8M0 bytes) -> tcgen05.cp chunks: for every (tile of rows_per_tile rows,
// 128-wide K block) `chunks` = ceil(rows_per_tile / 128) pieces of 512 bytes, byte(r, j) = (r % 32) * 16 + (r / 32) * 4 + j
// for row r of the piece and 32-block j of the K block; rows outside the matrix get 2^0
__global__ void __launch_bounds__(256) k_tile_scales(const uint8_t* __restrict__ src, int add127, int rows, int kblk32,
                                                     int rows_per_tile, int chunks, int kblocks, uint8_t* __restrict__ out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the GEMM behind us waits before it reads `out`
  // blockIdx.z = slab of `rows` rows (stacked activation terms): each slab is tiled on its own
  src += (long long)blockIdx.z * rows * kblk32;
  out += (long long)blockIdx.z * gridDim.y * kblocks * chunks * 512;
  const int kb = blockIdx.x, tile = blockIdx.y;
  const int r = threadIdx.x;
  if (r >= chunks * 128) return;
  const long long row = (long long)tile * rows_per_tile + r;
  uint32_t w = 0x7F7F7F7Fu;
  if (r < rows_per_tile && row < rows) {
    w = __ldg(reinterpret_cast<const uint32_t*>(src + row * kblk32) + kb);
    if (add127) w = __vadd4(w, 0x7F7F7F7Fu);            // e + 127; MSQ_EXP_NAN (-128) becomes 0xFF = NaN
  }
  const int rr = r & 127;
  *reinterpret_cast<uint32_t*>(out + (((long long)tile * kblocks + kb) * chunks + (r >> 7)) * 512 + (rr & 31) * 16 + (rr >> 5) * 4) = w;
}

// sparse outlier correction as a pass of its own behind the GEMM (every SM works on it; inside the GEMM only the four
// epilogue warps of a CTA can): one thread per (outlier, token), C[m, n_e] += x[m, k_e] * dw_e by atomics
template <bool X4>      // X4: the activations are packed FP4 codes (two per byte, ldx in bytes) instead of E4M3 bytes
__global__ void __launch_bounds__(256) k_sparse_outlier_corr(const int* __restrict__ o_n, const int* __restrict__ o_k,
                                                             const float* __restrict__ o_dw, int nnz, const uint8_t* __restrict__ xcodes,
                                                             long long ldx, const int8_t* __restrict__ xe, int M, int K, float* C,
                                                             long long ldc, int fold_m) {
  // lanes = 32 consecutive outliers (sorted by feature: their atomics for one token fall into a few sectors of C's row),
  // warps = 8 tokens
  const int e = blockIdx.x * 32 + (threadIdx.x & 31);
  const int m = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (e >= nnz || m >= M) return;
  const int k = __ldg(o_k + e), n = __ldg(o_n + e);
  const float dw = __ldg(o_dw + e);
  const int xb = __ldg(xe + (long long)m * (K >> 5) + (k >> 5));
  const float xs = xb == MSQ_EXP_NAN ? qnan() : pow2i(xb);
  float xq;
  if (X4) xq = decode_elem(((uint32_t)__ldg(xcodes + (long long)m * ldx + (k >> 1)) >> ((k & 1) * 4)) & 0xFu, 2, 3);
  else xq = decode_elem((uint32_t)__ldg(xcodes + (long long)m * ldx + k), 4, 5);
  const float xv = __fmul_rn(xq, xs);
  atomicAdd(C + (long long)(fold_m ? m % fold_m : m) * ldc + n, __fmul_rn(xv, dw));
}

// token-tile width: fewest (waves over the SMs) x (bytes a tile pulls into shared memory) -- the kernel is bound by the
// L2 -> shared-memory fill, not by the tensor pipe; one accumulator (BN = 256) pays its exposed epilogue
static void choose_bn(int M, int N, int K, int sms, int& bn, int& nacc, bool f4 = false) {
  if (const char* e = getenv("MSQ_MX8T_BN")) {
    const int v = atoi(e);
    if (v >= 16 && v <= 256 && v % 16 == 0) { bn = v; nacc = v <= 240 ? 2 : 1; return; }      // (A/B switch)
  }
  const int tiles_n = (N + YM - 1) / YM;
  const double kblocks = (K + YBK - 1) / YBK;
  double best = 1e300;
  bn = 240; nacc = 2;
  const int cand[] = {256, 240, 224, 208, 192, 176, 160, 144, 128, 112, 96, 80, 64, 48, 32, 16};
  for (int c : cand) {
    if (c > 16 && c - 16 >= M) continue;                 // a narrower tile already holds every token
    // kind::mxf4: a tile's MMAs take half as long, the exposed drain of a single accumulator costs twice as much
    // (8192^3: 4.66 PFLOP/s at 256, 5.4-5.6 at 240 / 224)
    if (f4 && c > 240) continue;                         // (f4 is also set for the two-CTA form of either kind)
    const long long tiles = (long long)tiles_n * ((M + c - 1) / c);
    const double waves = (double)((tiles + sms - 1) / sms);
    const double cost = waves * (kblocks * (YA_BYTES + c * YBK + 4096.0) + (c > 240 ? 135000.0 : 0.0));
    if (cost < best - 1e-6) { best = cost; bn = c; nacc = c <= 240 ? 2 : 1; }
  }
}

// tiled scale chunks as a 2-D uint32 tensor: row = one 512-byte chunk
static int make_tmap_sf(CUtensorMap* out, const void* base, long long rows, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return MSQ_ERR_CUDA;
  cuuint64_t dims[2] = {128u, (cuuint64_t)rows};
  cuuint64_t strides[1] = {512u};
  cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MSQ_OK : MSQ_ERR_CUDA;
}

// which kernel and which token-tile width a problem gets (shared by the GEMM and the workspace / tile-width queries)
static void plan(int M, int N, int K, int sms, int& bn, int& nacc, int& cta2, bool f4 = false, int* ksplit = nullptr,
                 bool allow_split = true) {
  // measured (plain FP4 weight, 4096^2 / 8192^2): M = 64: 10.0 / 16.8 us one CTA, 10.6 / 17.0 us pairs; M = 256: 12.8 / 23.1 vs
  // 11.8 / 19.7; M = 4096: 65.1 / 243.9 vs 55.6 / 201.0 -- pairs from 256 tokens on (MSQ_MX8T_CTA2 = 0 / 1 forces a form)
  const char* e = getenv("MSQ_MX8T_CTA2");
  cta2 = e ? (e[0] == '1') : (M >= 256 && N >= 512);
  if (sms < 2) cta2 = 0;
  // split-K for the token counts between the decode sizes and the compute-bound ones, whose tiles leave most SMs idle
  // (4096^2, M = 256: 32 pairs of CTAs on 148 SMs): one-CTA form, every tile cut into K slices that add into C by atomics
  // (MSQ_MX8T_KSPLIT=0 switches it off).  Slices of at least 4 stages; used when the tiles fill less than half the SMs
  const char* ek = getenv("MSQ_MX8T_KSPLIT");
  int ks = 1, bn_s = 0;
  // measured (scripts/splitk_ab.py): up to 64 tokens and from 32 stages of K on the slices win (FP8 activations, 4096^2 M = 48:
  // 10.7 -> 9.0 us, 8192^2: 17.8 -> 13.8 us; kind::mxf4 8192^2: 12.6 -> 11.0 us); from 128 tokens on the zero-fill of C and the
  // atomics cost more than the narrow tiles (4096^2 M = 256: 12.8 -> 15.7 us), so those keep the plan without slices
  if (allow_split && !(ek && ek[0] == '0') && !(e && e[0] == '1') && M <= 64 && K / (f4 ? 2 * YBK : YBK) >= 32) {
    // the alternative to split-K is what choose_bn does on its own: narrow token tiles (4096^2, M = 64: 16 tokens per tile),
    // i.e. the weight tile streamed once per token tile.  Same cost model for both: bytes a CTA pulls into shared memory
    int bn1, nacc1;
    choose_bn(M, N, K, sms, bn1, nacc1, f4);
    const int tiles_n = (N + YM - 1) / YM;
    const int kst = K / (f4 ? 2 * YBK : YBK);                           // stages along K
    const long long tiles1 = (long long)tiles_n * ((M + bn1 - 1) / bn1);
    const double cost1 = (double)((tiles1 + sms - 1) / sms) * (kst * (YA_BYTES + bn1 * YBK + 4096.0) + (bn1 > 240 ? 135000.0 : 0.0));
    double best = 0.85 * cost1;                                          // the slices pay a zero-fill and atomics
    const int cand[] = {240, 224, 208, 192, 176, 160, 144, 128, 112, 96, 80, 64, 48, 32, 16};
    for (int c : cand) {
      if (c > 16 && c - 16 >= M) continue;
      const long long tiles = (long long)tiles_n * ((M + c - 1) / c);
      if (tiles * 2 > sms || kst < 8) continue;
      int k = (int)(sms / tiles);
      if (k > kst / 4) k = kst / 4;
      if (k > 16) k = 16;
      if (ek && ek[0] >= '2' && ek[0] <= '9') k = ek[0] - '0';
      if (k < 2) continue;
      const double cost = (double)((kst + k - 1) / k) * (YA_BYTES + c * YBK + 4096.0) + 30000.0;
      if (cost < best) { best = cost; ks = k; bn_s = c; }
    }
  }
  if (ksplit) *ksplit = ks;
  if (ks > 1) { cta2 = 0; bn = bn_s; nacc = 2; return; }
  // pairs of CTAs over 256-row tiles; with eight epilogue warps per CTA two accumulators of <= 240 columns beat the single
  // 256-column one for FP8 activations as well (8192^3: 2.67 vs 2.94-2.98 PFLOP/s, scripts/mx8_bn_sweep.py)
  if (cta2) choose_bn(M, ((N + 2 * YM - 1) / (2 * YM)) * YM, K, sms / 2, bn, nacc, true);
  else choose_bn(M, N, K, sms, bn, nacc, f4);
}

// K slices of the one-CTA kernel: stages per slice, slices recounted so that none is empty; zero-fills C when slices add
static int setup_split(YParams& P, int ksplit, bool f4, cudaStream_t s) {
  const int kst = P.kblocks / (f4 ? 2 : 1);
  P.ksplit = 1; P.kb_per = kst;
  if (ksplit > 1) {
    P.kb_per = (kst + ksplit - 1) / ksplit;
    P.ksplit = (kst + P.kb_per - 1) / P.kb_per;
  }
  if (P.ksplit > 1 && !P.accumulate) {
    cudaError_t e = P.ldc == P.N ? cudaMemsetAsync(P.C, 0, (size_t)P.M * P.N * sizeof(float), s)
                                 : cudaMemset2DAsync(P.C, (size_t)P.ldc * sizeof(float), 0, (size_t)P.N * sizeof(float), (size_t)P.M, s);
    if (e != cudaSuccess) return MSQ_ERR_CUDA;
  }
  return MSQ_OK;
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" {

static int tile_scales(const void* scales, int is_int8_exponent, int rows, int K, int rows_per_tile, unsigned char* out, int slabs,
                       void* stream);

int msq_tile_scales_ue8m0(const void* scales, int is_int8_exponent, int rows, int K, int rows_per_tile, unsigned char* out,
                          void* stream) {
  return tile_scales(scales, is_int8_exponent, rows, K, rows_per_tile, out, 1, stream);
}

static int tile_scales(const void* scales, int is_int8_exponent, int rows, int K, int rows_per_tile, unsigned char* out, int slabs,
                       void* stream) {
  msq::DevGuard dev_guard_(scales);
  if (!scales || !out || rows <= 0 || K <= 0 || rows_per_tile <= 0 || rows_per_tile > 256 || slabs < 1) return MSQ_ERR_BAD_ARG;
  if (K % 128 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(scales) & 3) || (reinterpret_cast<uintptr_t>(out) & 15)) return MSQ_ERR_ALIGN;
  if (slabs > 1 && ((long long)rows * (K / 32)) % 4 != 0) return MSQ_ERR_ALIGN;          // every slab starts on a 32-bit word
  const int chunks = (rows_per_tile + 127) / 128;
  dim3 grid((unsigned)(K / YBK), (unsigned)((rows + rows_per_tile - 1) / rows_per_tile), (unsigned)slabs);
  k_tile_scales<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(scales), is_int8_exponent ? 1 : 0, rows,
                                                        K / 32, rows_per_tile, chunks, K / YBK, out);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

// C[m mod fold_m (fold_m = 0: m), o_n[e]] += dq(x)[m, o_k[e]] * o_dw[e] for every outlier e and token m: the sparse part of a
// linear whose dense part ran on the packed codes (msq_gemm_mxfp8_tma without o_ptr); COO triplets, best sorted by feature
static int sparse_correct(bool x4, const unsigned char* xcodes, long long ldx, const signed char* xe, const int* o_n, const int* o_k,
                          const float* o_dw, int nnz, float* C, long long ldc, int M, int K, int fold_m, void* stream);
int msq_sparse_outlier_correct(const unsigned char* xcodes, long long ldx, const signed char* xe, const int* o_n, const int* o_k,
                               const float* o_dw, int nnz, float* C, long long ldc, int M, int K, int fold_m, void* stream) {
  return sparse_correct(false, xcodes, ldx, xe, o_n, o_k, o_dw, nnz, C, ldc, M, K, fold_m, stream);
}
// the same behind the W4A4 GEMM: xcodes4 = packed FP4 activations (two codes per byte, low nibble = lower k; ldx in bytes)
int msq_sparse_outlier_correct_fp4(const unsigned char* xcodes4, long long ldx, const signed char* xe, const int* o_n, const int* o_k,
                                   const float* o_dw, int nnz, float* C, long long ldc, int M, int K, int fold_m, void* stream) {
  return sparse_correct(true, xcodes4, ldx, xe, o_n, o_k, o_dw, nnz, C, ldc, M, K, fold_m, stream);
}
static int sparse_correct(bool x4, const unsigned char* xcodes, long long ldx, const signed char* xe, const int* o_n, const int* o_k,
                          const float* o_dw, int nnz, float* C, long long ldc, int M, int K, int fold_m, void* stream) {
  msq::DevGuard dev_guard_(xcodes);
  if (!xcodes || !xe || !C || M <= 0 || K <= 0 || nnz < 0 || fold_m < 0) return MSQ_ERR_BAD_ARG;
  if (nnz == 0) return MSQ_OK;
  if (!o_n || !o_k || !o_dw) return MSQ_ERR_BAD_ARG;
  dim3 grid((unsigned)((nnz + 31) / 32), (unsigned)((M + 7) / 8));
  if (grid.y > 65535u) return MSQ_ERR_UNSUPPORTED;
  if (x4)
    k_sparse_outlier_corr<true><<<grid, 256, 0, (cudaStream_t)stream>>>(o_n, o_k, o_dw, nnz, xcodes, ldx, reinterpret_cast<const int8_t*>(xe),
                                                                        M, K, C, ldc, fold_m);
  else
    k_sparse_outlier_corr<false><<<grid, 256, 0, (cudaStream_t)stream>>>(o_n, o_k, o_dw, nnz, xcodes, ldx, reinterpret_cast<const int8_t*>(xe),
                                                                         M, K, C, ldc, fold_m);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_gemm_mxfp8_tma_tile_tokens(int M, int N, int K) {
  int dev = 0, sms = 148, bn, nacc;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  int cta2;
  plan(M, N, K, sms, bn, nacc, cta2);
  return bn;
}

static int tile_tokens(int M, int N, int K, bool f4) {
  int dev = 0, sms = 148, bn, nacc, cta2;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  plan(M, N, K, sms, bn, nacc, cta2, f4);
  return bn;
}
int msq_gemm_mxfp4_tma_tile_tokens(int M, int N, int K) { return tile_tokens(M, N, K, true); }
// K slices the plan gives a problem (1 = no split-K); w4a4 != 0: the kind::mxf4 pairing
int msq_gemm_mx_tma_k_slices(int M, int N, int K, int w4a4) {
  int dev = 0, sms = 148, bn, nacc, cta2, ks = 1;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  plan(M, N, K, sms, bn, nacc, cta2, w4a4 != 0, &ks);
  return ks;
}
long long msq_gemm_mxfp4_tma_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int bn = tile_tokens(M, N, K, true);
  const long long tiles_m = (M + bn - 1) / bn, kblocks = (K + YBK - 1) / YBK, chunks = (bn + 127) / 128;
  return tiles_m * kblocks * chunks * 512 + 1024;
}

// bytes for the tiled activation scales of up to three stacked terms
long long msq_gemm_mxfp8_tma_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int bn = msq_gemm_mxfp8_tma_tile_tokens(M, N, K);
  const long long tiles_m = (M + bn - 1) / bn, kblocks = (K + YBK - 1) / YBK, chunks = (bn + 127) / 128;
  return 3 * tiles_m * kblocks * chunks * 512 + 1024;
}

static int gemm_mxfp8(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* w, long long ldw,
                      int w_is_fp4, const unsigned char* wsf_tiled, const int* o_ptr, const int* o_k, const float* o_dw, void* Cv,
                      int c_is_bf16, long long ldc, int M, int N, int K, int terms, int accumulate, void* workspace,
                      long long workspace_bytes, void* stream);
int msq_gemm_mxfp8_tma(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* w, long long ldw,
                       int w_is_fp4, const unsigned char* wsf_tiled, const int* o_ptr, const int* o_k, const float* o_dw, float* C,
                       long long ldc, int M, int N, int K, int terms, int accumulate, void* workspace, long long workspace_bytes,
                       void* stream) {
  return gemm_mxfp8(xcodes, ldx, xe, w, ldw, w_is_fp4, wsf_tiled, o_ptr, o_k, o_dw, C, 0, ldc, M, N, K, terms, accumulate, workspace,
                    workspace_bytes, stream);
}
// the same with a bf16 result [M, ldc] (one term, no accumulate, no fused sparse epilogue, no K slices); workspace:
// msq_gemm_mxfp8_tma_bf16_workspace_bytes
long long msq_gemm_mxfp8_tma_bf16_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int dev = 0, sms = 148, bn, nacc, cta2;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  plan(M, N, K, sms, bn, nacc, cta2, false, nullptr, false);
  const long long tiles_m = (M + bn - 1) / bn, kblocks = (K + YBK - 1) / YBK, chunks = (bn + 127) / 128;
  return tiles_m * kblocks * chunks * 512 + 1024;
}
int msq_gemm_mxfp8_tma_bf16(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* w, long long ldw,
                            int w_is_fp4, const unsigned char* wsf_tiled, void* C_bf16, long long ldc, int M, int N, int K,
                            void* workspace, long long workspace_bytes, void* stream) {
  return gemm_mxfp8(xcodes, ldx, xe, w, ldw, w_is_fp4, wsf_tiled, nullptr, nullptr, nullptr, C_bf16, 1, ldc, M, N, K, 1, 0, workspace,
                    workspace_bytes, stream);
}
static int gemm_mxfp8(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* w, long long ldw,
                      int w_is_fp4, const unsigned char* wsf_tiled, const int* o_ptr, const int* o_k, const float* o_dw, void* Cv,
                      int c_is_bf16, long long ldc, int M, int N, int K, int terms, int accumulate, void* workspace,
                      long long workspace_bytes, void* stream) {
  float* C = reinterpret_cast<float*>(Cv);
  msq::DevGuard dev_guard_(xcodes);
  if (!xcodes || !xe || !w || !wsf_tiled || !C || !workspace || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (K % 128 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xe) & 3) || (reinterpret_cast<uintptr_t>(xcodes) & 15) || (ldx % 16) ||
      (reinterpret_cast<uintptr_t>(w) & 31) || (ldw % 16) || (reinterpret_cast<uintptr_t>(wsf_tiled) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return MSQ_ERR_ALIGN;
  if (w_is_fp4 && (ldw % 32)) return MSQ_ERR_ALIGN;
  if (o_ptr && (!o_k || !o_dw)) return MSQ_ERR_BAD_ARG;
  if (terms < 1 || terms > 3 || (terms > 1 && o_ptr)) return MSQ_ERR_BAD_ARG;     // stacked terms: corrections by msq_sparse_outlier_correct
  if (workspace_bytes < (c_is_bf16 ? msq_gemm_mxfp8_tma_bf16_workspace_bytes(M, N, K) : msq_gemm_mxfp8_tma_workspace_bytes(M, N, K)))
    return MSQ_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  YParams P;
  memset(&P, 0, sizeof(P));
  int cta2 = 0, ksplit = 1;
  P.c_bf16 = c_is_bf16 ? 1 : 0;
  plan(M, N, K, sms, P.bn, P.nacc, cta2, false, &ksplit, !c_is_bf16);
  if (o_ptr && cta2) { cta2 = 0; choose_bn(M, N, K, sms, P.bn, P.nacc); }       // the fused sparse epilogue lives in the one-CTA kernel
  P.wsf_t = wsf_tiled; P.xsf_t = reinterpret_cast<const uint8_t*>(workspace); P.C = C; P.ldc = ldc;
  P.M = M; P.N = N; P.K = K; P.accumulate = accumulate; P.terms = terms;
  P.o_ptr = o_ptr; P.o_k = o_k; P.o_dw = o_dw; P.xcodes = xcodes; P.ldx = ldx; P.xe = xe;
  P.tiles_n = (N + YM - 1) / YM; P.tiles_m = (M + P.bn - 1) / P.bn; P.kblocks = K / YBK; P.sfb_chunks = (P.bn + 127) / 128;
  P.w_fmt = w_is_fp4 ? 5 : 0;
  // a 16U4_ALIGN16B load reports the bytes it READ to the mbarrier (see msq_gemm_mx8.cu)
  P.w_tx_bytes = w_is_fp4 ? YA_BYTES / 2 : YA_BYTES;
  // (the zero-fill of a split-K result goes in front of the scale tiling, so that the GEMM stays the programmatic dependent
  // of the tiling kernel)
  int rc = setup_split(P, cta2 ? 1 : ksplit, false, s);
  if (rc) return rc;
  rc = tile_scales(xe, 1, M, K, P.bn, reinterpret_cast<unsigned char*>(workspace), terms, stream);
  if (rc) return rc;
  CUtensorMap tmX, tmW;
  rc = make_tmap_2d(&tmX, xcodes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (long long)terms * M, K, ldx, P.bn, YBK);
  if (rc) return rc;
  if (w_is_fp4) rc = make_tmap_2d_u4(&tmW, w, N, K, ldw, YM);
  else rc = make_tmap_2d(&tmW, w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, ldw, YM, YBK);
  if (rc) return rc;
  if (cta2) {
    // cluster of two CTAs per 256-row tile; each CTA loads its half of the token tile
    P.tiles_n = (N + 2 * YM - 1) / (2 * YM);
    CUtensorMap tmXh, tmXSF, tmWSF;
    rc = make_tmap_2d(&tmXh, xcodes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (long long)terms * M, K, ldx, P.bn / 2, YBK);
    if (rc) return rc;
    rc = make_tmap_sf(&tmXSF, workspace, (long long)terms * P.tiles_m * P.kblocks * P.sfb_chunks, P.sfb_chunks);
    if (rc) return rc;
    rc = make_tmap_sf(&tmWSF, wsf_tiled, (long long)((N + YM - 1) / YM) * P.kblocks, 1);
    if (rc) return rc;
    if (cudaFuncSetAttribute(k_gemm_mx8_tma2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Z_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
    int pairs = P.tiles_n * P.tiles_m;
    if (pairs > sms / 2) pairs = sms / 2;
    P.pdl = getenv("MSQ_MX8_NO_PDL") == nullptr ? 1 : 0;
    P.epi_h = getenv("MSQ_MX8T_EPI4") == nullptr ? 2 : 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = Z_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = P.pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, k_gemm_mx8_tma2<false>, tmXh, tmW, tmXSF, tmWSF, P) == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
  }
  if (cudaFuncSetAttribute(k_gemm_mx8_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Y_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
  int grid = P.tiles_n * P.tiles_m * P.ksplit;
  if (grid > sms) grid = sms;
  if (getenv("MSQ_MX8_NO_PDL") == nullptr) {
    P.pdl = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = Y_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_gemm_mx8_tma<false>, tmX, tmW, P) == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
  }
  k_gemm_mx8_tma<false><<<grid, 256, Y_SMEM, s>>>(tmX, tmW, P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

// W4A4: C[m, n] (+)= sum_k dq(X)[m, k] * dq(W)[n, k] with BOTH operands packed FP4 (E2M1, two codes per byte, low nibble
// first, K-major) and int8 shared exponents per 32 values of K (activations) / pre-tiled UE8M0 chunks (weight,
// msq_tile_scales_ue8m0): tcgen05.mma kind::mxf4.block_scale.scale_vec::2X, K = 64 per instruction.  ldx / ldw in BYTES.
// Workspace: msq_gemm_mxfp4_tma_workspace_bytes(M, N, K).  K % 256 == 0.
static int gemm_mxfp4(const unsigned char* xcodes4, long long ldx, const signed char* xe, const unsigned char* w4, long long ldw,
                      const unsigned char* wsf_tiled, void* Cv, int c_is_bf16, long long ldc, int M, int N, int K, int accumulate,
                      void* workspace, long long workspace_bytes, void* stream);
int msq_gemm_mxfp4_tma(const unsigned char* xcodes4, long long ldx, const signed char* xe, const unsigned char* w4, long long ldw,
                       const unsigned char* wsf_tiled, float* C, long long ldc, int M, int N, int K, int accumulate, void* workspace,
                       long long workspace_bytes, void* stream) {
  return gemm_mxfp4(xcodes4, ldx, xe, w4, ldw, wsf_tiled, C, 0, ldc, M, N, K, accumulate, workspace, workspace_bytes, stream);
}
// the same with a bf16 result [M, ldc] (rounded once from the fp32 accumulator; no accumulate, no K slices): saves the
// conversion pass behind the GEMM when the caller wants 16-bit activations back.  Workspace:
// msq_gemm_mxfp4_tma_bf16_workspace_bytes (the plan without K slices can pick another token-tile width)
long long msq_gemm_mxfp4_tma_bf16_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int dev = 0, sms = 148, bn, nacc, cta2;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  plan(M, N, K, sms, bn, nacc, cta2, true, nullptr, false);
  const long long tiles_m = (M + bn - 1) / bn, kblocks = (K + YBK - 1) / YBK, chunks = (bn + 127) / 128;
  return tiles_m * kblocks * chunks * 512 + 1024;
}
int msq_gemm_mxfp4_tma_bf16(const unsigned char* xcodes4, long long ldx, const signed char* xe, const unsigned char* w4, long long ldw,
                            const unsigned char* wsf_tiled, void* C_bf16, long long ldc, int M, int N, int K, void* workspace,
                            long long workspace_bytes, void* stream) {
  return gemm_mxfp4(xcodes4, ldx, xe, w4, ldw, wsf_tiled, C_bf16, 1, ldc, M, N, K, 0, workspace, workspace_bytes, stream);
}
static int gemm_mxfp4(const unsigned char* xcodes4, long long ldx, const signed char* xe, const unsigned char* w4, long long ldw,
                      const unsigned char* wsf_tiled, void* Cv, int c_is_bf16, long long ldc, int M, int N, int K, int accumulate,
                      void* workspace, long long workspace_bytes, void* stream) {
  float* C = reinterpret_cast<float*>(Cv);
  msq::DevGuard dev_guard_(xcodes4);
  if (!xcodes4 || !xe || !w4 || !wsf_tiled || !C || !workspace || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (K % 256 != 0) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xe) & 3) || (reinterpret_cast<uintptr_t>(xcodes4) & 15) || (ldx % 16) ||
      (reinterpret_cast<uintptr_t>(w4) & 15) || (ldw % 16) || (reinterpret_cast<uintptr_t>(wsf_tiled) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return MSQ_ERR_ALIGN;
  if (workspace_bytes < (c_is_bf16 ? msq_gemm_mxfp4_tma_bf16_workspace_bytes(M, N, K) : msq_gemm_mxfp4_tma_workspace_bytes(M, N, K)))
    return MSQ_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  YParams P;
  memset(&P, 0, sizeof(P));
  int cta2 = 0, ksplit = 1;
  P.c_bf16 = c_is_bf16 ? 1 : 0;
  plan(M, N, K, sms, P.bn, P.nacc, cta2, true, &ksplit, !c_is_bf16);       // same plan as the workspace query (the tiled activation scales depend on bn)
  P.wsf_t = wsf_tiled; P.xsf_t = reinterpret_cast<const uint8_t*>(workspace); P.C = C; P.ldc = ldc;
  P.M = M; P.N = N; P.K = K; P.accumulate = accumulate; P.terms = 1;
  P.tiles_n = (N + YM - 1) / YM; P.tiles_m = (M + P.bn - 1) / P.bn; P.kblocks = K / YBK; P.sfb_chunks = (P.bn + 127) / 128;
  P.w_tx_bytes = YA_BYTES;
  int rc = setup_split(P, cta2 ? 1 : ksplit, true, s);
  if (rc) return rc;
  rc = tile_scales(xe, 1, M, K, P.bn, reinterpret_cast<unsigned char*>(workspace), 1, stream);
  if (rc) return rc;
  CUtensorMap tmX, tmW;
  rc = make_tmap_2d(&tmX, xcodes4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, K / 2, ldx, P.bn, YBK);
  if (rc) return rc;
  rc = make_tmap_2d(&tmW, w4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K / 2, ldw, YM, YBK);
  if (rc) return rc;
  if (cta2) {
    P.tiles_n = (N + 2 * YM - 1) / (2 * YM);
    CUtensorMap tmXh, tmXSF, tmWSF;
    rc = make_tmap_2d(&tmXh, xcodes4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, K / 2, ldx, P.bn / 2, YBK);
    if (rc) return rc;
    rc = make_tmap_sf(&tmXSF, workspace, (long long)P.tiles_m * P.kblocks * P.sfb_chunks, 2 * P.sfb_chunks);
    if (rc) return rc;
    rc = make_tmap_sf(&tmWSF, wsf_tiled, (long long)((N + YM - 1) / YM) * P.kblocks, 2);
    if (rc) return rc;
    if (cudaFuncSetAttribute(k_gemm_mx8_tma2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Z_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
    int pairs = P.tiles_n * P.tiles_m;
    if (pairs > sms / 2) pairs = sms / 2;
    P.pdl = 1;
    P.epi_h = 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = Z_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, k_gemm_mx8_tma2<true>, tmXh, tmW, tmXSF, tmWSF, P) == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
  }
  if (cudaFuncSetAttribute(k_gemm_mx8_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Y_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
  int grid = P.tiles_n * P.tiles_m * P.ksplit;
  if (grid > sms) grid = sms;
  P.pdl = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = Y_SMEM; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_gemm_mx8_tma<true>, tmX, tmW, P) == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"
