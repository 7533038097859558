// msq_gemm_w4.cu -- packed-weight linear on tcgen05:   C[m, n] (+)= sum_k X[m, k] * W[n, k]
//
// W is the packed MicroScopiQ tensor [N, K] blocked along K (block 32, inner layout): codes +
// outlier bitmap + two int8 shared exponents per block (same code width for inliers and outliers,
// the w_elem_format/w_elem_format pairing of number_system/mx/linear.py:78-85).  X is bf16.
//
// The weight tile is the UMMA *M* operand (128 output features), the activations the UMMA *N*
// operand (BNX = 16..256 tokens), so a small batch does not waste the tensor core.  Per K-block of 64:
//   warp 0      : TMA loads the X tile [BNX x 64] (128B swizzle)
//   warps 8..11 : dequantize the W tile [128 x 64] from the packed planes (LUT + power-of-two scale,
//                 exact in bf16) and write it to shared memory in the same 128B-swizzled K-major
//                 layout TMA would produce; fence.proxy.async; arrive
//   warp 1      : tcgen05.mma.kind::f16 128 x BNX x 16, fp32 accumulators in TMEM (double buffered)
//   warps 4..7  : epilogue, TMEM -> registers -> C[m, n] (lane = n: coalesced rows of C)
// Dequantized weight values q * 2^e are exactly representable in bf16, products are exact in the
// fp32 accumulator: the result differs from the reference's fp32 F.linear on fake-quantized
// operands (linear.py:91) only by summation order.
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_tc.cuh"

namespace msq {

constexpr int WM = 128, WBK = 64, WSTAGES = 4, WMAXN = 256;
constexpr int WA_BYTES = WM * WBK * 2, WB_BYTES = WMAXN * WBK * 2;
constexpr int W4_SMEM = WSTAGES * (WA_BYTES + WB_BYTES) + 1024 + 256 + 2 * 256 * 2;

struct W4Params {
  const uint8_t* codes; const uint32_t* bitmap; const int8_t* e_in; const int8_t* e_out;
  float* C; long long ldc;
  int M, N, K;              // C is [M, N]; W is [N, K]; X is [M, K]
  int bnx;                  // activation tile (UMMA N), multiple of 16
  int tiles_n, tiles_m;
  int wi;                   // code width: 2, 4 or 8
  int in_ebits, in_bits, out_ebits, out_bits;
  int plain;                // no bitmap / outliers
  int accumulate;           // C += instead of C =
  int ksplit;               // K is cut into ksplit slices (atomicAdd epilogue when > 1)
  int kb_per_split;         // 64-wide K blocks per slice
};

__global__ void __launch_bounds__(512, 1) k_gemm_w4(const __grid_constant__ CUtensorMap tmX, const W4Params P) {
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
  __nv_bfloat16* lut_in = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(bars) + 256);
  __nv_bfloat16* lut_out = lut_in + 256;

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
    const bool valid = c < (1 << P.wi);
    lut_in[c] = __float2bfloat16_rn(valid ? decode_elem((uint32_t)c, P.in_ebits, P.in_bits) : 0.0f);
    lut_out[c] = __float2bfloat16_rn(valid ? decode_elem((uint32_t)c, P.out_ebits, P.out_bits) : 0.0f);
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
    // ---- weight dequant producers: 256 threads, 4 chunks (8 elements = 16 B of bf16) each per stage.
    // The packed bytes are prefetched PF-1 stages ahead into a register ring (statically indexed:
    // the stage loop is unrolled by PF), so ~PF stages of global loads per thread are in flight
    // while the current stage is converted.
    constexpr int PF = 3;
    const int ptid = threadIdx.x - 256;
    const int wi = P.wi;
    const long long code_row_bytes = ((long long)P.K * wi) >> 3;
    const int kblk32 = P.K >> 5;                        // MX blocks per weight row
    // raw loaded words only: nothing is computed from them until convert(), so the loads of the
    // prefetched stages stay in flight
    struct Raw { uint32_t c0, c1, bm; int ei, eo; };
    auto fetch = [&](int n0, int kb, Raw (&raw)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = ptid + 256 * i;
        const int r = c >> 3, kc = c & 7;
        const int n = n0 + r, k = kb * WBK + kc * 8;
        raw[i].c0 = raw[i].c1 = raw[i].bm = 0; raw[i].ei = 0; raw[i].eo = 0;
        if (n < P.N && k < P.K) {
          const uint8_t* cp = P.codes + (long long)n * code_row_bytes + (((long long)k * wi) >> 3);
          if (wi == 4) raw[i].c0 = __ldg(reinterpret_cast<const uint32_t*>(cp));
          else if (wi == 2) raw[i].c0 = __ldg(reinterpret_cast<const uint16_t*>(cp));
          else { const uint2 v = __ldg(reinterpret_cast<const uint2*>(cp)); raw[i].c0 = v.x; raw[i].c1 = v.y; }
          const int blk = n * kblk32 + (k >> 5);
          raw[i].ei = __ldg(P.e_in + blk);
          if (!P.plain) {
            raw[i].bm = __ldg(P.bitmap + (((long long)n * P.K + k) >> 5));
            raw[i].eo = __ldg(P.e_out + blk);
          }
        }
      }
    };
    // 4-bit sign-magnitude codes with one format for inliers and outliers (MXLinear's pairing):
    // magnitude -> bf16 through byte-permute lookups held in registers (8 magnitudes: low-byte and
    // high-byte tables of 8 bytes each), sign moved to bit 15, one HMUL2 per pair for the block
    // scale; the rare outlier elements are rescaled afterwards.
    const bool fast4 = (wi == 4) && (P.plain || (P.in_ebits == P.out_ebits && P.in_bits == P.out_bits));
    uint32_t tl0 = 0, tl1 = 0, th0 = 0, th1 = 0;
    if (fast4) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const uint32_t bb = (uint32_t)__bfloat16_as_ushort(lut_in[m]);
        if (m < 4) { tl0 |= (bb & 0xFFu) << (8 * m); th0 |= (bb >> 8) << (8 * m); }
        else { tl1 |= (bb & 0xFFu) << (8 * (m - 4)); th1 |= (bb >> 8) << (8 * (m - 4)); }
      }
    }
    auto convert = [&](uint8_t* dst, const Raw (&raw)[4], int n0, int kb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = ptid + 256 * i;
        const int r = c >> 3, kc = c & 7;
        uint4 outv = make_uint4(0, 0, 0, 0);
        const int k = kb * WBK + kc * 8;
        if (n0 + r < P.N && k < P.K) {
          const int ei = raw[i].ei, eo = raw[i].eo;
          uint32_t ob = (raw[i].bm >> (k & 31)) & 0xFFu;
          const bool nan_blk = ei == MSQ_EXP_NAN || eo == MSQ_EXP_NAN;
          // inlier: q * 2^e_in ; outlier: q * 2^(e_out - e_in)   (both exact in bf16)
          const __nv_bfloat16 s_in = __float2bfloat16_rn(nan_blk ? qnan() : pow2i(ei));
          const int eoi = max(-126, min(127, eo - ei));
          const __nv_bfloat16 s_out = __float2bfloat16_rn(nan_blk ? qnan() : pow2i(eoi));
          uint32_t w[4];
          if (fast4) {
            const uint32_t sbits = (uint32_t)__bfloat16_as_ushort(s_in);
            const uint32_t s2 = sbits | (sbits << 16);
            const uint32_t cw = raw[i].c0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {                      // 4 codes per half
              const uint32_t c4 = (cw >> (16 * h)) & 0xFFFFu;
              const uint32_t sel = c4 & 0x7777u;               // magnitudes as permute selectors
              const uint32_t lo = __byte_perm(tl0, tl1, sel);  // 4 low bytes
              const uint32_t hi = __byte_perm(th0, th1, sel);  // 4 high bytes
              uint32_t p0 = __byte_perm(lo, hi, 0x5140);       // (lo0,hi0,lo1,hi1)
              uint32_t p1 = __byte_perm(lo, hi, 0x7362);       // (lo2,hi2,lo3,hi3)
              p0 |= ((c4 & 0x8u) << 12) | ((c4 & 0x80u) << 24);
              p1 |= ((c4 & 0x800u) << 4) | ((c4 & 0x8000u) << 16);
              __nv_bfloat162 v0 = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&p0), *reinterpret_cast<const __nv_bfloat162*>(&s2));
              __nv_bfloat162 v1 = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&p1), *reinterpret_cast<const __nv_bfloat162*>(&s2));
              w[2 * h] = *reinterpret_cast<uint32_t*>(&v0);
              w[2 * h + 1] = *reinterpret_cast<uint32_t*>(&v1);
            }
            while (ob) {                                       // outliers: q * 2^(e_out - e_in)
              const int e = __ffs(ob) - 1;
              ob &= ob - 1;
              const uint32_t code = (cw >> (4 * e)) & 0xFu;
              const uint32_t bits = (uint32_t)__bfloat16_as_ushort(__hmul(lut_in[code], s_out));
              const int wsel = e >> 1;
              w[wsel] = (e & 1) ? ((w[wsel] & 0x0000FFFFu) | (bits << 16)) : ((w[wsel] & 0xFFFF0000u) | bits);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              uint32_t code;
              if (wi == 8) code = ((e < 4 ? raw[i].c0 : raw[i].c1) >> ((e & 3) * 8)) & 0xFFu;
              else code = (raw[i].c0 >> (e * wi)) & ((1u << wi) - 1u);
              const bool isout = (ob >> e) & 1u;
              const __nv_bfloat16 q = isout ? lut_out[code] : lut_in[code];
              const __nv_bfloat16 v = __hmul(q, isout ? s_out : s_in);
              const uint32_t bits = (uint32_t)__bfloat16_as_ushort(v);
              if (e & 1) w[e >> 1] |= bits << 16; else w[e >> 1] = bits;
            }
          }
          outv = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(dst + r * 128 + ((kc ^ (r & 7)) << 4)) = outv;
      }
    };
    // cursor over this CTA's (tile, K-block) sequence
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
    Raw ring[PF][4];
#pragma unroll
    for (int j = 0; j < PF - 1; ++j) {
      if (fc.ok) { fetch(fc.tn * WM, fc.kb0 + fc.kb, ring[j]); cur_next(fc); }
    }
    int stage = 0; uint32_t phase = 0;
    while (cc.ok) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        if (cc.ok) {
          if (fc.ok) { fetch(fc.tn * WM, fc.kb0 + fc.kb, ring[(j + PF - 1) % PF]); cur_next(fc); }
          tc::mbar_wait(&empty[stage], phase ^ 1);
          convert(sA + stage * WA_BYTES, ring[j], cc.tn * WM, cc.kb0 + cc.kb);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&full[stage]);      // one arrival per producer warp
          if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
          cur_next(cc);
        }
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

extern "C" {

// C[M,N] fp32 (+)= X[M,K] (bf16, K contiguous, ld = ldx) * W[N,K]^T with W given by its packed planes
// (block 32 along K, inner layout, no hi stream: inlier and outlier codes share one width).
int msq_gemm_packed_w_bf16(const void* X, long long ldx, const unsigned char* codes, const unsigned int* bitmap,
                           const signed char* e_in, const signed char* e_out, float* C, long long ldc, int M, int N, int K,
                           int in_ebits, int in_mbits, int out_ebits, int out_mbits, int plain, int accumulate,
                           void* stream) {
  msq::DevGuard dev_guard_(X);
  if (!X || !codes || !e_in || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if (!plain && (!bitmap || !e_out)) return MSQ_ERR_BAD_ARG;
  if (K % 32 != 0) return MSQ_ERR_UNSUPPORTED;
  Fmt fi = make_fmt(in_ebits, in_mbits, 0, 1.0f), fo = make_fmt(plain ? in_ebits : out_ebits, plain ? in_mbits : out_mbits, 0, 1.0f);
  if (fi.width != fo.width) return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(X) & 15) || (ldx % 8) || (reinterpret_cast<uintptr_t>(codes) & 15)) return MSQ_ERR_ALIGN;
  W4Params P;
  memset(&P, 0, sizeof(P));
  P.codes = codes; P.bitmap = bitmap; P.e_in = e_in; P.e_out = e_out; P.C = C; P.ldc = ldc; P.M = M; P.N = N; P.K = K;
  int bnx = M >= WMAXN ? WMAXN : ((M + 15) / 16) * 16;
  P.bnx = bnx;
  P.tiles_n = (N + WM - 1) / WM; P.tiles_m = (M + bnx - 1) / bnx;
  P.wi = fi.width; P.in_ebits = in_ebits; P.in_bits = in_mbits;
  P.out_ebits = plain ? in_ebits : out_ebits; P.out_bits = plain ? in_mbits : out_mbits;
  P.plain = plain; P.accumulate = accumulate;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // split K until the grid covers the SMs (each slice keeps at least 8 K-blocks of 64)
  const int kblocks_total = (K + WBK - 1) / WBK;
  // split K when the un-split grid cannot fill the machine: cost = rounds over the SMs x (K-blocks per item + a
  // fixed per-item charge for the TMEM drain and the atomics)
  int ksplit = 1;
  {
    const int tiles = P.tiles_n * P.tiles_m;
    double best = 1e30;
    for (int ks = 1; ks <= (tiles < sms ? 16 : 1); ks *= 2) {
      const int per = (kblocks_total + ks - 1) / ks;
      if (ks > 1 && (per < 4 || (ks - 1) * per >= kblocks_total)) break;      // every split must own K-blocks
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
  // the opt-in is per device: set it on every launch (microseconds) instead of once per process
  if (cudaFuncSetAttribute(k_gemm_w4, cudaFuncAttributeMaxDynamicSharedMemorySize, W4_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
  int grid = P.tiles_n * P.tiles_m * P.ksplit;
  if (grid > sms) grid = sms;
  k_gemm_w4<<<grid, 512, W4_SMEM, (cudaStream_t)stream>>>(tmX, P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"
