// msq_gemm_split3.cu -- fp32-accurate contractions on tcgen05 for the GPTQ solver (llm/gptq.py:98-165):
//
//     C[M,N] (fp32) = alpha * A[M,K] * B[N,K]^T + beta * C
//
// with A and B given as THREE bf16 planes each whose sum is the fp32 operand exactly (x = p0 + p1 + p2,
// p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1)).  The six plane pairs whose product is not below
// 2^-24 of the leading one -- (0,0) (0,1) (0,2) (1,0) (1,1) (2,0) -- are accumulated in the fp32 TMEM
// accumulator, every bf16 x bf16 product being exact there, so the result carries fp32-level error (the dropped
// pairs are <= 3 * 2^-24 |a||b|) while running on the tensor cores.  Used for
//   * the trailing update  W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:]        (llm/gptq.py:165)
//   * the blocked Cholesky / triangular-inverse recursion that replaces the cuSOLVER trio (llm/gptq.py:101-103)
//
// Structure = k_gemm_nt (msq_gemm.cu): TMA producer warp, single-thread tcgen05.mma issuer, TMEM double-buffered
// 128x256 fp32 accumulators, 4 epilogue warps, persistent round-robin tiles.  Differences: the K loop walks
// (k-block, plane pair) units; operands are addressed as sub-views of plane stacks through ONE tensor map per
// stack (plane p of a stack starts p * plane_rows rows further down); triangular operands restrict the K range per
// tile; the epilogue can emit the result as bf16 planes (row-major and/or transposed) for the next contraction.
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_tc.cuh"
#include "msq_split3.cuh"

namespace msq {

namespace s3 {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2, B_STAGE_BYTES = BN * BK * 2;
constexpr int SMEM = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + 256;

__device__ __forceinline__ void split3_store(float v, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
  p0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r1);
  p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
}

__global__ void __launch_bounds__(256, 1)
k_gemm_split3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Split3Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint64_t* acc_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (P.M + BM - 1) / BM, tiles_n = (P.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int kblocks = (P.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc<512>(tmem_base_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  auto tile_coords = [&](int t, int& tm, int& tn) { tn = t / tiles_m; tm = t - tn * tiles_m; };
  auto tile_skipped = [&](int tm, int tn) { return P.upper_only && (tm * BM >= (tn + 1) * BN); };
  // K range of a tile when one operand is triangular (zero blocks are never loaded)
  auto k_range = [&](int tm, int tn, int& kb0, int& kb1) {
    kb0 = 0; kb1 = kblocks;
    if (P.k_lo_mode == 1) kb0 = (tm * BM) / BK;
    else if (P.k_lo_mode == 2) kb0 = (tn * BN) / BK;
    if (P.k_hi_mode == 1) kb1 = min(kblocks, ((tm + 1) * BM + BK - 1) / BK);
    if (kb0 > kb1 - 1) kb0 = max(kb1 - 1, 0);
  };
  constexpr int npairs = 6;      // plane pairs (pa, pb) in accumulation order

  // role loops on the whole warp, TMA / tcgen05 instructions under tc::elect_one() (see msq_tc.cuh)
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn; tile_coords(t, tm, tn);
      if (tile_skipped(tm, tn)) continue;
      int kb0, kb1; k_range(tm, tn, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        for (int pr = 0; pr < npairs; ++pr) {
          const int pa = (0x211000 >> (4 * pr)) & 0xF;        // 0,0,0,1,1,2
          const int pb = (0x010210 >> (4 * pr)) & 0xF;        // 0,1,2,0,1,0
          tc::mbar_wait(&empty[stage], phase ^ 1);
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
            tc::tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], P.a_col0 + kb * BK,
                            P.a_row0 + pa * P.a_plane_rows + tm * BM);
            tc::tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, &full[stage], P.b_col0 + kb * BK,
                            P.b_row0 + pb * P.b_plane_rows + tn * BN);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = tc::make_idesc(1 /*bf16*/, BM, BN);
    const uint64_t ad0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sA)), bd0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sB));
    int stage = 0; uint32_t phase = 0;
    int buf = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn; tile_coords(t, tm, tn);
      if (tile_skipped(tm, tn)) continue;
      int kb0, kb1; k_range(tm, tn, kb0, kb1);
      tc::mbar_wait(&acc_empty[buf], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(buf * BN);
      uint32_t first = 1;
      for (int u = (kb1 - kb0) * npairs; u > 0; --u) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)stage * (uint32_t)(A_STAGE_BYTES >> 4));
        const uint64_t bd = bd0 + (uint64_t)((uint32_t)stage * (uint32_t)(B_STAGE_BYTES >> 4));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_f16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
          tc::umma_commit(&empty[stage]);
        }
        __syncwarp();
        first = 0;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&acc_full[buf]);
      __syncwarp();
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int buf = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn; tile_coords(t, tm, tn);
      if (tile_skipped(tm, tn)) continue;
      tc::mbar_wait(&acc_full[buf], acc_phase);
      tc::tc_fence_after();
      const int row = tm * BM + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32), r);
        tc::tmem_ld_wait();
        const int col0 = tn * BN + c * 32;
        if (col0 >= P.N) continue;                       // warp-uniform
        const bool full32 = col0 + 32 <= P.N && ((P.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);    // float4 rows need a 16-byte row pitch
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = P.alpha * __uint_as_float(r[j]);
        if (row < P.M) {
          if (P.C) {
            float* crow = P.C + (long long)row * P.ldc + col0;
            if (full32) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4* dst = reinterpret_cast<float4*>(crow + 4 * j);
                if (P.beta != 0.0f) {
                  const float4 old = *dst;
                  v[4 * j + 0] = fmaf(P.beta, old.x, v[4 * j + 0]); v[4 * j + 1] = fmaf(P.beta, old.y, v[4 * j + 1]);
                  v[4 * j + 2] = fmaf(P.beta, old.z, v[4 * j + 2]); v[4 * j + 3] = fmaf(P.beta, old.w, v[4 * j + 3]);
                }
                *dst = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (col0 + j < P.N) {
                  if (P.beta != 0.0f) v[j] = fmaf(P.beta, crow[j], v[j]);
                  crow[j] = v[j];
                }
              }
            }
          }
          if (P.out_planes) {      // row-major bf16 planes of the result (operand of a later contraction)
            __nv_bfloat16* p0 = P.out_planes + (long long)(P.op_row0 + row) * P.ldp + P.op_col0 + col0;
            const long long ps = (long long)P.op_plane_rows * P.ldp;
            if (col0 + 32 <= P.N) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                __align__(16) __nv_bfloat16 a0[8], a1[8], a2[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split3_store(v[8 * j + e], a0[e], a1[e], a2[e]);
                *reinterpret_cast<uint4*>(p0 + 8 * j) = *reinterpret_cast<const uint4*>(a0);
                *reinterpret_cast<uint4*>(p0 + ps + 8 * j) = *reinterpret_cast<const uint4*>(a1);
                *reinterpret_cast<uint4*>(p0 + 2 * ps + 8 * j) = *reinterpret_cast<const uint4*>(a2);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (col0 + j < P.N) {
                  __nv_bfloat16 b0, b1, b2;
                  split3_store(v[j], b0, b1, b2);
                  p0[j] = b0; p0[ps + j] = b1; p0[2 * ps + j] = b2;
                }
              }
            }
          }
          if (P.out_planes_t) {    // transposed planes: element (row, col) lands at [col][row]; lanes = consecutive rows
            const long long ps = (long long)P.opt_plane_rows * P.ldp;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < P.N) {
                __nv_bfloat16 b0, b1, b2;
                split3_store(v[j], b0, b1, b2);
                __nv_bfloat16* pt = P.out_planes_t + (long long)(P.opt_row0 + col0 + j) * P.ldp + P.opt_col0 + row;
                pt[0] = b0; pt[ps] = b1; pt[2 * ps] = b2;
              }
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

// fp32 [rows, cols] (ld) -> three bf16 planes at (row0, col0) of a plane stack; columns up to the next multiple of 64
// are zero-filled (the stack's K range is read in 64-element blocks).
__global__ void __launch_bounds__(256) k_split3(const float* __restrict__ src, long long ld, int rows, int cols,
                                                __nv_bfloat16* __restrict__ planes, long long ldp, int plane_rows, int row0,
                                                int col0) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, cc = blockIdx.x * 32 + tx;
  const long long ps = (long long)plane_rows * ldp;
  for (int r = ty; r < 32; r += 8) {
    const int rr = r0 + r;
    if (rr >= rows) continue;
    const float v = cc < cols ? src[(long long)rr * ld + cc] : 0.0f;
    __nv_bfloat16 b0, b1, b2;
    split3_store(v, b0, b1, b2);
    __nv_bfloat16* p = planes + (long long)(row0 + rr) * ldp + col0 + cc;
    p[0] = b0; p[ps] = b1; p[2 * ps] = b2;
  }
}

}  // namespace s3

int split3_make_tmap(CUtensorMap* out, const __nv_bfloat16* stack, long long stack_rows, long long cols, long long ldp,
                     int box_rows) {
  return make_tmap_2d(out, stack, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, stack_rows, cols, ldp, box_rows, s3::BK);
}

static int sm_count_s3() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

int split3_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const Split3Params& P, cudaStream_t stream) {
  if (P.M <= 0 || P.N <= 0 || P.K <= 0) return MSQ_OK;
  if (P.C && (reinterpret_cast<uintptr_t>(P.C) & 3)) return MSQ_ERR_ALIGN;
  if (cudaFuncSetAttribute(s3::k_gemm_split3, cudaFuncAttributeMaxDynamicSharedMemorySize, s3::SMEM) != cudaSuccess)
    return MSQ_ERR_CUDA;
  const int tiles = ((P.M + s3::BM - 1) / s3::BM) * ((P.N + s3::BN - 1) / s3::BN);
  static int sms = 0;
  if (!sms) sms = sm_count_s3();
  const int grid = tiles < sms ? tiles : sms;
  s3::k_gemm_split3<<<grid, 256, s3::SMEM, stream>>>(tmA, tmB, P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int split3_convert(const float* src, long long ld, int rows, int cols, __nv_bfloat16* planes, long long ldp, int plane_rows,
                   int row0, int col0, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return MSQ_OK;
  const int cpad = (cols + 63) / 64 * 64;
  dim3 g(cpad / 32, (rows + 31) / 32);
  s3::k_split3<<<g, 256, 0, stream>>>(src, ld, rows, cols, planes, ldp, plane_rows, row0, col0);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // namespace msq
