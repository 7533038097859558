// msq_gemm.cu -- persistent warp-specialized tcgen05 GEMM for sm_100a:
//     C[M,N] (fp32) = alpha * A[M,K] * B[N,K]^T + beta * C         (A, B fp16 or bf16, K contiguous)
//
// Used for the GPTQ Hessian (replaces `inp.matmul(inp.t())`, llm/gptq.py:58, with A = B = X^T)
// and for the packed-weight linear at large M.
//
//   warp 0     : TMA producer   (cp.async.bulk.tensor, 128B-swizzled tiles, STAGES-deep mbarrier ring)
//   warp 1     : MMA issuer     (one elected thread, tcgen05.mma.cta_group::1.kind::f16, 128x256x16)
//   warp 2     : TMEM allocator (512 columns = two 128x256 fp32 accumulators, double buffered)
//   warps 4..7 : epilogue       (tcgen05.ld 32x32b -> registers -> alpha/beta -> global)
// One CTA per SM, static round-robin tile schedule; the epilogue of tile i overlaps the MMAs of tile i+1.
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_tc.cuh"
#include "msq_split3.cuh"

namespace msq {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2, B_STAGE_BYTES = BN * BK * 2;
constexpr int GEMM_SMEM = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  float* C;
  long long ldc;
  int M, N, K;
  float alpha, beta;
  int ab_format;      // 0 fp16, 1 bf16
  int tiles_m, tiles_n;
  int upper_only;     // symmetric product: skip tiles entirely below the diagonal
};

__global__ void __launch_bounds__(256, 1)
k_gemm_nt(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;      // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2; // [2] epilogue -> MMA
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = P.tiles_m * P.tiles_n;
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

  auto tile_coords = [&](int t, int& tm, int& tn) { tn = t / P.tiles_m; tm = t - tn * P.tiles_m; };
  auto tile_skipped = [&](int tm, int tn) { return P.upper_only && (tm * BM >= (tn + 1) * BN); };

  // role loops on the whole warp, TMA / tcgen05 instructions under tc::elect_one() (see msq_tc.cuh)
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn; tile_coords(t, tm, tn);
      if (tile_skipped(tm, tn)) continue;
      for (int kb = 0; kb < kblocks; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          tc::tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], kb * BK, tm * BM);
          tc::tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, &full[stage], kb * BK, tn * BN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = tc::make_idesc(P.ab_format, BM, BN);
    const uint64_t ad0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sA)), bd0 = tc::make_kmajor_sw128_desc(tc::smem_u32(sB));
    int stage = 0; uint32_t phase = 0;
    int buf = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn; tile_coords(t, tm, tn);
      if (tile_skipped(tm, tn)) continue;
      tc::mbar_wait(&acc_empty[buf], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(buf * BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        // descriptor = base + stage * (stage bytes >> 4): the 14-bit address field cannot carry (shared memory < 256 KB)
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)stage * (uint32_t)(A_STAGE_BYTES >> 4));
        const uint64_t bd = bd0 + (uint64_t)((uint32_t)stage * (uint32_t)(B_STAGE_BYTES >> 4));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)      // 16 elements = 32 bytes = 2 descriptor address units per UMMA_K
            tc::umma_f16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          tc::umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&acc_full[buf]);
      __syncwarp();
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    int buf = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn; tile_coords(t, tm, tn);
      if (tile_skipped(tm, tn)) continue;
      tc::mbar_wait(&acc_full[buf], acc_phase);
      tc::tc_fence_after();
      const int row = tm * BM + q * 32 + lane;
      float* crow = P.C + (long long)row * P.ldc + (long long)tn * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32), r);
        tc::tmem_ld_wait();
        const int col0 = tn * BN + c * 32;
        if (row < P.M) {
          if (col0 + 32 <= P.N && ((P.ldc & 3) == 0)) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 o;
              o.x = P.alpha * __uint_as_float(r[4 * v + 0]);
              o.y = P.alpha * __uint_as_float(r[4 * v + 1]);
              o.z = P.alpha * __uint_as_float(r[4 * v + 2]);
              o.w = P.alpha * __uint_as_float(r[4 * v + 3]);
              float4* dst = reinterpret_cast<float4*>(crow + c * 32 + 4 * v);
              if (P.beta != 0.0f) {
                const float4 old = *dst;
                o.x = fmaf(P.beta, old.x, o.x); o.y = fmaf(P.beta, old.y, o.y);
                o.z = fmaf(P.beta, old.z, o.z); o.w = fmaf(P.beta, old.w, o.w);
              }
              *dst = o;
            }
          } else {
            for (int v = 0; v < 32; ++v) {
              if (col0 + v < P.N) {
                float o = P.alpha * __uint_as_float(r[v]);
                if (P.beta != 0.0f) o = fmaf(P.beta, crow[c * 32 + v], o);
                crow[c * 32 + v] = o;
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

// mirror the strictly-upper part of a symmetric fp32 matrix into the lower part (32x32 smem tiles)
__global__ void __launch_bounds__(256) k_symmetrize_lower(float* __restrict__ H, int n, long long ld) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x, by = blockIdx.y;     // tile (by, bx) of the upper part, bx >= by
  if (bx < by) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int row = by * 32 + r, col = bx * 32 + tx;
    tile[r][tx] = (row < n && col < n) ? H[(long long)row * ld + col] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int row = bx * 32 + r, col = by * 32 + tx;      // transposed position
    if (row < n && col < n && row > col) H[(long long)row * ld + col] = tile[tx][r];
  }
}

// X [T, K] (fp16/bf16/fp32) -> Xt [K, Tpad] 16-bit, zero padded to Tpad; fp32 input is split into
// hi = fp16(x), lo = fp16(x - hi) and emitted as the two operand streams
//   XtA = [hi | hi | lo],  XtB = [hi | lo | hi]   (so  XtA * XtB^T = hi hi^T + hi lo^T + lo hi^T)
template <typename TIN>
__global__ void __launch_bounds__(256) k_transpose_to16(const TIN* __restrict__ X, uint16_t* __restrict__ XtA,
                                                        uint16_t* __restrict__ XtB, int T, int K, long long ldx, int Tpad,
                                                        int split) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, k = k0 + tx;
    float v = 0.0f;
    if (t < T && k < K) {
      if (sizeof(TIN) == 4) v = reinterpret_cast<const float*>(X)[(long long)t * ldx + k];
      else if (split == 2) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(X)[(long long)t * ldx + k]);
      else v = __half2float(reinterpret_cast<const __half*>(X)[(long long)t * ldx + k]);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, t = t0 + tx;
    if (k >= K || t >= Tpad) continue;
    const float v = tile[tx][r];
    const long long row = (long long)k * (split == 1 ? 3LL * Tpad : Tpad);
    if (split == 1) {          // fp32 input, two-term fp16 split
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const uint16_t h = __half_as_ushort(hi), l = __half_as_ushort(lo);
      XtA[row + t] = h; XtA[row + Tpad + t] = h; XtA[row + 2LL * Tpad + t] = l;
      XtB[row + t] = h; XtB[row + Tpad + t] = l; XtB[row + 2LL * Tpad + t] = h;
    } else if (split == 2) {
      XtA[row + t] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    } else {
      XtA[row + t] = __half_as_ushort(__float2half_rn(v));
    }
  }
}

// X [T, K] fp32 -> three bf16 planes of X^T, each [Kpad, Tpad] (x = p0 + p1 + p2 exactly, zero padded): the operand
// of the split-plane GEMM.  bf16 keeps fp32's exponent range, so -- unlike a two-term fp16 split -- nothing
// overflows at |x| >= 65504 and small features keep their relative accuracy.
__global__ void __launch_bounds__(256) k_transpose_split3(const float* __restrict__ X, __nv_bfloat16* __restrict__ P, int T,
                                                          int K, long long ldx, int Tpad, int Kpad) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, k = k0 + tx;
    tile[r][tx] = (t < T && k < K) ? X[(long long)t * ldx + k] : 0.0f;
  }
  __syncthreads();
  const long long ps = (long long)Kpad * Tpad;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, t = t0 + tx;
    if (k >= Kpad || t >= Tpad) continue;
    const float v = tile[tx][r];
    const __nv_bfloat16 p0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(p0);
    const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
    __nv_bfloat16* dst = P + (long long)k * Tpad + t;
    dst[0] = p0; dst[ps] = p1; dst[2 * ps] = p2;
  }
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes, long long rows, long long cols,
                 long long ld, int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return MSQ_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MSQ_OK : MSQ_ERR_CUDA;
}

// packed 4-bit matrix [rows, cols elements] (row pitch in bytes) loaded as CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B: every 16 codes
// (8 bytes in global memory) land in a 16-byte chunk of shared memory -- the layout tcgen05.mma kind::mxf8f6f4 expects
// for 4-bit operands.  box = [box_rows, 128 elements]; cols % 128 == 0, base 32-byte aligned, pitch % 32 == 0.
int make_tmap_2d_u4(CUtensorMap* out, const void* base, long long rows, long long cols, long long row_bytes, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return MSQ_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MSQ_OK : MSQ_ERR_CUDA;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int gemm_nt_16(const void* A, const void* B, float* C, int M, int N, int K, long long lda, long long ldb,
                      long long ldc, float alpha, float beta, int ab_format, int upper_only, cudaStream_t stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return MSQ_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15) ||
      (lda % 8) || (ldb % 8))
    return MSQ_ERR_ALIGN;
  CUtensorMap tmA, tmB;
  const CUtensorMapDataType dt = ab_format == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  int rc = make_tmap_2d(&tmA, A, dt, 2, M, K, lda, BM, BK);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, B, dt, 2, N, K, ldb, BN, BK);
  if (rc) return rc;
  GemmParams P;
  P.C = C; P.ldc = ldc; P.M = M; P.N = N; P.K = K; P.alpha = alpha; P.beta = beta; P.ab_format = ab_format;
  P.tiles_m = (M + BM - 1) / BM; P.tiles_n = (N + BN - 1) / BN; P.upper_only = upper_only;
  // the opt-in is per device: set it on every launch (microseconds) instead of once per process
  if (cudaFuncSetAttribute(k_gemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) != cudaSuccess) return MSQ_ERR_CUDA;
  int grid = P.tiles_m * P.tiles_n;
  if (grid > sm_count()) grid = sm_count();
  k_gemm_nt<<<grid, 256, GEMM_SMEM, stream>>>(tmA, tmB, P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // namespace msq

using namespace msq;

extern "C" {

// C[M,N] fp32 = alpha * A[M,K] * B[N,K]^T + beta * C ; A, B 16-bit (dtype 1 = fp16, 2 = bf16), K contiguous
int msq_gemm_nt_f32out(const void* A, const void* B, float* C, int M, int N, int K, long long lda, long long ldb,
                       long long ldc, float alpha, float beta, int dtype, void* stream) {
  msq::DevGuard dev_guard_(A);
  if (dtype != 1 && dtype != 2) return MSQ_ERR_BAD_ARG;
  return gemm_nt_16(A, B, C, M, N, K, lda, ldb, ldc, alpha, beta, dtype == 2 ? 1 : 0, 0, (cudaStream_t)stream);
}

// workspace bytes for msq_hessian_accumulate
long long msq_hessian_workspace_bytes(int T, int K, int dtype) {
  const long long Tpad = ((long long)T + 63) / 64 * 64;
  const long long Kpad = ((long long)K + 127) / 128 * 128;
  return dtype == 0 ? 3LL * Kpad * Tpad * 2 : Tpad * (long long)K * 2;
}

// GPTQ.add_batch (llm/gptq.py:53-58) for nn.Linear:   H <- beta * H + alpha * X^T X
// with X [T, K] row-major tokens x features (dtype 0 fp32, 1 fp16, 2 bf16), H [K, K] fp32 (ld = ldh).
// The caller passes beta = n/(n+b), alpha = 2/(n+b).  fp16/bf16 products are exact in the fp32
// accumulator; fp32 input is split into three bf16 planes (full fp32 range, error ~2^-24 |x||x|^T per product).
// symmetric != 0: only tiles touching the upper triangle are computed, then mirrored.
int msq_hessian_accumulate(const void* X, int T, int K, long long ldx, int dtype, float* H, long long ldh, float alpha,
                           float beta, void* workspace, long long workspace_bytes, int symmetric, void* stream) {
  msq::DevGuard dev_guard_(X);
  if (!X || !H || !workspace || T <= 0 || K <= 0 || dtype < 0 || dtype > 2) return MSQ_ERR_BAD_ARG;
  if (workspace_bytes < msq_hessian_workspace_bytes(T, K, dtype)) return MSQ_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int Tpad = (T + 63) / 64 * 64;
  uint16_t* XtA = reinterpret_cast<uint16_t*>(workspace);
  uint16_t* XtB = XtA;
  dim3 g((Tpad + 31) / 32, (K + 31) / 32);
  int Keff = Tpad;
  if (dtype == 0) {
    // fp32 activations: three bf16 planes of X^T and the fp32-accurate split-plane GEMM (six plane pairs)
    const int Kpad = (K + 127) / 128 * 128;
    __nv_bfloat16* PL = reinterpret_cast<__nv_bfloat16*>(workspace);
    dim3 g3((Tpad + 31) / 32, (Kpad + 31) / 32);
    k_transpose_split3<<<g3, 256, 0, s>>>(reinterpret_cast<const float*>(X), PL, T, K, ldx, Tpad, Kpad);
    if (cudaGetLastError() != cudaSuccess) return MSQ_ERR_CUDA;
    CUtensorMap mA, mB;
    if (split3_make_tmap(&mA, PL, 3LL * Kpad, Tpad, Tpad, 128)) return MSQ_ERR_CUDA;
    if (split3_make_tmap(&mB, PL, 3LL * Kpad, Tpad, Tpad, 256)) return MSQ_ERR_CUDA;
    Split3Params SP;
    memset(&SP, 0, sizeof(SP));
    SP.M = K; SP.N = K; SP.K = Tpad;
    SP.a_plane_rows = Kpad; SP.b_plane_rows = Kpad;
    SP.alpha = alpha; SP.beta = beta; SP.C = H; SP.ldc = ldh;
    SP.upper_only = symmetric ? 1 : 0;
    int rc3 = split3_gemm(mA, mB, SP, s);
    if (rc3) return rc3;
    if (symmetric) {
      dim3 gs((K + 31) / 32, (K + 31) / 32);
      k_symmetrize_lower<<<gs, 256, 0, s>>>(H, K, ldh);
      if (cudaGetLastError() != cudaSuccess) return MSQ_ERR_CUDA;
    }
    return MSQ_OK;
  } else {
    k_transpose_to16<uint16_t><<<g, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(X), XtA, XtB, T, K, ldx, Tpad, dtype == 2 ? 2 : 0);
  }
  if (cudaGetLastError() != cudaSuccess) return MSQ_ERR_CUDA;
  // the fp32 split is not symmetric term by term (hi lo^T + lo hi^T is, as a sum) -> still symmetric overall
  int rc = gemm_nt_16(XtA, XtB, H, K, K, Keff, Keff, Keff, ldh, alpha, beta, dtype == 2 ? 1 : 0, symmetric ? 1 : 0, s);
  if (rc) return rc;
  if (symmetric) {
    dim3 gs((K + 31) / 32, (K + 31) / 32);
    k_symmetrize_lower<<<gs, 256, 0, s>>>(H, K, ldh);
    if (cudaGetLastError() != cudaSuccess) return MSQ_ERR_CUDA;
  }
  return MSQ_OK;
}

}  // extern "C"
