// msq_cholinv.cu -- the "Cholesky trio" of GPTQ.fasterquant (llm/gptq.py:101-103)
//
//     H = cholesky(H);  H = cholesky_inverse(H);  H = cholesky(H, upper=True)        ->  Hinv
//
// as ONE native routine.  The result is the upper-triangular U with U^T U = H^-1.  Algebra: factor H = R R^T with R
// UPPER triangular (a Cholesky that starts in the bottom-right corner); then H^-1 = R^-T R^-1 = (R^-1)^T (R^-1) and
// U = R^-1 is upper triangular with a positive diagonal -- by uniqueness of the Cholesky factor it IS the matrix the
// three library calls produce (1 factorization + 1 triangular inverse = 2/3 n^3 flops instead of 4/3 n^3).
//
// Recursion over 128-column blocks ("2" = bottom-right part, handled first):
//     (R22, U22) = f(A22)
//     R12 = A12 U22^T                      tcgen05, fp32-accurate split-plane GEMM (msq_gemm_split3.cu)
//     A11 -= R12 R12^T                     tcgen05, upper-triangle tiles only
//     (R11, U11) = f(A11)
//     U12 = -U11 (R12 U22)                 tcgen05 x2, triangular operands restrict the K range per tile
// Leaves (<= 128 x 128) run in one CTA: reverse Cholesky + inverse of 32 x 32 diagonal sub-blocks in the registers
// of one warp (shuffle broadcasts), panel / trailing products from shared memory.  Every product the recursion needs
// is kept as bf16 plane stacks (PU, PU^T, PR); PU^T doubles as the operand of the GPTQ trailing update.
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include "msq_tc.cuh"
#include "msq_split3.cuh"

namespace msq {

namespace ci {

constexpr int LB = 128;        // leaf size

__device__ __forceinline__ void split3(float v, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
  p0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r1);
  p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
}

// One warp: A (32 x 32 SPD, upper triangle in a[], lane = row) -> R upper with A = R R^T (in a[]) and u[] = R^-1.
// returns 0 or 1 + index of the first non-positive pivot
__device__ __forceinline__ int warp_chol_inv_32(float (&a)[32], float (&u)[32], int lane) {
  int bad = 0;
#pragma unroll
  for (int j = 31; j >= 0; --j) {
    const float ajj = __shfl_sync(0xffffffffu, a[j], j);
    if (!(ajj > 0.0f) && bad == 0) bad = j + 1;
    const float rjj = sqrtf(ajj);
    const float rij = lane < j ? a[j] / rjj : (lane == j ? rjj : 0.0f);
    a[j] = rij;
#pragma unroll
    for (int k = 0; k < j; ++k) {
      const float rkj = __shfl_sync(0xffffffffu, rij, k);
      a[k] = fmaf(-rij, rkj, a[k]);           // A[i][k] -= R[i][j] R[k][j]   (used for i <= k only)
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float rjj = __shfl_sync(0xffffffffu, a[j], j);
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int k = 0; k < j; ++k) {
      const float rkj = __shfl_sync(0xffffffffu, a[j], k);
      if (k & 1) s1 = fmaf(u[k], rkj, s1); else s0 = fmaf(u[k], rkj, s0);
    }
    u[j] = lane == j ? 1.0f / rjj : (lane < j ? -(s0 + s1) / rjj : 0.0f);
  }
  return bad;
}

struct LeafParams {
  const float* A; long long lda;     // matrix being factored (upper triangle valid)
  float* U; long long ldu;           // result matrix
  __nv_bfloat16* PU; __nv_bfloat16* PUT; long long ldp; int plane_rows;
  int o, m;                          // block offset and size (m <= 128)
  int* info;
};

// acc[a][j] = sum_{k < K} A[(tr + 32 a) * lda + k] * B[(tc + 8 j) * ldb + k]     (tr = tid / 8, tc = tid % 8)
// A . B^T from shared memory with both operands contiguous in k: 7 LDS.128 feed 48 FMAs per 4 k; the row -> thread
// maps keep the 16-byte accesses of a quarter warp on distinct banks (row pitch 132 words).
__device__ __forceinline__ void nt_tile(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int M,
                                        int K, int tr, int tc, float (&acc)[3][4]) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[a][j] = 0.0f;
  const bool r1 = tr + 32 < M, r2 = tr + 64 < M;
  if (tr >= M) return;
  for (int k = 0; k < K; k += 4) {
    float4 av[3], bv[4];
    av[0] = *reinterpret_cast<const float4*>(A + tr * lda + k);
    av[1] = r1 ? *reinterpret_cast<const float4*>(A + (tr + 32) * lda + k) : make_float4(0, 0, 0, 0);
    av[2] = r2 ? *reinterpret_cast<const float4*>(A + (tr + 64) * lda + k) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(B + (tc + 8 * j) * ldb + k);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[a][j] = fmaf(av[a].x, bv[j].x, acc[a][j]);
        acc[a][j] = fmaf(av[a].y, bv[j].y, acc[a][j]);
        acc[a][j] = fmaf(av[a].z, bv[j].z, acc[a][j]);
        acc[a][j] = fmaf(av[a].w, bv[j].w, acc[a][j]);
      }
  }
}

constexpr int UTS = 36;    // row pitch of the transposed 32 x 32 diagonal inverse
constexpr int TTS = 100;   // row pitch of T^T (32 x <= 96)
constexpr int LEAF_SMEM2 = (2 * LB * 132 + 32 * UTS + 32 * TTS) * (int)sizeof(float);

__global__ void __launch_bounds__(256, 1) k_chol_leaf(const LeafParams P) {
  extern __shared__ __align__(16) float sm_leaf[];
  constexpr int S = 132;             // row pitch: 16-byte aligned rows
  float* As = sm_leaf;               // [128][S]  A -> R
  float* Us = sm_leaf + LB * S;      // [128][S]  U = R^-1
  float* UTd = Us + LB * S;          // [32][UTS] transpose of the current diagonal block of U
  float* TT = UTd + 32 * UTS;        // [32][TTS] (R12 U22)^T
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tr = tid >> 3, tc = tid & 7;
  const int m = P.m;
  for (int idx = tid; idx < LB * LB; idx += 256) {
    const int i = idx >> 7, k = idx & 127;
    float v = 0.0f;
    if (i < m && k < m && i <= k) v = P.A[(long long)(P.o + i) * P.lda + P.o + k];
    else if (i == k) v = 1.0f;                           // identity padding
    As[i * S + k] = v;
    Us[i * S + k] = 0.0f;
  }
  __syncthreads();

  // phase 1: A = R R^T from the last 32-column block to the first; diagonal blocks are inverted on the way
  for (int jb = 3; jb >= 0; --jb) {
    const int c0 = jb * 32;
    if (warp == 0) {
      float a[32], u[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) { a[k] = k >= lane ? As[(c0 + lane) * S + c0 + k] : 0.0f; u[k] = 0.0f; }
      const int bad = warp_chol_inv_32(a, u, lane);
      if (bad && lane == 0 && c0 + bad - 1 < m) atomicCAS(P.info, 0, P.o + c0 + bad);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        As[(c0 + lane) * S + c0 + k] = k >= lane ? a[k] : 0.0f;
        Us[(c0 + lane) * S + c0 + k] = u[k];
      }
    }
    __syncthreads();
    if (c0 == 0) break;
    // panel: R12 = A12 U22^T   ([c0 x 32], K = 32), buffered in registers (it overwrites its own A operand)
    float acc[3][4];
    nt_tile(As + c0, S, Us + c0 * S + c0, S, c0, 32, tr, tc, acc);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (tr + 32 * a < c0)
#pragma unroll
        for (int j = 0; j < 4; ++j) As[(tr + 32 * a) * S + c0 + tc + 8 * j] = acc[a][j];
    __syncthreads();
    // trailing: A11 -= R12 R12^T, 32-column groups on or above the diagonal
    for (int g = 0; g < jb; ++g) {
      nt_tile(As + c0, S, As + (32 * g) * S + c0, S, 32 * (g + 1), 32, tr, tc, acc);
#pragma unroll
      for (int a = 0; a < 3; ++a)
        if (tr + 32 * a < 32 * (g + 1))
#pragma unroll
          for (int j = 0; j < 4; ++j) As[(tr + 32 * a) * S + 32 * g + tc + 8 * j] -= acc[a][j];
    }
    __syncthreads();
  }
  // phase 2: off-diagonal blocks of U = R^-1:  U12 = -U11 (R12 U22), block column by block column
  for (int jb = 1; jb < 4; ++jb) {
    const int c0 = jb * 32;
    for (int idx = tid; idx < 32 * 32; idx += 256) {     // transpose of the diagonal inverse: K-contiguous B operand
      const int r = idx >> 5, c = idx & 31;
      UTd[c * UTS + r] = Us[(c0 + r) * S + c0 + c];
    }
    __syncthreads();
    float acc[3][4];
    nt_tile(As + c0, S, UTd, UTS, c0, 32, tr, tc, acc);  // T = R12 U22, stored transposed
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (tr + 32 * a < c0)
#pragma unroll
        for (int j = 0; j < 4; ++j) TT[(tc + 8 * j) * TTS + tr + 32 * a] = acc[a][j];
    __syncthreads();
    nt_tile(Us, S, TT, TTS, c0, c0, tr, tc, acc);        // U12 = -U11 T   (K = c0; U11 is upper triangular)
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (tr + 32 * a < c0)
#pragma unroll
        for (int j = 0; j < 4; ++j) Us[(tr + 32 * a) * S + c0 + tc + 8 * j] = -acc[a][j];
    __syncthreads();
  }
  // phase 3: results -- fp32 block of U, planes of U and of U^T (full block, zeros below the diagonal)
  const long long ps = (long long)P.plane_rows * P.ldp;
  for (int idx = tid; idx < LB * LB; idx += 256) {
    const int i = idx >> 7, k = idx & 127;
    if (i >= m || k >= m) continue;
    const float v = Us[i * S + k];                       // 0 for i > k
    P.U[(long long)(P.o + i) * P.ldu + P.o + k] = v;
    __nv_bfloat16 b0, b1, b2;
    split3(v, b0, b1, b2);
    __nv_bfloat16* p = P.PU + (long long)(P.o + i) * P.ldp + P.o + k;
    p[0] = b0; p[ps] = b1; p[2 * ps] = b2;
    const float vt = Us[k * S + i];                      // U^T[i][k]
    split3(vt, b0, b1, b2);
    __nv_bfloat16* q = P.PUT + (long long)(P.o + i) * P.ldp + P.o + k;
    q[0] = b0; q[ps] = b1; q[2 * ps] = b2;
  }
}

struct Layout {
  int n, nA, nS;
  long long ldp;
  long long stack_elems, scratch_elems;
  long long off_pu, off_put, off_pr, off_ps, total_bytes;
};

static Layout make_layout(int n) {
  Layout L;
  L.n = n;
  L.nA = (n + 127) / 128 * 128;
  L.ldp = (n + 63) / 64 * 64;
  L.nS = ((L.nA / 128 + 1) / 2) * 128 + 128;
  if (L.nS > L.nA) L.nS = L.nA;
  L.stack_elems = 3LL * L.nA * L.ldp;
  L.scratch_elems = 3LL * L.nS * L.ldp;
  auto al = [](long long b) { return (b + 255) & ~255LL; };
  L.off_pu = 0;
  L.off_put = al(L.off_pu + L.stack_elems * 2);
  L.off_pr = al(L.off_put + L.stack_elems * 2);
  L.off_ps = al(L.off_pr + L.stack_elems * 2);
  L.total_bytes = al(L.off_ps + L.scratch_elems * 2);
  return L;
}

struct Ctx {
  float* A; long long lda;
  float* U; long long ldu;
  Layout L;
  __nv_bfloat16 *PU, *PUT, *PR, *PS;
  CUtensorMap mPU_a, mPU_b, mPUT_a, mPR_a, mPR_b, mPS_a, mPS_b;
  int* info;
  cudaStream_t s;
  int rc;
};

static void rec(Ctx& c, int lo_b, int hi_b) {
  if (c.rc) return;
  const int n = c.L.n;
  if (hi_b - lo_b == 1) {
    LeafParams P;
    P.A = c.A; P.lda = c.lda; P.U = c.U; P.ldu = c.ldu; P.PU = c.PU; P.PUT = c.PUT; P.ldp = c.L.ldp;
    P.plane_rows = c.L.nA; P.o = lo_b * LB; P.m = min(LB, n - P.o); P.info = c.info;
    k_chol_leaf<<<1, 256, LEAF_SMEM2, c.s>>>(P);
    if (cudaGetLastError() != cudaSuccess) c.rc = MSQ_ERR_CUDA;
    return;
  }
  const int mid_b = lo_b + (hi_b - lo_b) / 2;
  const int lo = lo_b * LB, mid = mid_b * LB, hi = min(hi_b * LB, n);
  const int n1 = mid - lo, n2 = hi - mid;
  rec(c, mid_b, hi_b);
  if (c.rc) return;
  // A12 -> planes (scratch)
  c.rc = split3_convert(c.A + (long long)lo * c.lda + mid, c.lda, n1, n2, c.PS, c.L.ldp, c.L.nS, 0, 0, c.s);
  if (c.rc) return;
  Split3Params P;
  memset(&P, 0, sizeof(P));
  P.ldp = c.L.ldp;
  // R12 = A12 U22^T  -> planes PR[lo:mid, mid:hi]
  P.M = n1; P.N = n2; P.K = n2;
  P.a_row0 = 0; P.a_col0 = 0; P.a_plane_rows = c.L.nS;
  P.b_row0 = mid; P.b_col0 = mid; P.b_plane_rows = c.L.nA;
  P.alpha = 1.0f; P.beta = 0.0f; P.C = nullptr;
  P.out_planes = c.PR; P.op_row0 = lo; P.op_col0 = mid; P.op_plane_rows = c.L.nA;
  P.k_lo_mode = 2;
  c.rc = split3_gemm(c.mPS_a, c.mPU_b, P, c.s);
  if (c.rc) return;
  // A11 -= R12 R12^T (upper tiles)
  memset(&P, 0, sizeof(P));
  P.ldp = c.L.ldp;
  P.M = n1; P.N = n1; P.K = n2;
  P.a_row0 = lo; P.a_col0 = mid; P.a_plane_rows = c.L.nA;
  P.b_row0 = lo; P.b_col0 = mid; P.b_plane_rows = c.L.nA;
  P.alpha = -1.0f; P.beta = 1.0f; P.C = c.A + (long long)lo * c.lda + lo; P.ldc = c.lda;
  P.upper_only = 1;
  c.rc = split3_gemm(c.mPR_a, c.mPR_b, P, c.s);
  if (c.rc) return;
  rec(c, lo_b, mid_b);
  if (c.rc) return;
  // T^T = U22^T R12^T  -> planes PS [n2, n1]
  memset(&P, 0, sizeof(P));
  P.ldp = c.L.ldp;
  P.M = n2; P.N = n1; P.K = n2;
  P.a_row0 = mid; P.a_col0 = mid; P.a_plane_rows = c.L.nA;
  P.b_row0 = lo; P.b_col0 = mid; P.b_plane_rows = c.L.nA;
  P.alpha = 1.0f; P.beta = 0.0f;
  P.out_planes = c.PS; P.op_row0 = 0; P.op_col0 = 0; P.op_plane_rows = c.L.nS;
  P.k_hi_mode = 1;
  c.rc = split3_gemm(c.mPUT_a, c.mPR_b, P, c.s);
  if (c.rc) return;
  // U12 = -U11 T  -> fp32 U[lo:mid, mid:hi], planes PU (same place) and PU^T (transposed place)
  memset(&P, 0, sizeof(P));
  P.ldp = c.L.ldp;
  P.M = n1; P.N = n2; P.K = n1;
  P.a_row0 = lo; P.a_col0 = lo; P.a_plane_rows = c.L.nA;
  P.b_row0 = 0; P.b_col0 = 0; P.b_plane_rows = c.L.nS;
  P.alpha = -1.0f; P.beta = 0.0f; P.C = c.U + (long long)lo * c.ldu + mid; P.ldc = c.ldu;
  P.out_planes = c.PU; P.op_row0 = lo; P.op_col0 = mid; P.op_plane_rows = c.L.nA;
  P.out_planes_t = c.PUT; P.opt_row0 = mid; P.opt_col0 = lo; P.opt_plane_rows = c.L.nA;
  P.k_lo_mode = 1;
  c.rc = split3_gemm(c.mPU_a, c.mPS_b, P, c.s);
}

// zero the strictly-lower part of U (the reference's cholesky(upper=True) returns exact zeros there)
__global__ void __launch_bounds__(256) k_zero_lower(float* __restrict__ U, long long ldu, int n) {
  const int r = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
  if (c < r && c < n) U[(long long)r * ldu + c] = 0.0f;
}

}  // namespace ci
}  // namespace msq

using namespace msq;

extern "C" {

long long msq_cholinv_workspace_bytes(int n) {
  if (n <= 0) return 0;
  return ci::make_layout(n).total_bytes;
}

// H (fp32 [n, n], ld = ldh, symmetric positive definite, upper triangle read, DESTROYED) -> U (fp32 [n, n], ld = ldu):
// the upper Cholesky factor of H^-1, i.e. what llm/gptq.py:101-103 leaves in `H`.  info (device int32, zeroed by the
// caller) receives 1 + the index of the first non-positive pivot (then U holds NaNs), like LAPACK's info.
int msq_cholinv_upper_f32(float* H, long long ldh, float* U, long long ldu, int n, void* workspace, long long workspace_bytes,
                          int* info, void* stream) {
  msq::DevGuard dev_guard_(H);
  if (!H || !U || !workspace || !info || n <= 0) return MSQ_ERR_BAD_ARG;
  if ((ldh % 4) || (ldu % 4) || (reinterpret_cast<uintptr_t>(H) & 15) || (reinterpret_cast<uintptr_t>(U) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 255))
    return MSQ_ERR_ALIGN;
  ci::Ctx c;
  c.L = ci::make_layout(n);
  if (workspace_bytes < c.L.total_bytes) return MSQ_ERR_BAD_ARG;
  c.A = H; c.lda = ldh; c.U = U; c.ldu = ldu; c.info = info; c.s = (cudaStream_t)stream; c.rc = 0;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  c.PU = reinterpret_cast<__nv_bfloat16*>(ws + c.L.off_pu);
  c.PUT = reinterpret_cast<__nv_bfloat16*>(ws + c.L.off_put);
  c.PR = reinterpret_cast<__nv_bfloat16*>(ws + c.L.off_pr);
  c.PS = reinterpret_cast<__nv_bfloat16*>(ws + c.L.off_ps);
  // blocks strictly below the diagonal are read as part of 256-wide tiles: they must hold zeros
  if (cudaMemsetAsync(c.PU, 0, c.L.stack_elems * 2, c.s) != cudaSuccess) return MSQ_ERR_CUDA;
  if (cudaMemsetAsync(c.PUT, 0, c.L.stack_elems * 2, c.s) != cudaSuccess) return MSQ_ERR_CUDA;
  if (cudaFuncSetAttribute(ci::k_chol_leaf, cudaFuncAttributeMaxDynamicSharedMemorySize, ci::LEAF_SMEM2) != cudaSuccess)
    return MSQ_ERR_CUDA;
  int rc = 0;
  rc |= split3_make_tmap(&c.mPU_a, c.PU, 3LL * c.L.nA, n, c.L.ldp, 128);
  rc |= split3_make_tmap(&c.mPU_b, c.PU, 3LL * c.L.nA, n, c.L.ldp, 256);
  rc |= split3_make_tmap(&c.mPUT_a, c.PUT, 3LL * c.L.nA, n, c.L.ldp, 128);
  rc |= split3_make_tmap(&c.mPR_a, c.PR, 3LL * c.L.nA, n, c.L.ldp, 128);
  rc |= split3_make_tmap(&c.mPR_b, c.PR, 3LL * c.L.nA, n, c.L.ldp, 256);
  rc |= split3_make_tmap(&c.mPS_a, c.PS, 3LL * c.L.nS, c.L.ldp, c.L.ldp, 128);
  rc |= split3_make_tmap(&c.mPS_b, c.PS, 3LL * c.L.nS, c.L.ldp, c.L.ldp, 256);
  if (rc) return MSQ_ERR_CUDA;
  ci::rec(c, 0, (n + ci::LB - 1) / ci::LB);
  if (c.rc) return c.rc;
  dim3 g((n + 255) / 256, n);
  ci::k_zero_lower<<<g, 256, 0, c.s>>>(U, ldu, n);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

// where the bf16 planes of U^T live inside the workspace after msq_cholinv_upper_f32: plane p of element (r, c) of
// U^T at planes[(p * plane_rows + r) * ld + c].  They are the B operand of msq_gptq_trailing_update_tc.
int msq_cholinv_ut_planes(void* workspace, int n, void** planes, long long* ld, int* plane_rows) {
  if (!workspace || n <= 0 || !planes || !ld || !plane_rows) return MSQ_ERR_BAD_ARG;
  const ci::Layout L = ci::make_layout(n);
  *planes = reinterpret_cast<uint8_t*>(workspace) + L.off_put;
  *ld = L.ldp;
  *plane_rows = L.nA;
  return MSQ_OK;
}

}  // extern "C"

extern "C" {

// fp32 [rows, cols] (ld) -> three bf16 planes whose sum is the input exactly: plane p of element (r, c) at
// planes[(p * plane_rows + r) * ldp + c]; columns up to the next multiple of 64 are zero-filled (ldp must cover them).
int msq_split3_planes_f32(const float* src, long long ld, int rows, int cols, void* planes, long long ldp, int plane_rows,
                          void* stream) {
  msq::DevGuard dev_guard_(src);
  if (!src || !planes || rows <= 0 || cols <= 0 || plane_rows < rows || ldp < (cols + 63) / 64 * 64) return MSQ_ERR_BAD_ARG;
  if ((ldp % 8) || (reinterpret_cast<uintptr_t>(planes) & 15)) return MSQ_ERR_ALIGN;
  return split3_convert(src, ld, rows, cols, reinterpret_cast<__nv_bfloat16*>(planes), ldp, plane_rows, 0, 0,
                        (cudaStream_t)stream);
}

// C[M,N] (fp32) = alpha * A[M,K] B[N,K]^T + beta * C with both operands given as bf16 plane stacks
// (msq_split3_planes_f32): the fp32-accurate tensor-core contraction behind the GPTQ trailing update
// (llm/gptq.py:165), the Cholesky recursion (llm/gptq.py:101-103) and mx.linear's dense fallback
// (number_system/mx/linear.py:91).  Error <= ~4 * 2^-24 * |A| |B|^T per element (fp32 accumulation included).
int msq_gemm_split3_f32(const void* a_planes, long long a_ld, int a_plane_rows, const void* b_planes, long long b_ld,
                        int b_plane_rows, float* C, long long ldc, int M, int N, int K, float alpha, float beta,
                        void* stream) {
  msq::DevGuard dev_guard_(C);
  if (!a_planes || !b_planes || !C || M <= 0 || N <= 0 || K <= 0 || a_plane_rows < M || b_plane_rows < N) return MSQ_ERR_BAD_ARG;
  if ((a_ld % 8) || (b_ld % 8) || (reinterpret_cast<uintptr_t>(a_planes) & 15) || (reinterpret_cast<uintptr_t>(b_planes) & 15))
    return MSQ_ERR_ALIGN;
  CUtensorMap mA, mB;
  // tensor width = K rounded up to the 64-element blocks the converter zero-filled
  const long long kc = ((long long)K + 63) / 64 * 64;
  if (split3_make_tmap(&mA, reinterpret_cast<const __nv_bfloat16*>(a_planes), 3LL * a_plane_rows, kc < a_ld ? kc : a_ld, a_ld, 128))
    return MSQ_ERR_CUDA;
  if (split3_make_tmap(&mB, reinterpret_cast<const __nv_bfloat16*>(b_planes), 3LL * b_plane_rows, kc < b_ld ? kc : b_ld, b_ld, 256))
    return MSQ_ERR_CUDA;
  Split3Params P;
  memset(&P, 0, sizeof(P));
  P.M = M; P.N = N; P.K = K;
  P.a_plane_rows = a_plane_rows; P.b_plane_rows = b_plane_rows;
  P.alpha = alpha; P.beta = beta; P.C = C; P.ldc = ldc;
  return split3_gemm(mA, mB, P, (cudaStream_t)stream);
}

}  // extern "C"
