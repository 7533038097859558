// msq_gptq.cu -- the column loop of GPTQ.fasterquant (llm/gptq.py:106-165) on B200.
//
// Rows are independent inside a 128-column GPTQ block except for the top-k "prune"
// (gptq.py:146-153), which is a value no-op whenever the column already holds at least k exact
// zeros (always the case for int2/int4/fp4 inliers).  k_gptq_block therefore runs the whole
// block without any cross-row synchronisation and records, per column, the number of zeros and
// the reference's k; msq_gptq_check_prune verifies zeros >= k afterwards.  If that ever fails the
// host re-runs the linear through the column-synchronous generic path (llm/gptq.py mirror).
//
// One warp owns 32 rows (lane = row): the [32 x 128] tile of W lives in shared memory (stride 129,
// conflict free), column i is quantized with the same block core as the stand-alone quantizer
// (BS lanes per MX block, xor-butterfly reductions), and the rank-1 update
// W1[:, j] -= err * Hinv1[i, j] (gptq.py:159) is applied to the remaining columns of the tile.
#include "msq_quant_core.cuh"
#include "msq_api_internal.h"
#include "msq_tc.cuh"
#include "msq_split3.cuh"

namespace msq {

constexpr int GB = 128;          // GPTQ block width (blocksize of gptq.py:61)
constexpr int TS = GB + 4;       // smem row stride: 16-byte aligned rows, float4 row updates conflict free

struct GptqParams {
  const float* W; long long ldw;     // [rows, >= i1+count] current weights
  float* Q; long long ldq;           // quantized weights out
  const float* Hinv; long long ldh;  // upper Cholesky factor of H^-1
  float* Err;                        // [rows, 128] (w - q) / d, zero padded (optional for k_gptq_block2)
  __nv_bfloat16* ErrPlanes;          // optional: the same values as three bf16 planes [3 * err_plane_rows, 128]
  int err_plane_rows;                //           (operand of the tensor-core trailing update)
  // optional packed-artefact outputs of k_gptq_block2<BS, true> (all or none; cols % 32 == 0):
  uint8_t* codes8; long long ldcodes;   // [rows, ldcodes] full element code (inlier or outlier format by the mask bit)
  uint32_t* bitmap;                  // 1 bit per element, row-major [rows, cols]
  int8_t* e_in; int8_t* e_out;       // [rows / BS, cols]
  int cols;
  double* loss;                      // sum (w-q)^2 / d^2 / 2
  int* zeros; int* kcnt;             // per column: #(q == 0), reference's k
  int* status;
  int rows, i1, count;
  QCfg c;
};

template <int BS>
__global__ void __launch_bounds__(32) k_gptq_block(const GptqParams P) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                 // [32][TS]  W1 -> Q1
  float* etile = sm + 32 * TS;      // [32][TS]  Err1
  float* hs = sm + 64 * TS;         // [128][GB] Hinv1 = Hinv[i1:i1+count, i1:i1+count] (zero padded)
  const int lane = threadIdx.x;
  const int r0 = blockIdx.x * 32;
  const int row = r0 + lane;
  const bool active = row < P.rows;
  const int count = P.count;

  for (int rr = 0; rr < 32; ++rr) {
    const bool rv = r0 + rr < P.rows;
    const float* src = P.W + (long long)(r0 + rr) * P.ldw + P.i1;
#pragma unroll
    for (int c = 0; c < GB / 32; ++c) {
      const int j = lane + 32 * c;
      tile[rr * TS + j] = (rv && j < count) ? __ldg(src + j) : 0.0f;
      etile[rr * TS + j] = 0.0f;
    }
  }
  for (int rr = 0; rr < GB; ++rr) {
    const float* src = P.Hinv + (long long)(P.i1 + rr) * P.ldh + P.i1;
#pragma unroll
    for (int c = 0; c < GB / 32; ++c) {
      const int j = lane + 32 * c;
      hs[rr * GB + j] = (rr < count && j < count) ? __ldg(src + j) : 0.0f;
    }
  }
  __syncwarp();

  double loss = 0.0;
  int st = 0;
  const bool counts_for_k = ((row / BS) % BS) == 0;      // quant.py:66  outlier_pos[::block_size]
  float* trow = tile + lane * TS;
#pragma unroll 1
  for (int i = 0; i < count; ++i) {
    const float* hrow = hs + i * GB;
    const float d = hrow[i];
    const float w = trow[i];
    float x[1] = {active ? w : 0.0f};
    uint32_t code[1];
    const BlockOut bo = core_outlier<1, BS, false, false>(x, code, P.c, nullptr, nullptr);
    const float q = x[0];
    st |= bo.st;
    const unsigned zmask = __ballot_sync(0xffffffffu, active && q == 0.0f);
    if (lane == 0) atomicAdd(P.zeros + P.i1 + i, __popc(zmask));
    if (active && counts_for_k && (lane % BS) == 0 && bo.cnt) atomicAdd(P.kcnt + P.i1 + i, bo.cnt);
    const float diff = __fsub_rn(w, q);
    const float err = __fdiv_rn(diff, d);
    if (active) loss += (double)__fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(d, d));
    trow[i] = q;
    etile[lane * TS + i] = active ? err : 0.0f;
    // W1[:, j] -= err * Hinv1[i, j] for j > i (gptq.py:159): scalar head up to the next multiple of 4,
    // then 128-bit row updates (Hinv1 row is a shared-memory broadcast)
    const int j4 = min((i + 4) & ~3, count);
    for (int j = i + 1; j < j4; ++j) trow[j] = __fsub_rn(trow[j], __fmul_rn(err, hrow[j]));
#pragma unroll 4
    for (int j = j4; j < count; j += 4) {
      float4 t = *reinterpret_cast<float4*>(trow + j);
      const float4 h = *reinterpret_cast<const float4*>(hrow + j);
      t.x = __fsub_rn(t.x, __fmul_rn(err, h.x));
      t.y = __fsub_rn(t.y, __fmul_rn(err, h.y));
      t.z = __fsub_rn(t.z, __fmul_rn(err, h.z));
      t.w = __fsub_rn(t.w, __fmul_rn(err, h.w));
      *reinterpret_cast<float4*>(trow + j) = t;
    }
  }
  __syncwarp();

  for (int rr = 0; rr < 32; ++rr) {
    if (r0 + rr >= P.rows) break;
    float* qd = P.Q + (long long)(r0 + rr) * P.ldq + P.i1;
    float* ed = P.Err + (long long)(r0 + rr) * GB;
#pragma unroll
    for (int c = 0; c < GB / 32; ++c) {
      const int j = lane + 32 * c;
      if (j < count) qd[j] = tile[rr * TS + j];
      ed[j] = etile[rr * TS + j];
    }
  }
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  if (lane == 0) atomicAdd(P.loss, loss * 0.5);
  st = __reduce_or_sync(0xffffffffu, st);
  if (lane == 0 && st) atomicOr(P.status, st);
}

// ---------------------------------------------------------------------------------------------------------------
// k_gptq_block2: the same column loop with the rank-1 updates taken off the critical path.
//
// The loop is a chain of 128 dependent column quantizations; what a column needs before it can be quantized is only
// ITS OWN value with every earlier error applied, in order.  Two warps share 32 rows (lane = row):
//   * warp Q owns the current micro-block of 8 columns in registers: it applies the 8 errors of the previous
//     micro-block to them, then quantizes the 8 columns one after the other, updating the rest of the micro-block in
//     registers after each (gptq.py:159 restricted to the micro-block);
//   * warp U prepares the NEXT micro-block while Q works: it applies every error older than the previous
//     micro-block (left-looking, accumulating in registers, Hinv rows as shared-memory broadcasts).
// Per weight the sequence of fp32 operations is exactly the reference's -- w_j = ((w_j - e_0 h_0j) - e_1 h_1j) - ...
// in ascending i, each product and difference rounded separately -- so results are bit-identical to k_gptq_block
// and to the reference's column loop; only who executes them, and when, changes.  One named barrier per micro-block.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MBK = 8;

__device__ __forceinline__ void bar_pair() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

template <int BS, bool EMIT>
__global__ void __launch_bounds__(64) k_gptq_block2(const GptqParams P) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                 // [32][TS]  W1 -> Q1
  float* etile = sm + 32 * TS;      // [32][TS]  Err1
  float* hs = sm + 64 * TS;         // [128][GB] Hinv1
  // EMIT: element codes, outlier bits and block exponents of the quantized columns (the packed artefact)
  uint8_t* ctile = reinterpret_cast<uint8_t*>(hs + GB * GB);          // [32][GB] codes
  uint32_t* mtile = reinterpret_cast<uint32_t*>(ctile + 32 * GB);      // [32][4]  outlier bits
  int8_t* eitile = reinterpret_cast<int8_t*>(mtile + 32 * 4);          // [32 / BS][GB]
  int8_t* eotile = eitile + (32 / BS) * GB;                            // [32 / BS][GB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * 32;
  const int row = r0 + lane;
  const bool active = row < P.rows;
  const int count = P.count;

  for (int rr = warp; rr < 32; rr += 2) {
    const bool rv = r0 + rr < P.rows;
    const float* src = P.W + (long long)(r0 + rr) * P.ldw + P.i1;
#pragma unroll
    for (int c = 0; c < GB / 32; ++c) {
      const int j = lane + 32 * c;
      tile[rr * TS + j] = (rv && j < count) ? __ldg(src + j) : 0.0f;
      etile[rr * TS + j] = 0.0f;
    }
    if (EMIT && lane < 4) mtile[rr * 4 + lane] = 0;
  }
  for (int rr = warp; rr < GB; rr += 2) {
    const float* src = P.Hinv + (long long)(P.i1 + rr) * P.ldh + P.i1;
#pragma unroll
    for (int c = 0; c < GB / 32; ++c) {
      const int j = lane + 32 * c;
      hs[rr * GB + j] = (rr < count && j < count) ? __ldg(src + j) : 0.0f;
    }
  }
  __syncthreads();

  const int nm = (count + MBK - 1) / MBK;
  float* trow = tile + lane * TS;
  float* erow = etile + lane * TS;
  if (warp == 0) {
    double loss = 0.0;
    int st = 0;
    const bool counts_for_k = ((row / BS) % BS) == 0;      // quant.py:66  outlier_pos[::block_size]
    float ep[MBK];
#pragma unroll
    for (int c = 0; c < MBK; ++c) ep[c] = 0.0f;
#pragma unroll 1
    for (int m = 0; m < nm; ++m) {
      const int j0 = m * MBK;
      if (m > 0) bar_pair();
      float w[MBK];
      {
        const float4 a = *reinterpret_cast<const float4*>(trow + j0);
        const float4 b = *reinterpret_cast<const float4*>(trow + j0 + 4);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
      }
      if (m > 0) {       // errors of the previous micro-block, in order
#pragma unroll
        for (int i = 0; i < MBK; ++i) {
          const float* hrow = hs + (j0 - MBK + i) * GB + j0;
          const float4 a = *reinterpret_cast<const float4*>(hrow);
          const float4 b = *reinterpret_cast<const float4*>(hrow + 4);
          const float h[MBK] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int c = 0; c < MBK; ++c) w[c] = __fsub_rn(w[c], __fmul_rn(ep[i], h[c]));
        }
      }
#pragma unroll
      for (int c = 0; c < MBK; ++c) {
        const int i = j0 + c;
        if (i < count) {                        // warp-uniform
          const float* hrow = hs + i * GB + j0;
          const float4 ha = *reinterpret_cast<const float4*>(hrow);
          const float4 hb = *reinterpret_cast<const float4*>(hrow + 4);
          const float h[MBK] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
          const float d = h[c];
          const float wv = w[c];
          float x[1] = {active ? wv : 0.0f};
          uint32_t code[1];
          const BlockOut bo = core_outlier<1, BS, EMIT, false>(x, code, P.c, nullptr, nullptr);
          const float q = x[0];
          st |= bo.st;
          const unsigned zmask = __ballot_sync(0xffffffffu, active && q == 0.0f);
          if (lane == 0) atomicAdd(P.zeros + P.i1 + i, __popc(zmask));
          if (active && counts_for_k && (lane % BS) == 0 && bo.cnt) atomicAdd(P.kcnt + P.i1 + i, bo.cnt);
          const float diff = __fsub_rn(wv, q);
          const float err = __fdiv_rn(diff, d);
          if (active) loss += (double)__fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(d, d));
          ep[c] = active ? err : 0.0f;
          trow[i] = q;
          if (EMIT) {
            ctile[lane * GB + i] = (uint8_t)code[0];
            if (bo.mask & 1u) mtile[lane * 4 + (i >> 5)] |= 1u << (i & 31);
            if ((lane % BS) == 0) {
              eitile[(lane / BS) * GB + i] = (int8_t)bo.e_in;
              eotile[(lane / BS) * GB + i] = (int8_t)bo.e_out;
            }
          }
#pragma unroll
          for (int c2 = c + 1; c2 < MBK; ++c2) w[c2] = __fsub_rn(w[c2], __fmul_rn(err, h[c2]));
        } else {
          ep[c] = 0.0f;
        }
      }
      *reinterpret_cast<float4*>(erow + j0) = make_float4(ep[0], ep[1], ep[2], ep[3]);
      *reinterpret_cast<float4*>(erow + j0 + 4) = make_float4(ep[4], ep[5], ep[6], ep[7]);
    }
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) atomicAdd(P.loss, loss * 0.5);
    st = __reduce_or_sync(0xffffffffu, st);
    if (lane == 0 && st) atomicOr(P.status, st);
  } else {
#pragma unroll 1
    for (int r = 1; r < nm; ++r) {
      bar_pair();                         // errors of micro-block r-1 are now visible
      if (r + 1 >= nm) break;
      const int j0 = (r + 1) * MBK;       // prepare micro-block r+1 with errors 0 .. 8r-1
      float w[MBK];
      {
        const float4 a = *reinterpret_cast<const float4*>(trow + j0);
        const float4 b = *reinterpret_cast<const float4*>(trow + j0 + 4);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
      }
#pragma unroll 1
      for (int i4 = 0; i4 < r * MBK; i4 += 4) {
        const float4 e4 = *reinterpret_cast<const float4*>(erow + i4);
        const float e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const float* hrow = hs + (i4 + ii) * GB + j0;
          const float4 a = *reinterpret_cast<const float4*>(hrow);
          const float4 b = *reinterpret_cast<const float4*>(hrow + 4);
          const float h[MBK] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int c = 0; c < MBK; ++c) w[c] = __fsub_rn(w[c], __fmul_rn(e[ii], h[c]));
        }
      }
      *reinterpret_cast<float4*>(trow + j0) = make_float4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<float4*>(trow + j0 + 4) = make_float4(w[4], w[5], w[6], w[7]);
    }
  }
  __syncthreads();

  const long long eps = (long long)P.err_plane_rows * GB;
  for (int rr = warp; rr < 32; rr += 2) {
    if (r0 + rr >= P.rows) break;
    float* qd = P.Q + (long long)(r0 + rr) * P.ldq + P.i1;
#pragma unroll
    for (int c = 0; c < GB / 32; ++c) {
      const int j = lane + 32 * c;
      if (j < count) qd[j] = tile[rr * TS + j];
      const float e = etile[rr * TS + j];
      if (P.Err) P.Err[(long long)(r0 + rr) * GB + j] = e;
      if (P.ErrPlanes) {
        const __nv_bfloat16 p0 = __float2bfloat16_rn(e);
        const float r1 = e - __bfloat162float(p0);
        const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
        __nv_bfloat16* ep = P.ErrPlanes + (long long)(r0 + rr) * GB + j;
        ep[0] = p0; ep[eps] = p1; ep[2 * eps] = p2;
      }
      if (EMIT && j < count) P.codes8[(long long)(r0 + rr) * P.ldcodes + P.i1 + j] = ctile[rr * GB + j];
    }
    if (EMIT && lane < 4 && P.i1 + lane * 32 < P.cols)
      P.bitmap[(((long long)(r0 + rr) * P.cols + P.i1) >> 5) + lane] = mtile[rr * 4 + lane];
    if (EMIT && (rr % BS) == 0) {
      const long long eb = (long long)((r0 + rr) / BS) * P.cols + P.i1;
#pragma unroll
      for (int c = 0; c < GB / 32; ++c) {
        const int j = lane + 32 * c;
        if (j < count) { P.e_in[eb + j] = eitile[(rr / BS) * GB + j]; P.e_out[eb + j] = eotile[(rr / BS) * GB + j]; }
      }
    }
  }
}

// W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:]   (gptq.py:165), strict fp32 FMA accumulation.
// 64 x 64 output tile per CTA, 4 x 4 outputs per thread, whole K = 128 staged in shared memory.
__global__ void __launch_bounds__(256) k_gptq_trailing(float* __restrict__ W, long long ldw, const float* __restrict__ Err,
                                                       const float* __restrict__ Hinv, long long ldh, int rows, int i1,
                                                       int count, int i2, int cols) {
  __shared__ float sA[64][64 + 4];     // Err tile  [m][k-chunk]
  __shared__ float sB[64][64 + 4];     // Hinv tile [k-chunk][n]
  const int m0 = blockIdx.y * 64, n0 = i2 + blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < count; k0 += 64) {
    __syncthreads();
    for (int idx = tid; idx < 64 * 64; idx += 256) {
      const int m = idx / 64, k = idx - m * 64;
      sA[m][k] = (m0 + m < rows && k0 + k < count) ? __ldg(Err + (long long)(m0 + m) * GB + k0 + k) : 0.0f;
      const int kk = idx / 64, n = idx - kk * 64;
      sB[kk][n] = (k0 + kk < count && n0 + n < cols) ? __ldg(Hinv + (long long)(i1 + k0 + kk) * ldh + n0 + n) : 0.0f;
    }
    __syncthreads();
    const int kend = min(64, count - k0);
    for (int k = 0; k < kend; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = sA[tm + u][k]; b[u] = sB[k][tn + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int m = m0 + tm + u;
    if (m >= rows) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int n = n0 + tn + v;
      if (n < cols) W[(long long)m * ldw + n] -= acc[u][v];
    }
  }
}

__global__ void k_gptq_check(const int* __restrict__ zeros, const int* __restrict__ kcnt, int cols, int* violation) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cols) {
    const int k = (int)(short)kcnt[i];               // .to(torch.int16), gptq.py:147
    if (zeros[i] < k) atomicAdd(violation, 1);
  }
}

// rank-1 update used by the generic (column-synchronous) path: W1[:, i:] -= err (x) h
__global__ void __launch_bounds__(256) k_rank1_sub(float* __restrict__ W, long long ldw, const float* __restrict__ err,
                                                   const float* __restrict__ h, int rows, int n) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)rows * n) return;
  const int r = (int)(t / n), j = (int)(t - (long long)r * n);
  float* p = W + (long long)r * ldw + j;
  *p = __fsub_rn(*p, __fmul_rn(__ldg(err + r), __ldg(h + j)));
}

}  // namespace msq

using namespace msq;

extern "C" {

// One 128-column GPTQ block [i1, i1+count) for all rows.  zeros/kcnt are [cols] int32 (caller zeroes
// them once per linear), loss a device double, Err a [rows,128] fp32 scratch, status OR of MSQ_ST_*.
int msq_gptq_block(const float* W, long long ldw, float* Q, long long ldq, const float* Hinv, long long ldh, int rows, int i1,
                   int count, float* Err, double* loss, int* zeros, int* kcnt, int* status, int block_size,
                   int in_scale_bits, int out_scale_bits, int in_ebits, int in_mbits, int in_emax, float in_max_norm,
                   int out_ebits, int out_mbits, int out_emax, float out_max_norm, float std_dev, int flush_fp32_subnorms,
                   int rmode, void* stream) {
  msq::DevGuard dev_guard_(W);
  if (!W || !Q || !Hinv || !Err || !loss || !zeros || !kcnt || !status || rows <= 0 || count <= 0 || count > GB)
    return MSQ_ERR_BAD_ARG;
  if (in_ebits > 5 || out_ebits > 5) return MSQ_ERR_UNSUPPORTED;
  GptqParams P;
  P.W = W; P.ldw = ldw; P.Q = Q; P.ldq = ldq; P.Hinv = Hinv; P.ldh = ldh; P.Err = Err; P.loss = loss; P.zeros = zeros;
  P.ErrPlanes = nullptr; P.err_plane_rows = 0;
  P.codes8 = nullptr; P.ldcodes = 0; P.bitmap = nullptr; P.e_in = nullptr; P.e_out = nullptr; P.cols = 0;
  P.kcnt = kcnt; P.status = status; P.rows = rows; P.i1 = i1; P.count = count;
  P.c.fin = make_fmt(in_ebits, in_mbits, in_emax, in_max_norm);
  P.c.fout = make_fmt(out_ebits, out_mbits, out_emax, out_max_norm);
  P.c.sb_in = in_scale_bits; P.c.sb_out = out_scale_bits; P.c.std_dev = std_dev; P.c.flush = flush_fp32_subnorms;
  P.c.variant = 0; P.c.mode = 0; P.c.rmode = rmode;
  finish_cfg(P.c);
  const unsigned grid = (unsigned)((rows + 31) / 32);
  const size_t smem = (2 * 32 * TS + GB * GB) * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  void (*k)(const GptqParams) = nullptr;
  if (block_size == 16) k = k_gptq_block<16>;
  else if (block_size == 32) k = k_gptq_block<32>;
  else if (block_size == 8) k = k_gptq_block<8>;
  else return MSQ_ERR_UNSUPPORTED;
  // the opt-in is per device: set it on every launch instead of once per process
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MSQ_ERR_CUDA;
  k<<<grid, 32, smem, s>>>(P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_gptq_trailing_update(float* W, long long ldw, const float* Err, const float* Hinv, long long ldh, int rows, int i1,
                             int count, int i2, int cols, void* stream) {
  msq::DevGuard dev_guard_(W);
  if (!W || !Err || !Hinv || rows <= 0 || count <= 0 || count > GB) return MSQ_ERR_BAD_ARG;
  if (i2 >= cols) return MSQ_OK;
  dim3 grid((unsigned)((cols - i2 + 63) / 64), (unsigned)((rows + 63) / 64));
  k_gptq_trailing<<<grid, 256, 0, (cudaStream_t)stream>>>(W, ldw, Err, Hinv, ldh, rows, i1, count, i2, cols);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_gptq_check_prune(const int* zeros, const int* kcnt, int cols, int* violation, void* stream) {
  msq::DevGuard dev_guard_(zeros);
  if (!zeros || !kcnt || !violation || cols <= 0) return MSQ_ERR_BAD_ARG;
  k_gptq_check<<<(cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(zeros, kcnt, cols, violation);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

int msq_rank1_sub_f32(float* W, long long ldw, const float* err, const float* h, int rows, int n, void* stream) {
  msq::DevGuard dev_guard_(W);
  if (!W || !err || !h || rows <= 0 || n < 0) return MSQ_ERR_BAD_ARG;
  if (n == 0) return MSQ_OK;
  const long long tot = (long long)rows * n;
  k_rank1_sub<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, ldw, err, h, rows, n);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}


long long msq_gptq_solve_workspace_bytes(int rows) {
  if (rows <= 0) return 0;
  const long long rp = ((long long)rows + 127) / 128 * 128;
  return rp * GB * 4 /* Err fp32 */ + 3 * rp * GB * 2 /* Err planes */;
}

// The whole fused column loop of GPTQ.fasterquant (llm/gptq.py:106-165) for one linear, queued on `stream` without a
// host round trip: per 128-column block k_gptq_block2, then the trailing update W[:, i2:] -= Err1 Hinv[i1:i2, i2:] on
// tcgen05 (split-plane GEMM, B operand = the U^T planes msq_cholinv_upper_f32 left in its workspace) or, when
// ut_planes is NULL or the weight is not 16-byte aligned, on the fp32 SIMT kernel; finally the prune check.
// W is the working copy (modified), Q the result; loss / zeros[cols] / kcnt[cols] / status / violation as in
// msq_gptq_block / msq_gptq_check_prune (the caller zeroes them); workspace from msq_gptq_solve_workspace_bytes.
int msq_gptq_solve_f32(float* W, long long ldw, float* Q, long long ldq, const float* Hinv, long long ldh,
                       const void* ut_planes, long long ut_ld, int ut_plane_rows, int rows, int cols, void* workspace,
                       long long workspace_bytes, double* loss, int* zeros, int* kcnt, int* status, int* violation,
                       int block_size, int in_scale_bits, int out_scale_bits, int in_ebits, int in_mbits, int in_emax,
                       float in_max_norm, int out_ebits, int out_mbits, int out_emax, float out_max_norm, float std_dev,
                       int flush_fp32_subnorms, int rmode, unsigned char* codes8, unsigned int* bitmap,
                       signed char* e_in, signed char* e_out, void* stream) {
  msq::DevGuard dev_guard_(W);
  if (!W || !Q || !Hinv || !workspace || !loss || !zeros || !kcnt || !status || !violation || rows <= 0 || cols <= 0)
    return MSQ_ERR_BAD_ARG;
  const bool emit = codes8 != nullptr;
  if (emit && (!bitmap || !e_in || !e_out || (cols % 32) || (rows % block_size))) return MSQ_ERR_BAD_ARG;
  if (workspace_bytes < msq_gptq_solve_workspace_bytes(rows)) return MSQ_ERR_BAD_ARG;
  if (in_ebits > 5 || out_ebits > 5) return MSQ_ERR_UNSUPPORTED;
  void (*k)(const GptqParams) = nullptr;
  if (block_size == 16) k = emit ? k_gptq_block2<16, true> : k_gptq_block2<16, false>;
  else if (block_size == 32) k = emit ? k_gptq_block2<32, true> : k_gptq_block2<32, false>;
  else if (block_size == 8) k = emit ? k_gptq_block2<8, true> : k_gptq_block2<8, false>;
  else return MSQ_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int rp = (rows + 127) / 128 * 128;
  float* Err = reinterpret_cast<float*>(workspace);
  __nv_bfloat16* EP = reinterpret_cast<__nv_bfloat16*>(Err + (long long)rp * GB);
  const bool tc = ut_planes && !(reinterpret_cast<uintptr_t>(W) & 15) && (ldw % 4 == 0) &&
                  !(reinterpret_cast<uintptr_t>(ut_planes) & 15) && (ut_ld % 8 == 0);
  GptqParams P;
  P.W = W; P.ldw = ldw; P.Q = Q; P.ldq = ldq; P.Hinv = Hinv; P.ldh = ldh; P.loss = loss; P.zeros = zeros;
  P.kcnt = kcnt; P.status = status; P.rows = rows;
  P.Err = tc ? nullptr : Err; P.ErrPlanes = tc ? EP : nullptr; P.err_plane_rows = rp;
  P.codes8 = codes8; P.ldcodes = cols; P.bitmap = bitmap; P.e_in = e_in; P.e_out = e_out; P.cols = cols;
  P.c.fin = make_fmt(in_ebits, in_mbits, in_emax, in_max_norm);
  P.c.fout = make_fmt(out_ebits, out_mbits, out_emax, out_max_norm);
  P.c.sb_in = in_scale_bits; P.c.sb_out = out_scale_bits; P.c.std_dev = std_dev; P.c.flush = flush_fp32_subnorms;
  P.c.variant = 0; P.c.mode = 0; P.c.rmode = rmode;
  finish_cfg(P.c);
  const size_t smem = (2 * 32 * TS + GB * GB) * sizeof(float) + (emit ? 32 * GB + 32 * 16 + 2 * (32 / block_size) * GB : 0);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MSQ_ERR_CUDA;
  CUtensorMap mE, mUT;
  if (tc) {
    if (split3_make_tmap(&mE, EP, 3LL * rp, GB, GB, 128)) return MSQ_ERR_CUDA;
    if (split3_make_tmap(&mUT, reinterpret_cast<const __nv_bfloat16*>(ut_planes), 3LL * ut_plane_rows, cols, ut_ld, 256))
      return MSQ_ERR_CUDA;
  }
  const unsigned grid = (unsigned)((rows + 31) / 32);
  for (int i1 = 0; i1 < cols; i1 += GB) {
    const int count = cols - i1 < GB ? cols - i1 : GB;
    const int i2 = i1 + count;
    P.i1 = i1; P.count = count;
    k<<<grid, 64, smem, s>>>(P);
    if (cudaGetLastError() != cudaSuccess) return MSQ_ERR_CUDA;
    if (i2 >= cols) break;
    if (tc) {
      Split3Params G;
      memset(&G, 0, sizeof(G));
      G.M = rows; G.N = cols - i2; G.K = count;
      G.a_row0 = 0; G.a_col0 = 0; G.a_plane_rows = rp;
      G.b_row0 = i2; G.b_col0 = i1; G.b_plane_rows = ut_plane_rows;
      G.alpha = -1.0f; G.beta = 1.0f; G.C = W + i2; G.ldc = ldw;
      const int rc = split3_gemm(mE, mUT, G, s);
      if (rc) return rc;
    } else {
      dim3 g2((unsigned)((cols - i2 + 63) / 64), (unsigned)((rows + 63) / 64));
      k_gptq_trailing<<<g2, 256, 0, s>>>(W, ldw, Err, Hinv, ldh, rows, i1, count, i2, cols);
      if (cudaGetLastError() != cudaSuccess) return MSQ_ERR_CUDA;
    }
  }
  k_gptq_check<<<(cols + 255) / 256, 256, 0, s>>>(zeros, kcnt, cols, violation);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // extern "C"
