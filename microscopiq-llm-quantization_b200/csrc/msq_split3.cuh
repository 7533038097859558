// msq_split3.cuh -- interface of the split-plane tensor-core GEMM (msq_gemm_split3.cu) used by the GPTQ solver
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace msq {

// A plane stack is a row-major bf16 matrix [3 * plane_rows, ldp]: plane p of logical element (r, c) lives at row
// p * plane_rows + r.  Operands are sub-views (row0, col0, extent M x K / N x K) of a stack; the tensor map covers
// the whole stack with its true column count, so only the stack's right edge is zero-filled by TMA: sub-view K
// extents must end on a multiple of 64 or on that edge.  Rows read beyond M / N only feed masked outputs.
struct Split3Params {
  int M, N, K;
  int a_row0, a_col0, a_plane_rows;
  int b_row0, b_col0, b_plane_rows;
  float alpha, beta;
  float* C; long long ldc;                 // fp32 result (optional): C = alpha * A B^T + beta * C
  __nv_bfloat16* out_planes;               // result as planes at (op_row0 + i, op_col0 + j) of a stack (optional)
  int op_row0, op_col0, op_plane_rows;
  __nv_bfloat16* out_planes_t;             // result transposed: element (i, j) at (opt_row0 + j, opt_col0 + i) (optional)
  int opt_row0, opt_col0, opt_plane_rows;
  long long ldp;                           // leading dimension of the OUTPUT stacks
  int upper_only;                          // symmetric result: skip tiles strictly below the diagonal
  int k_lo_mode;                           // 0: from 0; 1: from the tile's first row (A upper triangular); 2: from its first column (B upper triangular)
  int k_hi_mode;                           // 0: to K;   1: to the tile's last row + 1 (A lower triangular)
};

int split3_make_tmap(CUtensorMap* out, const __nv_bfloat16* stack, long long stack_rows, long long cols, long long ldp,
                     int box_rows);
int split3_gemm(const CUtensorMap& tmA /*box 128*/, const CUtensorMap& tmB /*box 256*/, const Split3Params& P,
                cudaStream_t stream);
int split3_convert(const float* src, long long ld, int rows, int cols, __nv_bfloat16* planes, long long ldp, int plane_rows,
                   int row0, int col0, cudaStream_t stream);

}  // namespace msq
