/*
 * msq_b200.h -- C ABI of libmsq_b200.so, the B200 (sm_100a) implementation of the
 * MicroScopiQ quantization hot path.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the device-pointer entry points own nothing: the caller allocates every buffer, nothing is
 *     allocated behind its back and they keep no state between calls; they run on the device that
 *     owns their first buffer (the calling thread's current device is restored on return) and on
 *     the caller's `stream` (a cudaStream_t of that device; NULL = legacy default stream);
 *   - the ONE exception is the host-buffer pipeline (msq_host_*, at the end of this header): it
 *     keeps, per device, three streams, staging buffers sized to the largest chunk seen and a status
 *     word, created on first use, guarded by a mutex and freed by msq_host_release();
 *   - every function returns MSQ_OK (0) or a negative MSQ_ERR_*; nothing calls exit()
 *     (the reference's gpuErrchk does: number_system/mx/cpp/common.cuh:211-219);
 *   - there is no CPU fallback behind any of these symbols.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef MSQ_B200_H
#define MSQ_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MSQ_OK 0
#define MSQ_ERR_BAD_ARG (-1)
#define MSQ_ERR_UNSUPPORTED (-2)   /* geometry/format combination has no packed fast path */
#define MSQ_ERR_CUDA (-3)
#define MSQ_ERR_ALIGN (-4)

/* status word bits (device int32, OR-accumulated by the quantizers) */
#define MSQ_ST_NONFINITE_INPUT 1   /* reference: assert, utils/quant.py:479,103 */
#define MSQ_ST_EXP_IN_OVERFLOW 2   /* reference: NaN scale, utils/quant.py:88 */
#define MSQ_ST_EXP_OUT_OVERFLOW 4  /* reference: assert, utils/quant.py:119-122 */

/* rounding modes == RoundingMode, number_system/mx/formats.py:15-18 / cpp/common.cuh:130-134 */
#define MSQ_ROUND_NEAREST 0        /* half away from zero */
#define MSQ_ROUND_FLOOR 1
#define MSQ_ROUND_EVEN 2

/* ---- library / layout queries --------------------------------------------------------- */
int msq_version(void);
int msq_seg_blocks(void);                                   /* blocks per packed segment (1024) */
int msq_packed_code_width(int ebits, int mbits);            /* 2, 4 or 8 bits per stored code */
int msq_packed_hi_width(int in_ebits, int in_mbits, int out_ebits, int out_mbits); /* 0,2,4,8 */
long long msq_packed_hi_seg_stride(int block_size, int hi_w);  /* bytes between segments' hi regions */

/* ---- element-wise rounding ------------------------------------------------------------
 * replaces quantize_elemwise_func_cuda   number_system/mx/cpp/funcs.cpp:178-195
 *          (kernel cpp/elemwise.cuh:18-38) and the torch path of
 *          _quantize_elemwise_core       number_system/mx/elemwise_ops.py:84-174.
 * dtype: 0 = fp32, 1 = fp16, 2 = bf16 (in and out).  bits includes sign + implicit one.
 * native = 0: float formulation of elemwise_ops.py (what custom_cuda=False computes);
 * native = 1: integer-mantissa formulation of cpp/quantize.cuh:88-149. */
int msq_quantize_elemwise(const void* in, void* out, long long n, int dtype,
                          int bits, int exp_bits, float max_norm, int rmode,
                          int saturate_normals, int allow_denorm, int native, void* stream);

/* ---- plain MX block quantization ------------------------------------------------------
 * replaces quantize_mx_by_tile_func_cuda  cpp/funcs.cpp:157-176 (mx.cu:77-172),
 *          quantize_mx_func_cuda          cpp/funcs.cpp:135-155 (mx.cu:13-72; pass ext_max),
 *          and the torch path of _quantize_mx   number_system/mx/mx_ops.py:332-457.
 * The tensor is [pre, L, post] row-major, blocks of block_size along L (zero padded).
 * mode 0 = Python-path arithmetic (mx_ops.py:426-451, divisor 2^e + 1e-6 of :444),
 * mode 1 = native arithmetic (cpp/mx.cuh:37-52, cpp/shared_exp.cuh:14-53).
 * Optional outputs (NULL to skip): deq [same shape], codes (packed, code-width bits per
 * element, linear order), e_shared int8 [pre*nblk*post] (-128 = NaN scale). */
int msq_quantize_mx_f32(const float* in, float* deq, long long pre, long long L, long long post,
                        int block_size, int scale_bits, int ebits, int mbits, int emax,
                        float max_norm, int flush_fp32_subnorms, int rmode, int mode,
                        const float* ext_max, unsigned char* codes, signed char* e_shared,
                        int* status, int* host_path, void* stream);

/* The activation side of the block-scaled linears (number_system/mx/linear.py:66-73 quantizes the activations of every
 * call): [rows, K] fp32 (x_dtype 0), bf16 (1) or fp16 (2) activations, blocks of 32 along K, straight to packed FP4 (E2M1,
 * elem_is_fp4 != 0) or FP8 (E4M3) codes + int8 shared exponents [rows, K/32].  Arithmetic and planes of
 * msq_quantize_mx_f32(mode = 1, block 32, round to nearest) bit for bit; one launch, the input read in its own dtype.
 * ldx in elements, ldc_bytes / lde in bytes; K % 32 == 0, 16-byte aligned rows. */
int msq_quantize_mx_act(const void* x, int x_dtype, long long ldx, long long rows, int K, int elem_is_fp4,
                        int scale_bits, int flush_fp32_subnorms, unsigned char* codes, long long ldc_bytes,
                        signed char* e_shared, long long lde, int* status, void* stream);

/* ---- MX + outlier quantization (the MicroScopiQ quantizer) ----------------------------
 * replaces quantize_mx_outlier_v1        utils/quant.py:147-266   (variant 0)
 *          quantize_mx_outlier_hessian   utils/quant.py:23-146    (variant 0 + blk_cnt)
 *          _quantize_mx_outlier_v1       number_system/mx/mx_ops.py:210-330 (variant 1,
 *                                        lo_ext/hi_ext from msq_outlier_bounds_across_blocks_f32)
 * Outputs, all optional:
 *   deq      fp32 dequantized tensor (what the reference returns)
 *   codes    inlier-width slot per element (outlier slots hold the low bits of the outlier code)
 *   e_in / e_out   int8 per block, block-linear order (p*nblk + b)*post + j
 *   bitmap   1 bit per element, linear element order, 32-bit little-endian words
 *   hi       segment-compacted high bits of outlier codes (see DESIGN.md "packed layout")
 *   seg_cnt  outliers per segment;  blk_cnt  outliers per block (uint8);  mask  1 byte/element
 *   status   OR of MSQ_ST_*;  host_path  which kernel ran (1 strided, 2 innermost, 3 generic, 4 strided through the
 *            persistent TMA-staged kernel: 2-D input, values only; 5 = its quantize+pack sibling) */
int msq_quantize_mx_outlier_f32(
    const float* in, float* deq, long long pre, long long L, long long post, int block_size,
    int in_scale_bits, int out_scale_bits,
    int in_ebits, int in_mbits, int in_emax, float in_max_norm,
    int out_ebits, int out_mbits, int out_emax, float out_max_norm,
    float std_dev, int flush_fp32_subnorms, int rmode, int variant,
    const float* lo_ext, const float* hi_ext,
    unsigned char* codes, signed char* e_in, signed char* e_out, unsigned int* bitmap,
    unsigned char* hi, int* seg_cnt, unsigned char* blk_cnt, unsigned char* mask,
    int* status, int* host_path, void* stream);

/* The same quantizer (variant 0, blocks along the rows: axes=[0]) for up to 16 weight matrices [rows[i], cols[i]]
 * in ONE launch: what llm/llama.py:238-253 does linear by linear for a decoder layer.  in / out are HOST arrays of
 * n DEVICE pointers; bit-identical to n calls of msq_quantize_mx_outlier_f32(pre = 1, L = rows[i], post = cols[i]). */
int msq_quantize_mx_outlier_multi_f32(int n, const float* const* in, float* const* out, const long long* rows,
                                      const long long* cols, int block_size, int in_scale_bits, int out_scale_bits,
                                      int in_ebits, int in_mbits, int in_emax, float in_max_norm,
                                      int out_ebits, int out_mbits, int out_emax, float out_max_norm,
                                      float std_dev, int flush_fp32_subnorms, int rmode, int* status, void* stream);

/* detector statistics of number_system/mx/mx_ops.py:49-83 as called at :248 (mean / unbiased
 * std of signed A over the block-count axis).  lo, hi: fp32 [pre, block_size, post]. */
int msq_outlier_bounds_across_blocks_f32(const float* in, float* lo, float* hi, long long pre,
                                         long long L, long long post, int block_size,
                                         float std_dev, void* stream);

/* ---- innermost-dimension reductions ------------------------------------------------------
 * replaces reduce_sum_inner_dim / reduce_max_inner_dim   number_system/mx/cpp/funcs.cpp:197-211
 *          (kernels cpp/reduce.cuh:150-210).  in [rows, inner] fp32 contiguous, out [rows];
 * op 0 = sum, 1 = max (start value -65504, cpp/reduce.cuh:87); 32 <= inner <= 1024, inner % 32 == 0
 * (the reference's AT_ASSERTM, reduce.cu:34-35), else MSQ_ERR_UNSUPPORTED.  The fp32 sum uses the
 * reference's two-level shuffle-tree association order. */
int msq_reduce_inner_dim_f32(const float* in, float* out, long long rows, int inner, int op, void* stream);

/* ---- packed planes -> dense -------------------------------------------------------------
 * Inverse of the packed outputs of msq_quantize_mx_outlier_f32 / msq_quantize_mx_f32 (plain != 0).
 * The reference has no packed artefact (llm/opt.py:254-294 is dead code); the invariant checked is
 * unpack(pack(W)) == quantize_mx_outlier_v1(W) bit for bit (utils/quant.py:147-266).
 * out_f32 and/or out_bf16 (device, same shape as the original tensor). */
int msq_unpack_mx(const unsigned char* codes, const signed char* e_in, const signed char* e_out,
                  const unsigned int* bitmap, const unsigned char* hi, float* out_f32, void* out_bf16,
                  long long pre, long long L, long long post, int block_size,
                  int in_ebits, int in_mbits, int in_emax, float in_max_norm,
                  int out_ebits, int out_mbits, int out_emax, float out_max_norm, int plain, void* stream);

/* ---- GPTQ uniform quantizer -------------------------------------------------------------
 * replaces quantize(x, scale, zero, maxq)        utils/quant.py:268-272
 *          Quantizer.find_params                 utils/quant.py:298-357
 * x is viewed as [pre, C, post] with one (scale, zero) per C; find_params works on rows of x [rows, n]. */
int msq_uniform_quantize_f32(const float* x, float* out, const float* scale, const float* zero, int maxq,
                             long long pre, long long C, long long post, void* stream);
int msq_uniform_find_params_f32(const float* x, float* scale, float* zero, long long rows, long long n,
                                int maxq, int sym, int mse, float norm, int grid, float maxshrink, void* stream);

/* _shared_exponents(method="max")               utils/quant.py:498-541  (exact floor(log2)) */
int msq_shared_exponents_f32(const float* x, float* out, long long pre, long long L, long long post,
                             int ebits, void* stream);

/* ---- tensor-core GEMM / Hessian ---------------------------------------------------------
 * C[M,N] fp32 = alpha * A[M,K] * B[N,K]^T + beta * C ; A, B 16-bit (dtype 1 fp16, 2 bf16), K contiguous.
 * replaces the dense contractions the reference sends to cuBLAS SGEMM (llm/gptq.py:58, mx/linear.py:91). */
int msq_gemm_nt_f32out(const void* A, const void* B, float* C, int M, int N, int K, long long lda,
                       long long ldb, long long ldc, float alpha, float beta, int dtype, void* stream);

/* GPTQ.add_batch for nn.Linear                  llm/gptq.py:53-58
 *   H <- beta * H + alpha * X^T X,  X [T, K] tokens x features (dtype 0 fp32, 1 fp16, 2 bf16), H [K,K] fp32.
 *   caller passes beta = n/(n+b), alpha = 2/(n+b); workspace from msq_hessian_workspace_bytes. */
long long msq_hessian_workspace_bytes(int T, int K, int dtype);
int msq_hessian_accumulate(const void* X, int T, int K, long long ldx, int dtype, float* H, long long ldh,
                           float alpha, float beta, void* workspace, long long workspace_bytes,
                           int symmetric, void* stream);

/* ---- GPTQ solver ------------------------------------------------------------------------
 * msq_gptq_block: columns [i1, i1+count) of GPTQ.fasterquant's loop   llm/gptq.py:106-160
 *   (quantize_mx_outlier_hessian per column, loss, rank-1 updates inside the 128-column block).
 *   W [rows, ldw] current weights (read), Q out, Hinv upper Cholesky factor of H^-1, Err [rows,128] out,
 *   loss device double (+=), zeros/kcnt [cols] int32 (+=, caller zeroes once per linear), status OR.
 * msq_gptq_trailing_update: W[:, i2:] -= Err @ Hinv[i1:i1+count, i2:]   llm/gptq.py:165
 * msq_gptq_check_prune: *violation += #columns whose zero count is below the reference's k
 *   (llm/gptq.py:146-153 would then change values -> use the column-synchronous path).
 * msq_rank1_sub_f32: W[:, :n] -= err (x) h                              llm/gptq.py:159 */
int msq_gptq_block(const float* W, long long ldw, float* Q, long long ldq, const float* Hinv, long long ldh,
                   int rows, int i1, int count, float* Err, double* loss, int* zeros, int* kcnt, int* status,
                   int block_size, int in_scale_bits, int out_scale_bits,
                   int in_ebits, int in_mbits, int in_emax, float in_max_norm,
                   int out_ebits, int out_mbits, int out_emax, float out_max_norm,
                   float std_dev, int flush_fp32_subnorms, int rmode, void* stream);
/* msq_gptq_solve_f32: the whole column loop of one linear (llm/gptq.py:106-165) queued on `stream` without host
 *   round trips: per 128-column block the two-warp column kernel (rank-1 updates off the critical path, same fp32
 *   operation order per weight as the reference), then the trailing update on tcgen05 when ut_planes (from
 *   msq_cholinv_ut_planes, or msq_split3_planes_f32 of Hinv^T) is given, else in fp32 SIMT; then the prune check. */
long long msq_gptq_solve_workspace_bytes(int rows);
int msq_gptq_solve_f32(float* W, long long ldw, float* Q, long long ldq, const float* Hinv, long long ldh,
                       const void* ut_planes, long long ut_ld, int ut_plane_rows, int rows, int cols,
                       void* workspace, long long workspace_bytes, double* loss, int* zeros, int* kcnt,
                       int* status, int* violation, int block_size, int in_scale_bits, int out_scale_bits,
                       int in_ebits, int in_mbits, int in_emax, float in_max_norm,
                       int out_ebits, int out_mbits, int out_emax, float out_max_norm,
                       float std_dev, int flush_fp32_subnorms, int rmode,
                       unsigned char* codes8, unsigned int* bitmap, signed char* e_in, signed char* e_out,
                       void* stream);
/* codes8 .. e_out (optional, all or none; cols % 32 == 0, rows % block_size == 0): the quantizer's own artefact of the
 * GPTQ result -- one code byte per weight (inlier or outlier format by its bitmap bit), the outlier bitmap and the
 * block exponents [rows / block_size, cols].  msq_pack_from_codes turns code bytes + bitmap into the standard packed
 * planes (inlier-width slots, segment-compacted hi stream, seg_cnt) of ANY [pre, L, post] geometry, so that
 * msq_unpack_mx(planes) == Q bit for bit and msq_gemm_packed_rowblock_bf16 can run the GPTQ result directly. */
int msq_pack_from_codes(const unsigned char* codes8, const unsigned int* bitmap, long long pre, long long L,
                        long long post, int block_size, int in_ebits, int in_mbits, int out_ebits, int out_mbits,
                        unsigned char* slots, unsigned char* hi, int* seg_cnt, void* stream);
int msq_gptq_trailing_update(float* W, long long ldw, const float* Err, const float* Hinv, long long ldh,
                             int rows, int i1, int count, int i2, int cols, void* stream);
int msq_gptq_check_prune(const int* zeros, const int* kcnt, int cols, int* violation, void* stream);
int msq_rank1_sub_f32(float* W, long long ldw, const float* err, const float* h, int rows, int n, void* stream);

/* ---- fp32-accurate tensor-core contraction (bf16 plane stacks) ------------------------------
 * x = p0 + p1 + p2 exactly with bf16 planes; the six significant plane pairs are accumulated by
 * tcgen05.mma kind::f16 into the fp32 TMEM accumulator (every product exact), so the result has
 * fp32-level error.  replaces the fp32 cuBLAS/cuSOLVER contractions of llm/gptq.py:101-103,165 and
 * number_system/mx/linear.py:91.
 * msq_split3_planes_f32: src fp32 [rows, cols] -> planes[(p * plane_rows + r) * ldp + c], columns
 *   zero-filled up to the next multiple of 64 (ldp >= that, ldp % 8 == 0).
 * msq_gemm_split3_f32:   C[M,N] = alpha * A[M,K] B[N,K]^T + beta * C. */
int msq_split3_planes_f32(const float* src, long long ld, int rows, int cols, void* planes, long long ldp,
                          int plane_rows, void* stream);
int msq_gemm_split3_f32(const void* a_planes, long long a_ld, int a_plane_rows, const void* b_planes,
                        long long b_ld, int b_plane_rows, float* C, long long ldc, int M, int N, int K,
                        float alpha, float beta, void* stream);

/* ---- Cholesky trio of GPTQ.fasterquant -------------------------------------------------------
 * replaces  H = cholesky(H); H = cholesky_inverse(H); H = cholesky(H, upper=True)   llm/gptq.py:101-103
 * H fp32 [n, n] (ld = ldh, % 4), symmetric positive definite, upper triangle read, DESTROYED;
 * U fp32 [n, n] (ld = ldu, % 4) receives the upper-triangular factor with U^T U = H^-1 (zeros below
 * the diagonal).  One reverse Cholesky H = R R^T (R upper) + one triangular inverse U = R^-1, blocked
 * recursively with every panel product on tcgen05 (split-plane GEMM).  info: device int32 zeroed by
 * the caller, receives 1 + index of the first non-positive pivot (LAPACK's info), U then holds NaNs.
 * The workspace keeps bf16 planes of U^T afterwards (msq_cholinv_ut_planes): operand of
 * msq_gptq_solve_f32's tensor-core trailing update. */
long long msq_cholinv_workspace_bytes(int n);
int msq_cholinv_upper_f32(float* H, long long ldh, float* U, long long ldu, int n, void* workspace,
                          long long workspace_bytes, int* info, void* stream);
int msq_cholinv_ut_planes(void* workspace, int n, void** planes, long long* ld, int* plane_rows);

/* ---- packed-weight linear ---------------------------------------------------------------
 * C[M,N] fp32 (+)= X[M,K] (bf16) * dequant(W)[N,K]^T, W = packed planes blocked by 32 along K with equal
 * inlier/outlier code widths.  replaces LinearFunction.forward's fake-quant + F.linear,
 * number_system/mx/linear.py:78-91. */
int msq_gemm_packed_w_bf16(const void* X, long long ldx, const unsigned char* codes, const unsigned int* bitmap,
                           const signed char* e_in, const signed char* e_out, float* C, long long ldc,
                           int M, int N, int K, int in_ebits, int in_mbits, int out_ebits, int out_mbits,
                           int plain, int accumulate, void* stream);

/* Same contraction for the artefact the quantizers emit (utils/quant.py:147-266 with axes=[0]; what GPTQ.fasterquant
 * leaves, llm/llama.py:130-137): W packed in ROW blocks of block_size (16 or 32) output rows per input column,
 * inlier-width slots + bitmap + e_in/e_out [N/block_size, K] + the segment-compacted hi stream.
 * msq_packed_build_hi_index: rank index over the bitmap (stream offset of every 8th block inside its segment),
 *   built once per weight; hi_off8 holds ceil(nblocks / 8) uint32.
 * msq_gemm_packed_rowblock_bf16: N % block_size == 0, K % 64 == 0, slot width 2 or 4 bits (else
 *   MSQ_ERR_UNSUPPORTED -> msq_unpack_mx + msq_gemm_nt_f32out). */
int msq_packed_build_hi_index(const unsigned int* bitmap, long long pre, long long L, long long post, int block_size,
                              unsigned int* hi_off8, void* stream);
int msq_gemm_packed_rowblock_bf16(const void* X, long long ldx, const unsigned char* codes, const unsigned int* bitmap,
                                  const signed char* e_in, const signed char* e_out, const unsigned char* hi,
                                  const unsigned int* hi_off8, long long hi_seg_stride, float* C, long long ldc,
                                  int M, int N, int K, int block_size, int in_ebits, int in_mbits, int out_ebits,
                                  int out_mbits, int accumulate, void* stream);

/* native block-scaled tensor-core path: tcgen05.mma kind::mxf8f6f4.block_scale (E4M3 x E4M3, UE8M0 scales
 * per 32 elements along K).  X: plain MX-FP8 (E4M3) activations, codes [M,K] bytes (ld = ldx) + int8
 * exponents [M, K/32]; W: packed FP4 (E2M1) weight blocked by 32 along K, outliers folded into the E4M3
 * element value as q * 2^(e_out - 2 e_in) (caller guarantees -6 <= e_out - 2 e_in <= 6 for outlier blocks).
 * Same contraction as msq_gemm_packed_w_bf16; replaces number_system/mx/linear.py:66-91 when
 * a_elem_format = fp8_e4m3 and w_elem_format = fp4_e2m1. */
int msq_gemm_mxfp8_blockscaled(const unsigned char* xcodes, long long ldx, const signed char* xe,
                               const unsigned char* wcodes, const unsigned int* wbitmap, const signed char* we_in,
                               const signed char* we_out, float* C, long long ldc, int M, int N, int K,
                               int plain, int accumulate, void* stream);

/* FP4-by-TMA variant (decode sizes): the packed 4-bit codes of the weight ARE the tensor-core operand -- TMA expands
 * 16 codes into a 16-byte chunk (CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B, the layout kind::mxf8f6f4 reads E2M1 from), the
 * block scale is we_in, no thread touches an inlier -- and the outlier contribution is a fused sparse correction in the
 * epilogue: CSR over output features (o_ptr [N+1], o_k, o_dw = q * (2^(e_out - e_in) - 2^e_in)); o_ptr NULL = none.
 * Same contraction and operands as msq_gemm_mxfp8_blockscaled, without its -6 <= e_out - 2 e_in <= 6 restriction.
 * Needs K % 128 == 0, wcodes 32-byte aligned. */
int msq_gemm_mxfp4w_blockscaled(const unsigned char* xcodes, long long ldx, const signed char* xe,
                                const unsigned char* wcodes, const signed char* we_in, const int* o_ptr,
                                const int* o_k, const float* o_dw, float* C, long long ldc, int M, int N, int K,
                                int accumulate, void* stream);

/* GEMM-ready layout of a K-blocked FP4 weight for the decode sizes: per (128-row tile, 128-wide K block) one contiguous
 * 8 KB chunk of codes and one 512-byte chunk of UE8M0 scales in the tcgen05.cp layout, so that a pipeline stage is two
 * contiguous DRAM reads instead of 128 64-byte row segments and 128 scattered scale words.  Built once per weight
 * (the reference has no packed inference path: number_system/mx/linear.py:66-91 multiplies fake-quantized tensors).
 *   codes_tiled: ceil(N/128) * (K/128) * 8192 bytes, 32-byte aligned;  sf_tiled: ceil(N/128) * (K/128) * 512 bytes. */
int msq_tile_fp4_weight(const unsigned char* wcodes, const signed char* we_in, unsigned char* codes_tiled,
                        unsigned char* sf_tiled, int N, int K, void* stream);
int msq_gemm_mxfp4w_tiled_blockscaled(const unsigned char* xcodes, long long ldx, const signed char* xe,
                                      const unsigned char* codes_tiled, const unsigned char* sf_tiled,
                                      const int* o_ptr, const int* o_k, const float* o_dw, float* C, long long ldc,
                                      int M, int N, int K, int accumulate, void* stream);

/* msq_gemm_mxfp4w_tiled_blockscaled over STACKED activation terms (msq_split_a16_e4m3 below: M = terms * fold_m rows):
 * row m is added into C[m % fold_m, :] by the epilogue (C [fold_m, N] is zero-filled by the call), so a W4A16 linear at
 * decode sizes is the split launch + this call. */
int msq_gemm_mxfp4w_tiled_folded(const unsigned char* xcodes, long long ldx, const signed char* xe,
                                 const unsigned char* codes_tiled, const unsigned char* sf_tiled, const int* o_ptr,
                                 const int* o_k, const float* o_dw, float* C, long long ldc, int M, int fold_m, int N,
                                 int K, void* stream);

/* compute-bound variant of the block-scaled path: widen the packed FP4 weight once to dense MX-FP8
 * (E4M3 bytes [N,K] + UE8M0 scales [N,K/32]), then run the GEMM with both operands streamed by TMA. */
int msq_widen_fp4_to_mxfp8(const unsigned char* wcodes, const unsigned int* wbitmap, const signed char* we_in,
                           const signed char* we_out, unsigned char* out_bytes, unsigned char* out_sf,
                           int N, int K, int plain, void* stream);
int msq_gemm_mxfp8_dense(const unsigned char* xcodes, long long ldx, const signed char* xe,
                         const unsigned char* wbytes, long long ldw, const unsigned char* wsf, float* C,
                         long long ldc, int M, int N, int K, int accumulate, void* stream);

/* compute-bound shape of the same path with NO thread on an operand: element tiles by tensor loads, the UE8M0 scales of
 * BOTH operands by bulk copies from copies already in the tcgen05.cp chunk layout, two accumulators in tensor memory
 * (the epilogue of a tile overlaps the MMAs of the next), persistent CTAs over an L2-friendly tile raster.
 * Replaces number_system/mx/linear.py:66-91 (f_linear on fake-quantized operands) at prefill / calibration sizes.
 *   msq_tile_scales_ue8m0: scales [rows, K/32] (int8 shared exponents when is_int8_exponent, else UE8M0 bytes) ->
 *     ceil(rows / rows_per_tile) * (K/128) * ceil(rows_per_tile/128) chunks of 512 bytes.  The weight's copy uses
 *     rows_per_tile = 128 and is built once per weight.
 *   msq_gemm_mxfp8_tma: w = dense E4M3 bytes [N, K] (ld = ldw bytes; msq_widen_fp4_to_mxfp8) or, w_is_fp4 != 0, the packed
 *     4-bit codes of a plain FP4 weight [N, K/2 bytes]; wsf_tiled = its tiled scales; xe = int8 activation exponents
 *     [M, K/32], tiled into `workspace` (msq_gemm_mxfp8_tma_workspace_bytes) by a launch in front of the GEMM.
 *     o_ptr / o_k / o_dw (may be NULL): CSR of the outlier corrections of an FP4 weight, as in
 *     msq_gemm_mxfp4w_blockscaled, applied by the epilogue warps to the accumulator tile they hold in registers.
 *     terms (1..3): the activations are `terms` stacked slabs of M rows (msq_split_a16_e4m3: row t * M + m) whose
 *     products are ACCUMULATED into C [M, N] -- the K loop runs over (term, K block) -- so a W4A16 linear needs neither a
 *     [terms * M, N] result nor a pass that adds its slabs (o_ptr must be NULL then: msq_sparse_outlier_correct with
 *     fold_m = M applies the corrections).
 *     K % 128 == 0; w 32-byte aligned.  msq_gemm_mxfp8_tma_tile_tokens reports the token-tile width it will use. */
/* the sparse part as a pass of its own (every SM works on it): C[m (mod fold_m if > 0), o_n[e]] += dq(x)[m, o_k[e]] * o_dw[e]
 * for every outlier e (COO triplets) and token m, by atomics, behind a GEMM that ran on the packed codes alone. */
int msq_sparse_outlier_correct(const unsigned char* xcodes, long long ldx, const signed char* xe, const int* o_n,
                               const int* o_k, const float* o_dw, int nnz, float* C, long long ldc, int M, int K,
                               int fold_m, void* stream);
/* the same behind the W4A4 GEMM (msq_gemm_mxfp4_tma): xcodes4 = packed FP4 activations, ldx in bytes */
int msq_sparse_outlier_correct_fp4(const unsigned char* xcodes4, long long ldx, const signed char* xe, const int* o_n,
                                   const int* o_k, const float* o_dw, int nnz, float* C, long long ldc, int M, int K,
                                   int fold_m, void* stream);
int msq_tile_scales_ue8m0(const void* scales, int is_int8_exponent, int rows, int K, int rows_per_tile,
                          unsigned char* out, void* stream);
int msq_gemm_mxfp8_tma_tile_tokens(int M, int N, int K);
long long msq_gemm_mxfp8_tma_workspace_bytes(int M, int N, int K);
int msq_gemm_mxfp8_tma(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* w,
                       long long ldw, int w_is_fp4, const unsigned char* wsf_tiled, const int* o_ptr, const int* o_k,
                       const float* o_dw, float* C, long long ldc, int M, int N, int K, int terms, int accumulate,
                       void* workspace, long long workspace_bytes, void* stream);
/* ..._bf16: one term, the result as bf16 [M, ldc] rounded once from the fp32 accumulator (no accumulate, no fused sparse
 * epilogue -- apply msq_sparse_outlier_correct to an fp32 result instead --, no K slices); own workspace query */
long long msq_gemm_mxfp8_tma_bf16_workspace_bytes(int M, int N, int K);
int msq_gemm_mxfp8_tma_bf16(const unsigned char* xcodes, long long ldx, const signed char* xe, const unsigned char* w,
                            long long ldw, int w_is_fp4, const unsigned char* wsf_tiled, void* C_bf16, long long ldc,
                            int M, int N, int K, void* workspace, long long workspace_bytes, void* stream);
/* W4A4 (replaces F.linear on the two fake-quantized operands of number_system/mx/linear.py:66-91 when a_elem_format and
 * w_elem_format are both fp4_e2m1): BOTH operands packed FP4 (two codes per byte, low nibble = lower k, K-major; ldx / ldw in
 * bytes), activations with int8 shared exponents [M, K/32] (tiled into `workspace`, msq_gemm_mxfp4_tma_workspace_bytes), weight with
 * pre-tiled UE8M0 chunks: tcgen05.mma kind::mxf4.block_scale.scale_vec::2X (K = 64 per instruction, twice the tensor rate
 * of kind::mxf8f6f4).  Products and fp32 accumulation as the tensor core does them: equal to the fp32 matmul of the
 * dequantized operands up to summation order.  K % 256 == 0; 16-byte aligned operands and pitches. */
int msq_gemm_mxfp4_tma_tile_tokens(int M, int N, int K);
/* K slices (split-K through atomics into a zero-filled C) the plan of the TMA-fed kernels gives a problem; 1 = none */
int msq_gemm_mx_tma_k_slices(int M, int N, int K, int w4a4);
long long msq_gemm_mxfp4_tma_workspace_bytes(int M, int N, int K);
/* ..._bf16: the result as bf16 [M, ldc] (rounded once from the fp32 accumulator) for callers that want 16-bit activations
 * back; own workspace query (no K slices there) */
long long msq_gemm_mxfp4_tma_bf16_workspace_bytes(int M, int N, int K);
int msq_gemm_mxfp4_tma_bf16(const unsigned char* xcodes4, long long ldx, const signed char* xe, const unsigned char* w4,
                            long long ldw, const unsigned char* wsf_tiled, void* C_bf16, long long ldc, int M, int N,
                            int K, void* workspace, long long workspace_bytes, void* stream);
int msq_gemm_mxfp4_tma(const unsigned char* xcodes4, long long ldx, const signed char* xe, const unsigned char* w4,
                       long long ldw, const unsigned char* wsf_tiled, float* C, long long ldc, int M, int N, int K,
                       int accumulate, void* workspace, long long workspace_bytes, void* stream);

/* W4A16 on the block-scaled tensor-core path: 16-bit activations as a sum of T = 2 or 3 block-scaled E4M3 terms
 *     x[m, k] = sum_{t<T} a_t[m, k] * 2^(e_t[m, k/32]),  e_t = e_1 - 3 (t - 1),  e_1 = exponent(block max) - 7,
 * each term the truncation of the remainder onto the E4M3 grid (exact fp32 steps).  T = 2 (8 significant bits): bf16
 * elements within 2^12 of their block's maximum are reproduced exactly, smaller ones are truncated at 2^-19 of the block
 * maximum; T = 3 (12 bits): bf16 within 2^15, fp16 within 2^11, truncation at 2^-22.
 * The terms are stacked along the token dimension (row t * M + m of codes [T M, K] / e [T M, K/32]): the block-scaled
 * GEMMs above run once with T M tokens -- the packed 4-bit weight codes are the other operand, read once -- and the caller
 * adds the T M-row slabs.  Replaces the bf16 fake-quant matmul of number_system/mx/linear.py:78-91 for
 * a_elem_format = None (16-bit activations) without dequantizing the weight.
 * dtype: 1 = fp16, 2 = bf16; K % 32 == 0; x 16-byte aligned, ldx % 8 == 0. */
int msq_split_a16_e4m3(const void* x, long long ldx, int dtype, int M, int K, int terms, unsigned char* codes,
                       signed char* e, void* stream);

/* ---- host-buffer entry points --------------------------------------------------------------
 * quantize_mx_outlier_v1 (utils/quant.py:147-266) as the reference's drivers call it on CPU-resident
 * weights (llm/llama.py:94-121): h_in / h_out are HOST pointers (pinned for full PCIe speed), the
 * tensor is cut into chunks of ~chunk_bytes (0 = 16 MiB) along its outermost blocked dimension and
 * upload / kernel / download of successive chunks overlap on three streams.  Results are bit-identical
 * to msq_quantize_mx_outlier_f32 on the whole tensor (MX blocks are independent).
 * sync = 0 only queues the work: calls may be issued back to back without draining the pipeline;
 * msq_host_sync waits for everything queued on the current device and returns the OR of the MSQ_ST_*
 * bits raised since the previous msq_host_sync. */
int msq_host_quantize_mx_outlier_f32(const float* h_in, float* h_out, long long pre, long long L, long long post,
                                     int block_size, int in_scale_bits, int out_scale_bits,
                                     int in_ebits, int in_mbits, int in_emax, float in_max_norm,
                                     int out_ebits, int out_mbits, int out_emax, float out_max_norm,
                                     float std_dev, int flush_fp32_subnorms, int rmode,
                                     long long chunk_bytes, int sync);
int msq_host_sync(int* status_out);
int msq_host_release(void);

/* (m, k) such that floor(n / d) == ((unsigned long long)n * m) >> k for all 0 <= n < 2^31 -- the exact
 * divider behind the kernels' block-linear index arithmetic (replaces the integer divisions of the
 * [pre, L, post] view, utils/quant.py:544-616); exported for host-side testing. */
int msq_magic_div(long long d, unsigned int* m, int* k);

#ifdef __cplusplus
}
#endif
#endif /* MSQ_B200_H */
