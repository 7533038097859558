/*
 * msq_b200.h -- C ABI of libmsq_b200.so, the B200 (sm_100a) implementation of the
 * MicroScopiQ quantization hot path.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller owns every buffer, nothing is allocated behind its back and there is no
 *     global state; `stream` is the caller's cudaStream_t (NULL = legacy default stream);
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
 *   status   OR of MSQ_ST_*;  host_path  which kernel ran (1 strided, 2 innermost, 3 generic) */
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

/* detector statistics of number_system/mx/mx_ops.py:49-83 as called at :248 (mean / unbiased
 * std of signed A over the block-count axis).  lo, hi: fp32 [pre, block_size, post]. */
int msq_outlier_bounds_across_blocks_f32(const float* in, float* lo, float* hi, long long pre,
                                         long long L, long long post, int block_size,
                                         float std_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSQ_B200_H */
