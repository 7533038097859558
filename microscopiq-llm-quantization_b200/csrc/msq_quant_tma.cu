// msq_quant_tma.cu -- the MX+outlier fake-quantizer of 2-D weights blocked along the rows (axes=[0], the layout of the
// reference's GPTQ / RTN drivers, utils/quant.py:147-266) as a PERSISTENT, TMA-staged kernel.
//
// A tile = BS rows x TW columns of fp32 = the MX blocks of TW consecutive columns.  One producer thread per CTA moves
// tiles global -> shared (cp.async.bulk.tensor, mbarrier complete_tx) and shared -> global (bulk_group) through a ring
// of NST stages; TW consumer threads (thread = column = one MX block) read their block from the tile with immediate-
// offset LDS, quantize it in registers with the same core_outlier as every other mapping, and write the result back
// IN PLACE.  Consequences against k_strided_multi:
//   * no per-element address arithmetic (32 IMAD.WIDE per block gone), no per-CTA table search / division in the
//     compute threads (tile -> (matrix, row block, column) is decoded by the producer thread only);
//   * global-memory latency is covered by the ring, not by occupancy: consumers never wait on a load they issued;
//   * ragged edges need no code: out-of-bounds rows / columns are zero-filled by the TMA load (the reference's zero
//     padding, utils/quant.py:563-583) and clipped by the TMA store.
// Requirements: cols % 4 == 0 (16-byte global strides), 16-byte aligned bases; otherwise the caller uses k_strided_multi.
#include "msq_quant_kernels.cuh"
#include "msq_tc.cuh"

namespace msq {

constexpr int TQ_NST = 4;      // ring depth

struct TmaQParams {
  CUtensorMap tin[MSQ_MULTI_MAX], tout[MSQ_MULTI_MAX];
  QCfg c;
  int* status;
  int n;
  unsigned int total;                         // tiles of all matrices
  unsigned int tile_start[MSQ_MULTI_MAX + 1]; // first tile of matrix t
  int gx[MSQ_MULTI_MAX];                      // column tiles per row block
  unsigned int gx_m[MSQ_MULTI_MAX];           // magic divider of gx
  int gx_k[MSQ_MULTI_MAX];
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int BS, int TW, int MINB>
__global__ void __launch_bounds__(TW + 32, MINB) k_strided_tma(const __grid_constant__ TmaQParams P) {
  // stages first (the only shared memory of the kernel, so the declared alignment holds), barriers behind them
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw;
  constexpr int STAGE_FLOATS = BS * TW;
  constexpr uint32_t STAGE_BYTES = STAGE_FLOATS * 4;
  float* tiles = reinterpret_cast<float*>(base);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)TQ_NST * STAGE_BYTES);
  uint64_t* done = full + TQ_NST;
  const int tid = (int)threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TQ_NST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&done[s], TW); }
    tc::mbar_fence_init();
  }
  __syncthreads();
  const unsigned int G = gridDim.x, first = blockIdx.x;
  const int n_mine = first < P.total ? (int)((P.total - first + G - 1) / G) : 0;

  if (tid >= TW) {
    // ------------------------------------------------ producer: one thread moves every tile of this CTA both ways
    if (tid == TW) {
      auto decode = [&](int k, int& t, int& c0, int& c1) {
        const unsigned int id = first + (unsigned int)k * G;
        t = 0;
        while (t + 1 < P.n && id >= P.tile_start[t + 1]) ++t;
        const int local = (int)(id - P.tile_start[t]);
        const int b = fast_div(local, P.gx_m[t], P.gx_k[t]);
        c0 = (local - b * P.gx[t]) * TW;
        c1 = b * BS;
      };
      auto load = [&](int k) {
        int t, c0, c1;
        decode(k, t, c0, c1);
        const int s = k % TQ_NST;
        tc::mbar_expect_tx(&full[s], STAGE_BYTES);
        tc::tma_load_2d(tiles + (size_t)s * STAGE_FLOATS, &P.tin[t], &full[s], c0, c1);
      };
      const int pre = n_mine < TQ_NST ? n_mine : TQ_NST;
      for (int k = 0; k < pre; ++k) load(k);
      for (int k = 0; k < n_mine; ++k) {
        const int s = k % TQ_NST;
        tc::mbar_wait(&done[s], (uint32_t)(k / TQ_NST) & 1u);     // every consumer has written its column back
        int t, c0, c1;
        decode(k, t, c0, c1);
        tma_store_2d(&P.tout[t], tiles + (size_t)s * STAGE_FLOATS, c0, c1);
        bulk_commit();
        // the stage of tile k-1 (its store was committed one iteration ago) takes tile k-1+NST
        if (k >= 1 && k - 1 + TQ_NST < n_mine) {
          bulk_wait_read<1>();
          load(k - 1 + TQ_NST);
        }
      }
      bulk_wait_all();
    }
    return;
  }

  // -------------------------------------------------- consumers: thread = column = one MX block per tile
  int st = 0;
#pragma unroll 1
  for (int k = 0; k < n_mine; ++k) {
    const int s = k % TQ_NST;
    tc::mbar_wait(&full[s], (uint32_t)(k / TQ_NST) & 1u);
    float* tp = tiles + (size_t)s * STAGE_FLOATS + tid;
    float x[BS];
    uint32_t code[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) x[i] = tp[i * TW];
    const BlockOut bo = core_outlier<BS, 1, false, false, false>(x, code, P.c, nullptr, nullptr);
    st |= bo.st;
#pragma unroll
    for (int i = 0; i < BS; ++i) tp[i * TW] = x[i];
    fence_proxy_async_smem();        // generic-proxy writes -> visible to the bulk store
    tc::mbar_arrive(&done[s]);
  }
  if (P.status) {
    st = __reduce_or_sync(0xffffffffu, st);
    if ((tid & 31) == 0 && st) atomicOr(P.status, st);
  }
}

static int make_tmap_plain_f32(CUtensorMap* out, const void* basep, long long rows, long long cols, int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return MSQ_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(basep), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MSQ_OK : MSQ_ERR_CUDA;
}

template <int BS, int TW, int MINB>
static int launch_one(TmaQParams& P, cudaStream_t stream) {
  auto k = k_strided_tma<BS, TW, MINB>;
  const size_t smem = (size_t)TQ_NST * BS * TW * 4 + 2 * TQ_NST * sizeof(uint64_t) + 128;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MSQ_ERR_CUDA;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, TW + 32, smem) != cudaSuccess || occ < 1) return MSQ_ERR_CUDA;
  unsigned long long g = (unsigned long long)sms * (unsigned)occ;
  if (g > P.total) g = P.total;
  k<<<(unsigned)g, TW + 32, smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

// n matrices [rows[i], cols[i]] (device pointers in / out); returns MSQ_ERR_UNSUPPORTED when a matrix cannot be
// addressed by a tensor map (the caller then uses k_strided_multi)
int launch_strided_tma(int n, const float* const* in, float* const* out, const long long* rows, const long long* cols,
                       int bs, const QCfg& c, int* status, cudaStream_t stream) {
  if (n <= 0 || n > MSQ_MULTI_MAX) return MSQ_ERR_BAD_ARG;
  if (bs != 8 && bs != 16 && bs != 32) return MSQ_ERR_UNSUPPORTED;
  const int TW = bs == 32 ? 128 : 256;
  TmaQParams P;
  memset(&P, 0, sizeof(P));
  P.c = c; P.status = status; P.n = n;
  unsigned long long total = 0;
  for (int i = 0; i < n; ++i) {
    if ((cols[i] & 3) || (reinterpret_cast<uintptr_t>(in[i]) & 15) || (reinterpret_cast<uintptr_t>(out[i]) & 15))
      return MSQ_ERR_UNSUPPORTED;
    if (rows[i] <= 0 || cols[i] <= 0 || rows[i] >= 0x7fffffffLL || cols[i] >= 0x7fffffffLL) return MSQ_ERR_UNSUPPORTED;
    if (make_tmap_plain_f32(&P.tin[i], in[i], rows[i], cols[i], bs, TW)) return MSQ_ERR_UNSUPPORTED;
    if (make_tmap_plain_f32(&P.tout[i], out[i], rows[i], cols[i], bs, TW)) return MSQ_ERR_UNSUPPORTED;
    P.gx[i] = (int)((cols[i] + TW - 1) / TW);
    magic_div(P.gx[i], &P.gx_m[i], &P.gx_k[i]);
    P.tile_start[i] = (unsigned)total;
    total += (unsigned long long)P.gx[i] * (unsigned long long)((rows[i] + bs - 1) / bs);
    if (total >= 0x7fffffffULL) return MSQ_ERR_UNSUPPORTED;
  }
  P.tile_start[n] = (unsigned)total;
  P.total = (unsigned)total;
  if (bs == 16) return launch_one<16, 256, 3>(P, stream);
  if (bs == 32) return launch_one<32, 128, 3>(P, stream);
  return launch_one<8, 256, 4>(P, stream);
}

}  // namespace msq
