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
#include <stdlib.h>

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

// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or the time is up, so the
// producer warp does not take issue slots from the compute warps of its scheduler while it waits (a plain try_wait
// loop was measured at 14 % of all issued instructions)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {}
}

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
        mbar_wait_sleep(&done[s], (uint32_t)(k / TQ_NST) & 1u);   // every consumer has written its column back
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
    mbar_wait_sleep(&full[s], (uint32_t)(k / TQ_NST) & 1u);
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


// ---- quantize + pack (codes, bitmap, exponents, hi stream) in the same TMA-staged style ---------------------------
// Work unit = one packed SEGMENT (MSQ_SEG_BLOCKS = 1024 consecutive block-linear indices = 4 tiles of 256 columns; with
// post % 256 == 0 every tile lies inside one row block, a segment may cross into the next one).  Per tile the consumers
// leave the tile's slice of every plane in a staging area of shared memory -- 4-bit codes through the 8x8 byte transpose
// across lanes, one ballot per row for the bitmap, one byte per block for each exponent -- and the producer thread
// writes them with two tensor stores and two linear bulk stores: the compute threads do no global address arithmetic.
// The hi fields of a block (<= 64 bits) are compacted in registers and stashed; one scan per segment places them.
struct TmaPackParams {
  CUtensorMap tin, tcodes, tbitmap;
  QCfg c;
  int* status;
  int8_t* e_in;
  int8_t* e_out;
  uint8_t* hi;
  int* seg_cnt;
  long long hi_seg_stride;
  int hi_w;
  unsigned int ntiles, nseg;
  int gx;
  unsigned int gx_m;
  int gx_k;
};

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// exclusive scan over the 256 consumer threads (named barrier 1: the producer warp does not take part)
__device__ __forceinline__ int cons_excl_scan(int v, int* s_warp, int* sum) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int wt;
  const int ex = warp_excl_scan(v, lane, wt);
  cons_bar();
  if (lane == 0) s_warp[w] = wt;
  cons_bar();
  int basev = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int t = s_warp[i]; if (i < w) basev += t; tot += t; }
  *sum = tot;
  return ex + basev;
}

// store_codes4_16 (msq_quant_strided.cu) with the destination in shared memory: cd = the tile's code image, 16 rows of
// 128 bytes; this warp's 32 columns are bytes [16*warp, 16*warp + 16) of every row
// W0 / W1 = this column's slot nibbles of rows 0-7 / 8-15 (slots_from_codes16)
__device__ __forceinline__ void stage_slots16(uint32_t* cd, uint32_t W0, uint32_t W1, int lane, int warp) {
  const int g = lane & 7;
  {
    const uint32_t r = __shfl_xor_sync(0xffffffffu, (g & 4) ? W0 : W1, 4);
    if (g & 4) W0 = r; else W1 = r;
  }
  {
    const uint32_t r = __shfl_xor_sync(0xffffffffu, __byte_perm(W0, W1, (g & 2) ? 0x5410 : 0x7632), 2);
    W0 = __byte_perm(W0, r, (g & 2) ? 0x3254 : 0x5410);
    W1 = __byte_perm(W1, r, (g & 2) ? 0x3276 : 0x7610);
  }
  {
    const uint32_t r = __shfl_xor_sync(0xffffffffu, __byte_perm(W0, W1, (g & 1) ? 0x6420 : 0x7531), 1);
    W0 = __byte_perm(W0, r, (g & 1) ? 0x3514 : 0x5240);
    W1 = __byte_perm(W1, r, (g & 1) ? 0x3716 : 0x7260);
  }
  const uint32_t a0 = W0 & 0x0F0F0F0Fu, a1 = W1 & 0x0F0F0F0Fu;
  const uint32_t b0 = (W0 >> 4) & 0x0F0F0F0Fu, b1 = (W1 >> 4) & 0x0F0F0F0Fu;
  const uint32_t row_lo = __byte_perm(a0 | (a0 >> 4), a1 | (a1 >> 4), 0x6420);
  const uint32_t row_hi = __byte_perm(b0 | (b0 >> 4), b1 | (b1 >> 4), 0x6420);
  uint32_t* dst = cd + (2 * g) * 32 + warp * 4 + (lane >> 3);      // 32 words per row
  dst[0] = row_lo;
  dst[32] = row_hi;
}
__device__ __forceinline__ void stage_codes4_16(uint32_t* cd, const uint32_t* c, int lane, int warp) {
  uint32_t W0, W1;
  slots_from_codes16(c, W0, W1);
  stage_slots16(cd, W0, W1, lane, warp);
}

// bitmap of a tile without ballots: the 32 x 16 bit matrix (lane = column, bit = row) is transposed by a 5-stage
// butterfly -- lanes j and j^16 first share one 32-bit word [M_(j|16) : M_(j&15)], then four swap stages on the two
// 16 x 16 halves at once (one shuffle, one rotate, one bitwise select each) -- after which lane r < 16 holds row r's word
template <int D>
__device__ __forceinline__ uint32_t bfly_stage(uint32_t v, int lane) {
  constexpr uint32_t m = D == 8 ? 0x00FF00FFu : (D == 4 ? 0x0F0F0F0Fu : (D == 2 ? 0x33333333u : 0x55555555u));
  const uint32_t t = __shfl_xor_sync(0xffffffffu, v, D);
  const bool up = (lane & D) != 0;
  // upper lane: its left columns take the partner's right columns (t >> D); lower lane: the mirror image
  return up ? ((v & ~m) | ((t >> D) & m)) : ((v & m) | ((t << D) & ~m));
}
__device__ __forceinline__ void stage_bitmap16(uint32_t* bm_row0, uint32_t M, int lane) {
  const uint32_t o = __shfl_xor_sync(0xffffffffu, M, 16);
  uint32_t v = (lane & 16) ? (o | (M << 16)) : (M | (o << 16));
  v = bfly_stage<8>(v, lane);
  v = bfly_stage<4>(v, lane);
  v = bfly_stage<2>(v, lane);
  v = bfly_stage<1>(v, lane);
  if (lane < 16) bm_row0[lane * (256 / 32)] = v;
}

constexpr int TP_NST = 5;      // input ring
constexpr int TP_NOUT = 3;     // staging buffers of the output planes (a freed buffer is needed one whole tile later)
constexpr int TP_TW = 256;

template <int BS, bool CVT, int NST = TP_NST, int NOUT = TP_NOUT, int MINB = 2>
__global__ void __launch_bounds__(TP_TW + 32, MINB) k_pack_tma(const __grid_constant__ TmaPackParams P) {
  static_assert(BS == 16, "4-bit slots, 16-row blocks");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int STAGE_FLOATS = BS * TP_TW;
  constexpr uint32_t STAGE_BYTES = STAGE_FLOATS * 4;
  constexpr int CD_BYTES = BS * (TP_TW / 2), BM_BYTES = BS * (TP_TW / 8), EX_BYTES = TP_TW;
  constexpr int OUT_BYTES = CD_BYTES + BM_BYTES + 2 * EX_BYTES;      // staging per tile (3 KB), double-buffered
  float* tiles = reinterpret_cast<float*>(smem_raw);
  unsigned char* outs = smem_raw + (size_t)NST * STAGE_BYTES;
  unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(outs + NOUT * OUT_BYTES);
  uint32_t* s_hi = reinterpret_cast<uint32_t*>(s_acc + MSQ_SEG_BLOCKS);          // MSQ_SEG_BLOCKS * BS * 4 / 8 bytes max
  unsigned char* s_cnt = reinterpret_cast<unsigned char*>(s_hi + MSQ_SEG_BLOCKS * BS * 4 / 32);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_cnt + MSQ_SEG_BLOCKS);   // producer -> consumers: tile landed
  uint64_t* empty = full + NST;                                         // consumers -> producer: tile is in registers
  uint64_t* sdone = empty + NST;                                        // consumers -> producer: staging written
  uint64_t* sfree = sdone + NOUT;                                       // producer -> consumers: staging read out
  int* s_warp = reinterpret_cast<int*>(sfree + NOUT);
  const int tid = (int)threadIdx.x;
  const int hi_w = P.hi_w;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], TP_TW); }
#pragma unroll
    for (int s = 0; s < NOUT; ++s) { tc::mbar_init(&sdone[s], TP_TW); tc::mbar_init(&sfree[s], 1); }
    tc::mbar_fence_init();
  }
  for (int i = tid; i < MSQ_SEG_BLOCKS * BS * 4 / 32; i += TP_TW + 32) s_hi[i] = 0;
  for (int i = tid; i < MSQ_SEG_BLOCKS; i += TP_TW + 32) s_cnt[i] = 0;
  __syncthreads();
  const unsigned int G = gridDim.x, first = blockIdx.x;
  const int nseg_mine = first < P.nseg ? (int)((P.nseg - first + G - 1) / G) : 0;
  int n_mine = 4 * nseg_mine;
  if (nseg_mine > 0 && first + (unsigned)(nseg_mine - 1) * G == P.nseg - 1) n_mine -= (int)(4u * P.nseg - P.ntiles);

  if (tid >= TP_TW) {
    if (tid == TP_TW) {
      auto tile_of = [&](int k) { return 4u * (first + (unsigned)(k >> 2) * G) + (unsigned)(k & 3); };
      auto coords = [&](unsigned int q, int& cx, int& r0) {
        const int b = fast_div((int)q, P.gx_m, P.gx_k);
        cx = (int)q - b * P.gx;
        r0 = b * BS;
      };
      auto load = [&](int k) {
        int cx, r0;
        coords(tile_of(k), cx, r0);
        const int s = k % NST;
        tc::mbar_expect_tx(&full[s], STAGE_BYTES);
        tc::tma_load_2d(tiles + (size_t)s * STAGE_FLOATS, &P.tin, &full[s], cx * TP_TW, r0);
      };
      const int pre = n_mine < NST ? n_mine : NST;
      for (int k = 0; k < pre; ++k) load(k);
      int si = 0, so = 0;                    // input stage / staging buffer of tile k
      uint32_t pi = 0, po = 0;               // their phase parities
      for (int k = 0; k < n_mine; ++k) {
        // the input stage is free as soon as every consumer holds its block in registers: refill it right away
        if (k + NST < n_mine) {
          mbar_wait_sleep(&empty[si], pi);
          load(k + NST);
        }
        mbar_wait_sleep(&sdone[so], po);
        const unsigned int q = tile_of(k);
        int cx, r0;
        coords(q, cx, r0);
        const unsigned char* o = outs + (size_t)so * OUT_BYTES;
        tma_store_2d(&P.tcodes, o, cx * (TP_TW / 2), r0);
        tma_store_2d(&P.tbitmap, o + CD_BYTES, cx * (TP_TW / 8), r0);
        bulk_store_1d(P.e_in + (size_t)q * TP_TW, o + CD_BYTES + BM_BYTES, EX_BYTES);
        bulk_store_1d(P.e_out + (size_t)q * TP_TW, o + CD_BYTES + BM_BYTES + EX_BYTES, EX_BYTES);
        bulk_commit();
        if (k >= 1) {                     // the stores of tile k-1 have left their staging buffer
          bulk_wait_read<1>();
          tc::mbar_arrive(&sfree[so == 0 ? NOUT - 1 : so - 1]);
        }
        if (++si == NST) { si = 0; pi ^= 1u; }
        if (++so == NOUT) { so = 0; po ^= 1u; }
      }
      bulk_wait_all();
    }
    return;
  }

  const int lane = tid & 31, warp = tid >> 5;
  int st = 0;
  int s = 0, b2 = 0;                       // input stage / staging buffer of tile k
  uint32_t ph = 0, pf = 1;                 // parity of full[s]; parity that makes the first sfree wait pass at once
#pragma unroll 1
  for (int k = 0; k < n_mine; ++k) {
    const int u = k & 3;
    mbar_wait_sleep(&full[s], ph);
    const float* tp = tiles + (size_t)s * STAGE_FLOATS + tid;
    float x[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) x[i] = tp[i * TP_TW];
    tc::mbar_arrive(&empty[s]);
    BlockPk bo;
    if (CVT) {
      bo = core_outlier_cvt(x, P.c);
    } else {
      uint32_t code[BS];
      const BlockOut b = core_outlier<BS, 1, true, false, true, true>(x, code, P.c, nullptr, nullptr);
      slots_from_codes16(code, bo.s0, bo.s1);
      bo.h0 = b.hacc[0]; bo.h1 = b.hacc[1]; bo.mask = b.mask; bo.e_in = b.e_in; bo.e_out = b.e_out; bo.cnt = b.cnt; bo.st = b.st;
    }
    st |= bo.st;
    // staging buffer b2 was last used by tile k - NOUT: wait until its stores have read it
    mbar_wait_sleep(&sfree[b2], pf);
    unsigned char* o = outs + (size_t)b2 * OUT_BYTES;
    stage_slots16(reinterpret_cast<uint32_t*>(o), bo.s0, bo.s1, lane, warp);
    stage_bitmap16(reinterpret_cast<uint32_t*>(o + CD_BYTES) + warp, bo.mask, lane);
    o[CD_BYTES + BM_BYTES + tid] = (unsigned char)(int8_t)bo.e_in;
    o[CD_BYTES + BM_BYTES + EX_BYTES + tid] = (unsigned char)(int8_t)bo.e_out;
    if (hi_w > 0) {
      // the two 32-bit halves of the block's compacted hi fields, joined by the number of outliers in the low half
      const int sh = __popc(bo.mask & 0xFFu) * hi_w;                       // <= 32
      s_acc[u * TP_TW + tid] = (unsigned long long)bo.h0 | ((unsigned long long)bo.h1 << sh);
      s_cnt[u * TP_TW + tid] = (unsigned char)bo.cnt;
    }
    fence_proxy_async_smem();
    tc::mbar_arrive(&sdone[b2]);
    if (++s == NST) { s = 0; ph ^= 1u; }
    if (++b2 == NOUT) { b2 = 0; pf ^= 1u; }

    if (u == 3 || k == n_mine - 1) {
      const unsigned int sg = first + (unsigned)(k >> 2) * G;
      if (hi_w > 0) {
        if (u < 3)
          for (int i = (u + 1) * TP_TW + tid; i < MSQ_SEG_BLOCKS; i += TP_TW) s_cnt[i] = 0;
        cons_bar();
        int c4[4], mine = 0;
#pragma unroll
        for (int v = 0; v < 4; ++v) { c4[v] = s_cnt[4 * tid + v]; mine += c4[v]; }
        int tot;
        int off = cons_excl_scan(mine, s_warp, &tot);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (c4[v]) {
            const unsigned long long acc = s_acc[4 * tid + v];
            const int bit0 = off * hi_w;
            const int w = bit0 >> 5, b = bit0 & 31;
            const unsigned long long lo = acc << b;
            const uint32_t top = b ? (uint32_t)(acc >> (64 - b)) : 0u;
            atomicOr(&s_hi[w], (uint32_t)lo);
            if ((uint32_t)(lo >> 32)) atomicOr(&s_hi[w + 1], (uint32_t)(lo >> 32));
            if (top) atomicOr(&s_hi[w + 2], top);
            off += c4[v];
          }
        }
        cons_bar();
        const int words = (tot * hi_w + 31) >> 5;
        uint32_t* dst = reinterpret_cast<uint32_t*>(P.hi + (size_t)sg * P.hi_seg_stride);
        for (int i = tid; i < words; i += TP_TW) { dst[i] = s_hi[i]; s_hi[i] = 0; }
        if (tid == 0 && P.seg_cnt) P.seg_cnt[sg] = tot;
      } else if (tid == 0 && P.seg_cnt) {
        P.seg_cnt[sg] = 0;
      }
    }
  }
  if (P.status) {
    st = __reduce_or_sync(0xffffffffu, st);
    if (lane == 0 && st) atomicOr(P.status, st);
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
  if (bs == 16) {
    // MSQ_QUANT_TMA_OCC=2 (A/B): two CTAs per SM instead of three -- measured in bench.py on an uncapped box: 0.851 vs 0.875
    const char* eo = getenv("MSQ_QUANT_TMA_OCC");
    if (eo && eo[0] == '2') return launch_one<16, 256, 2>(P, stream);
    return launch_one<16, 256, 3>(P, stream);
  }
  if (bs == 32) return launch_one<32, 128, 3>(P, stream);
  return launch_one<8, 256, 4>(P, stream);
}


static int make_tmap_bytes(CUtensorMap* out, const void* basep, long long rows, long long row_bytes, int box_rows, int box_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return MSQ_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_bytes, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(basep), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MSQ_OK : MSQ_ERR_CUDA;
}

// quantize + pack of ONE 2-D weight [L, post] blocked along its rows: 16-row blocks, 4-bit slots, hi fields of a block
// within 64 bits, post % 256 == 0, no dequantized output.  MSQ_ERR_UNSUPPORTED -> the caller uses k_strided<.., PACK>.
int launch_pack_tma(const QParams& Q, cudaStream_t stream) {
  if (Q.pre != 1 || Q.bs != 16 || Q.c.fin.width != 4 || (Q.post & 255) || Q.deq || Q.mask || Q.blk_cnt || Q.c.variant != 0 ||
      Q.bs * Q.hi_w > 64 || Q.hi_w == 8 || !Q.codes || !Q.bitmap || !Q.e_in || !Q.e_out || (Q.hi_w > 0 && !Q.hi))
    return MSQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(Q.in) & 15) || (reinterpret_cast<uintptr_t>(Q.codes) & 15) ||
      (reinterpret_cast<uintptr_t>(Q.bitmap) & 15) || (reinterpret_cast<uintptr_t>(Q.e_in) & 15) ||
      (reinterpret_cast<uintptr_t>(Q.e_out) & 15) || (reinterpret_cast<uintptr_t>(Q.hi) & 3))
    return MSQ_ERR_UNSUPPORTED;
  if (Q.nblocks >= 0x7fffffffLL || (long long)Q.L * Q.post >= 0x7fffffffLL) return MSQ_ERR_UNSUPPORTED;
  TmaPackParams P;
  memset(&P, 0, sizeof(P));
  if (make_tmap_plain_f32(&P.tin, Q.in, Q.L, Q.post, 16, TP_TW)) return MSQ_ERR_UNSUPPORTED;
  if (make_tmap_bytes(&P.tcodes, Q.codes, Q.L, Q.post / 2, 16, TP_TW / 2)) return MSQ_ERR_UNSUPPORTED;
  if (make_tmap_bytes(&P.tbitmap, Q.bitmap, Q.L, Q.post / 8, 16, TP_TW / 8)) return MSQ_ERR_UNSUPPORTED;
  P.c = Q.c; P.status = Q.status; P.e_in = Q.e_in; P.e_out = Q.e_out; P.hi = Q.hi; P.seg_cnt = Q.seg_cnt;
  P.hi_seg_stride = Q.hi_seg_stride; P.hi_w = Q.hi_w;
  P.ntiles = (unsigned)(Q.nblocks / TP_TW);
  P.nseg = (unsigned)((Q.nblocks + MSQ_SEG_BLOCKS - 1) / MSQ_SEG_BLOCKS);
  P.gx = (int)(Q.post / TP_TW);
  magic_div(P.gx, &P.gx_m, &P.gx_k);
  // the drivers' pairing (fp4_e2m1 inliers, fp8_e4m3 outliers, round to nearest) rounds and encodes by the narrow-format
  // conversion instructions (core_outlier_cvt); MSQ_PACK_CVT=0 keeps the generic arithmetic for comparison
  const bool cvt = Q.c.fin.ebits == 2 && Q.c.fin.bits == 3 && Q.c.fin.max_norm == 6.0f && Q.c.fout.ebits == 4 && Q.c.fout.bits == 5 &&
                   Q.c.fout.max_norm == 448.0f && Q.c.fin.emax == 2 && Q.c.fout.emax == 8 && Q.c.rmode == RD_NEAREST && Q.hi_w == 4 &&
                   !(getenv("MSQ_PACK_CVT") && getenv("MSQ_PACK_CVT")[0] == '0');
  // MSQ_PACK_OCC3=1 (experiment): three CTAs per SM with a 3-stage ring and two staging buffers
  const bool occ3 = cvt && getenv("MSQ_PACK_OCC3") && getenv("MSQ_PACK_OCC3")[0] == '1';
  void (*k)(const TmaPackParams) = cvt ? (occ3 ? k_pack_tma<16, true, 3, 2, 3> : k_pack_tma<16, true>) : k_pack_tma<16, false>;
  const int nst = occ3 ? 3 : TP_NST, nout = occ3 ? 2 : TP_NOUT;
  const size_t smem = (size_t)nst * (16 * TP_TW * 4) + nout * (16 * (TP_TW / 2) + 16 * (TP_TW / 8) + 2 * TP_TW) +
                      (size_t)MSQ_SEG_BLOCKS * 8 + (size_t)MSQ_SEG_BLOCKS * 16 * 4 / 8 + MSQ_SEG_BLOCKS +
                      (2 * nst + 2 * nout) * sizeof(uint64_t) + 8 * sizeof(int) + 64;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MSQ_ERR_CUDA;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, TP_TW + 32, smem) != cudaSuccess || occ < 1) return MSQ_ERR_CUDA;
  unsigned long long g = (unsigned long long)sms * (unsigned)occ;
  if (g > P.nseg) g = P.nseg;
  k<<<(unsigned)g, TP_TW + 32, smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? MSQ_OK : MSQ_ERR_CUDA;
}

}  // namespace msq
