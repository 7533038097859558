// msq_host.cu -- host-buffer entry points: the quantizer fed from (pinned) host memory and read back to host memory.
//
// The reference's quantize_mx_outlier_v1 (utils/quant.py:147-266) is called on whole weight tensors that live on the host
// (llm/llama.py:94-121 walks a CPU-resident model layer by layer).  Over PCIe the copies dominate, so this entry point
// cuts the tensor into contiguous chunks along its outermost blocked dimension (MX blocks are independent, so chunk
// boundaries that are multiples of the block size give bit-identical results) and runs upload / kernel / download of
// successive chunks on three streams.  Calls may be queued back to back (sync = 0): the slots are recycled through
// events, so the pipeline never drains between tensors; msq_host_sync() waits and returns the status word.
#include "msq_api_internal.h"
#include "msq_common.cuh"
#include <mutex>

extern "C" int msq_quantize_mx_outlier_f32(
    const float* in, float* deq, long long pre, long long L, long long post, int block_size, int in_scale_bits,
    int out_scale_bits, int in_ebits, int in_mbits, int in_emax, float in_max_norm, int out_ebits, int out_mbits,
    int out_emax, float out_max_norm, float std_dev, int flush_fp32_subnorms, int rmode, int variant, const float* lo_ext,
    const float* hi_ext, unsigned char* codes, signed char* e_in, signed char* e_out, unsigned int* bitmap,
    unsigned char* hi, int* seg_cnt, unsigned char* blk_cnt, unsigned char* mask, int* status, int* host_path, void* stream);

namespace {

constexpr int NSLOT = 3;
constexpr int MAXDEV = 16;

struct Slot {
  float* d_in = nullptr;
  float* d_out = nullptr;
  size_t cap = 0;
  cudaEvent_t up = nullptr, done = nullptr, down = nullptr;   // upload finished / kernel finished / download finished
  bool used = false;
};

struct Pipe {
  bool init = false;
  cudaStream_t s_up = nullptr, s_k = nullptr, s_down = nullptr;
  Slot slot[NSLOT];
  int* d_status = nullptr;
  long long next = 0;       // chunks queued so far (slot = next % NSLOT)
};

Pipe g_pipe[MAXDEV];
std::mutex g_mu;

#define CK(x) do { if ((x) != cudaSuccess) return MSQ_ERR_CUDA; } while (0)

int pipe_init(Pipe& p) {
  if (p.init) return MSQ_OK;
  CK(cudaStreamCreateWithFlags(&p.s_up, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&p.s_k, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&p.s_down, cudaStreamNonBlocking));
  for (int i = 0; i < NSLOT; ++i) {
    CK(cudaEventCreateWithFlags(&p.slot[i].up, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p.slot[i].done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p.slot[i].down, cudaEventDisableTiming));
  }
  CK(cudaMalloc(&p.d_status, sizeof(int)));
  CK(cudaMemset(p.d_status, 0, sizeof(int)));
  p.init = true;
  return MSQ_OK;
}

int slot_reserve(Pipe& p, Slot& s, size_t bytes) {
  if (s.cap >= bytes) return MSQ_OK;
  if (s.used) CK(cudaEventSynchronize(s.down));        // the old buffers may still be in flight
  if (s.d_in) cudaFree(s.d_in);
  if (s.d_out) cudaFree(s.d_out);
  s.d_in = s.d_out = nullptr; s.cap = 0;
  CK(cudaMalloc(&s.d_in, bytes));
  CK(cudaMalloc(&s.d_out, bytes));
  s.cap = bytes;
  return MSQ_OK;
}

}  // namespace

extern "C" {

int msq_host_quantize_mx_outlier_f32(const float* h_in, float* h_out, long long pre, long long L, long long post,
                                     int block_size, int in_scale_bits, int out_scale_bits, int in_ebits, int in_mbits,
                                     int in_emax, float in_max_norm, int out_ebits, int out_mbits, int out_emax,
                                     float out_max_norm, float std_dev, int flush_fp32_subnorms, int rmode,
                                     long long chunk_bytes, int sync) {
  if (!h_in || !h_out || pre <= 0 || L <= 0 || post <= 0 || block_size <= 0) return MSQ_ERR_BAD_ARG;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= MAXDEV) return MSQ_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(g_mu);
  Pipe& p = g_pipe[dev];
  int rc = pipe_init(p);
  if (rc) return rc;
  if (chunk_bytes <= 0) chunk_bytes = 16ll << 20;

  // chunk along the outermost dimension that keeps MX blocks whole: whole [L, post] planes when pre > 1,
  // otherwise groups of block rows
  const bool by_pre = pre > 1;
  const long long unit_elems = by_pre ? L * post : (long long)block_size * post;      // smallest legal chunk
  const long long units = by_pre ? pre : (L + block_size - 1) / block_size;
  long long upc = chunk_bytes / (unit_elems * 4);                                      // units per chunk
  if (upc < 1) upc = 1;
  const long long total_elems = pre * L * post;

  for (long long u0 = 0; u0 < units; u0 += upc) {
    const long long u1 = (u0 + upc < units) ? u0 + upc : units;
    const long long e0 = u0 * unit_elems;
    const long long e1 = (u1 * unit_elems < total_elems) ? u1 * unit_elems : total_elems;
    const size_t bytes = (size_t)(e1 - e0) * 4;
    Slot& s = p.slot[p.next % NSLOT];
    rc = slot_reserve(p, s, bytes);
    if (rc) return rc;
    // upload may overwrite d_in once the kernel that last read it is done
    if (s.used) CK(cudaStreamWaitEvent(p.s_up, s.done, 0));
    CK(cudaMemcpyAsync(s.d_in, h_in + e0, bytes, cudaMemcpyHostToDevice, p.s_up));
    CK(cudaEventRecord(s.up, p.s_up));
    // kernel needs the upload, and d_out free again (its previous download finished)
    CK(cudaStreamWaitEvent(p.s_k, s.up, 0));
    if (s.used) CK(cudaStreamWaitEvent(p.s_k, s.down, 0));
    const long long c_pre = by_pre ? (u1 - u0) : 1;
    const long long c_L = by_pre ? L : (e1 - e0) / post;
    rc = msq_quantize_mx_outlier_f32(s.d_in, s.d_out, c_pre, c_L, post, block_size, in_scale_bits, out_scale_bits, in_ebits,
                                     in_mbits, in_emax, in_max_norm, out_ebits, out_mbits, out_emax, out_max_norm, std_dev,
                                     flush_fp32_subnorms, rmode, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                     nullptr, nullptr, nullptr, nullptr, p.d_status, nullptr, p.s_k);
    if (rc) return rc;
    CK(cudaEventRecord(s.done, p.s_k));
    CK(cudaStreamWaitEvent(p.s_down, s.done, 0));
    CK(cudaMemcpyAsync(h_out + e0, s.d_out, bytes, cudaMemcpyDeviceToHost, p.s_down));
    CK(cudaEventRecord(s.down, p.s_down));
    s.used = true;
    ++p.next;
  }
  if (sync) CK(cudaStreamSynchronize(p.s_down));
  return MSQ_OK;
}

// Wait for everything queued by the host entry points on the current device.  *status_out receives the OR of the
// MSQ_ST_* bits raised since the previous call (and the device word is cleared).
int msq_host_sync(int* status_out) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= MAXDEV) return MSQ_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(g_mu);
  Pipe& p = g_pipe[dev];
  if (!p.init) { if (status_out) *status_out = 0; return MSQ_OK; }
  CK(cudaStreamSynchronize(p.s_down));
  int st = 0;
  CK(cudaMemcpyAsync(&st, p.d_status, sizeof(int), cudaMemcpyDeviceToHost, p.s_k));
  CK(cudaMemsetAsync(p.d_status, 0, sizeof(int), p.s_k));
  CK(cudaStreamSynchronize(p.s_k));
  if (status_out) *status_out = st;
  return MSQ_OK;
}

// Release the staging buffers of the current device (tests, memory pressure).
int msq_host_release(void) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= MAXDEV) return MSQ_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(g_mu);
  Pipe& p = g_pipe[dev];
  if (!p.init) return MSQ_OK;
  CK(cudaStreamSynchronize(p.s_down));
  CK(cudaStreamSynchronize(p.s_k));
  for (int i = 0; i < NSLOT; ++i) {
    Slot& s = p.slot[i];
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    s.d_in = s.d_out = nullptr; s.cap = 0; s.used = false;
  }
  return MSQ_OK;
}

}  // extern "C"
