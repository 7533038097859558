// msq_api_internal.h -- includes shared by the translation units of libmsq_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

namespace msq {

// Every device-pointer entry point runs on the device that OWNS its first buffer, whatever the calling thread's
// current device is (the reference's torch ops work on any device; the stream handle the caller passes belongs to
// that device as well).  The previous current device is restored on return.
struct DevGuard {
  int prev = -1;
  explicit DevGuard(const void* p) {
    cudaPointerAttributes a;
    if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
      int cur = 0;
      if (cudaGetDevice(&cur) == cudaSuccess && cur != a.device) {
        prev = cur;
        cudaSetDevice(a.device);
      }
    } else {
      cudaGetLastError();      // a host pointer is not an error here; the entry point's own checks decide
    }
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DevGuard(const DevGuard&) = delete;
  DevGuard& operator=(const DevGuard&) = delete;
};

}  // namespace msq
