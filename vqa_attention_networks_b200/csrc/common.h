// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vqa_b200.h"

namespace vqa {

int set_error(int code, const char* fmt, ...);   // records a thread-local message, returns code
int sm_count();                                   // SMs of the current device (cached per device)

#define VQA_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return ::vqa::set_error((int)_e, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define VQA_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) return ::vqa::set_error((int)_e, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vqa
