// common.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/walker_b200.h"

namespace wb {

// last error message of the calling thread (wb_last_error)
char* last_error_buffer();
int32_t fail(int32_t code, const char* fmt, ...);
// true when a usable sm_100 device is current; otherwise records the error
int32_t require_device();

#define WB_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return wb::fail(WB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define WB_REQUIRE(cond, msg)                             \
  do {                                                    \
    if (!(cond)) return wb::fail(WB_ERR_INVALID, "%s", msg); \
  } while (0)

}  // namespace wb
