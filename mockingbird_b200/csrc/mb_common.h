// Shared host-side helpers for libmockingbird_b200 (error reporting, launch counting).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/mockingbird_b200.h"

namespace mb {

// thread-local error string behind mb_last_error()
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
// every kernel launch of the library goes through this counter (mb_launch_count())
void count_launch(int n = 1);

#define MB_CUDA_CHECK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::mb::fail(MB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                 \
  } while (0)

#define MB_LAUNCH_CHECK(what)                                                              \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::mb::fail(MB_ERR_CUDA, "launch of %s failed: %s (%s:%d)", what,              \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                       \
    ::mb::count_launch();                                                                  \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mb
