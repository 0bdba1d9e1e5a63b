// Shared helpers for libfrcnn_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/frcnn_b200.h"

namespace frcnn {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

extern thread_local char g_last_error[512];

inline int fail(int code, const char *msg)
{
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}

inline int cuda_fail(cudaError_t e, const char *where)
{
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

// launch check: cudaGetLastError only (never synchronises)
#define FRCNN_CHECK_LAUNCH(where)                              \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) return frcnn::cuda_fail(e__, where); \
  } while (0)

#define FRCNN_REQUIRE(cond, msg)                               \
  do {                                                         \
    if (!(cond)) return frcnn::fail(FRCNN_E_BADARG, msg);       \
  } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// grid size for a grid-stride elementwise kernel: a multiple of the SM count
inline int elementwise_grid(size_t work_items, int threads, int max_ctas_per_sm = 8)
{
  size_t want = ceil_div<size_t>(work_items, (size_t)threads);
  size_t cap = (size_t)kNumSMs * max_ctas_per_sm;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace frcnn
