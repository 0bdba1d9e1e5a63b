// Shared helpers for libfrcnn_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <utility>
#include <atomic>
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

// ---- programmatic dependent launch (opt-in: FRCNN_PDL=1 or frcnn_set_pdl(1)) ------------------------------------------------
// Every kernel of this library starts with pdl_enter(): `launch_dependents` lets the NEXT kernel of the stream become resident as
// soon as all CTAs of this one have started (its barrier init / TMEM allocation / index arithmetic then overlap this kernel's tail
// and the launch-to-launch gap disappears), `wait` blocks until the PREVIOUS kernel has completed and its writes are visible.
// Because every kernel waits before its first global access and before it can finish, completion stays transitive along the
// stream (kernel n+1 cannot finish before kernel n), so buffers are never read early or overwritten while still in use.
// Launched without the attribute (the default) both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

extern std::atomic<int> g_pdl;        // api.cu: -1 = read FRCNN_PDL on first use
bool pdl_enabled();

// SMs the persistent (one CTA per SM) GEMM launches may occupy: kNumSMs minus the ones the caller set aside for a concurrent
// collective's CTAs (frcnn_set_sm_reserve; api.cu).  A persistent grid that finds some SMs taken runs a second, nearly empty wave.
extern std::atomic<int> g_sm_reserve;
// SMs of the current device (cudaDevAttrMultiProcessorCount, looked up once per process: one process drives one GPU), capped at the
// kNumSMs the workspace layouts are sized for.  A part with fewer SMs (or an MPS / green-context partition reporting fewer) gets
// persistent grids that still fit in one wave; kNumSMs remains the layout constant (grid_max).
int device_sm_count();
inline int sm_budget()
{
  const int sms = device_sm_count();
  int n = sms - g_sm_reserve.load(std::memory_order_relaxed);
  return n < 16 ? (sms < 16 ? sms : 16) : (n > kNumSMs ? kNumSMs : n);
}

// the one way kernels are launched: <<<>>> semantics, plus the programmatic-serialization attribute when PDL is on
template <typename... Params, typename... Args>
inline void launch(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through FRCNN_CHECK_LAUNCH (cudaGetLastError)
}

// same, as thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... Params, typename... Args>
inline void launch_cluster(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, bool allow_pdl, Args &&...args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.numAttrs = 1;
  if (allow_pdl && pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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
