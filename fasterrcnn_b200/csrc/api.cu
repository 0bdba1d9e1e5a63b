// Engine dispatch for the GEMM-shaped entry points + library-level symbols.
#include "common.cuh"

namespace frcnn {

thread_local char g_last_error[512] = "";

// the two process-wide switches of the library (launch plumbing; see the header)
std::atomic<int> g_pdl{-1};
std::atomic<int> g_sm_reserve{0};

int device_sm_count()
{
  static std::atomic<int> cached{0};
  int v = cached.load(std::memory_order_relaxed);
  if (v <= 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    v = n > kNumSMs ? kNumSMs : n;
    cached.store(v, std::memory_order_relaxed);
  }
  return v;
}

bool pdl_enabled()
{
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char *e = getenv("FRCNN_PDL");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

// conv_simt.cu
size_t simt_fwd_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
size_t simt_dgrad_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
size_t simt_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int simt_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual, float *y,
                    int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                    void *workspace, size_t workspace_bytes, cudaStream_t st);
int simt_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx,
                      int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                      void *workspace, size_t workspace_bytes, cudaStream_t st);
int simt_conv2d_wgrad(const float *dy, const float *x, float *dw,
                      int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                      void *workspace, size_t workspace_bytes, cudaStream_t st);

// conv_tc.cu (tcgen05 engines; mode 0 = fwd, 1 = dgrad, 2 = wgrad; f16 selects the fp16 engine)
bool tc_supported(int mode, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, bool f16);
size_t tc_workspace(int mode, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, bool f16);
int tc_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual, float *y,
                  int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                  void *workspace, size_t workspace_bytes, cudaStream_t st, const void *x_split, const void *w_split, bool f16, void *amax_out = nullptr);
int tc_amax_slots(int mode, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, bool f16);
size_t tf32_split_bytes(size_t count);
size_t f16_split_bytes(size_t count);
void tc_set_trace(void *buf);
int tf32_split(const float *x, size_t count, void *out, cudaStream_t st);
int f16_split(const float *x, size_t count, void *out, cudaStream_t st, const void *partials, int G);
int f16_split_carried(const float *x, size_t count, void *out, int ctas_per_sm, cudaStream_t st);
int tc_pair_max_active_clusters();
int tc_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx,
                    int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                    void *workspace, size_t workspace_bytes, cudaStream_t st, const void *dy_split, const void *w_split, bool f16, void *amax_out = nullptr);
int tc_conv2d_wgrad(const float *dy, const float *x, float *dw,
                    int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                    void *workspace, size_t workspace_bytes, cudaStream_t st, const void *dy_split, const void *x_split, bool f16);

// FRCNN_ENGINE_AUTO = the 3xTF32 tcgen05 engine wherever the shape allows, else the CUDA-core engine; the fp16 engine is chosen explicitly
static bool engine_f16(int engine) { return engine == FRCNN_ENGINE_TC_3XF16; }
// TC_3XTF32 is strict (unsupported shapes are an error: the kernel tests rely on it); TC_3XF16 behaves like AUTO with the fp16 tensor
// engine: shapes it does not take (the RGB stem, narrow heads, strided convs) run on the CUDA-core engine
static bool engine_forced_tc(int engine) { return engine == FRCNN_ENGINE_TC_3XTF32; }

}  // namespace frcnn

using namespace frcnn;

#define GEOM_ARGS N, H, W, Cin, Cout, KH, KW, stride, pad

extern "C" {

int frcnn_version(void) { return 100; }

const char *frcnn_last_error_string(void) { return g_last_error; }

int frcnn_set_sm_reserve(int sms)
{
  return g_sm_reserve.exchange(sms < 0 ? 0 : (sms > kNumSMs - 16 ? kNumSMs - 16 : sms), std::memory_order_relaxed);
}

int frcnn_set_pdl(int enabled)
{
  const int before = pdl_enabled() ? 1 : 0;
  g_pdl.store(enabled ? 1 : 0, std::memory_order_relaxed);
  return before;
}

size_t frcnn_conv2d_fwd_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  size_t a = simt_fwd_workspace(GEOM_ARGS);
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_supported(0, GEOM_ARGS, engine_f16(engine))) {
    size_t b = tc_workspace(0, GEOM_ARGS, engine_f16(engine));
    return engine_forced_tc(engine) ? b : (a > b ? a : b);
  }
  return a;
}

int frcnn_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual,
                     float *y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                     int engine, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(x && w && y, "conv2d_fwd: null pointer");
  FRCNN_REQUIRE(act >= FRCNN_ACT_NONE && act <= FRCNN_ACT_SIGMOID, "conv2d_fwd: unknown activation");
  const bool f16 = engine_f16(engine);
  if (engine_forced_tc(engine) && !tc_supported(0, GEOM_ARGS, f16)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_fwd: shape not supported by the tcgen05 engine");
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_supported(0, GEOM_ARGS, f16))
    return tc_conv2d_fwd(x, w, scale, bias, residual, y, GEOM_ARGS, act, workspace, workspace_bytes, as_stream(stream), nullptr, nullptr, f16);
  return simt_conv2d_fwd(x, w, scale, bias, residual, y, GEOM_ARGS, act, workspace, workspace_bytes, as_stream(stream));
}

size_t frcnn_conv2d_dgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  size_t a = simt_dgrad_workspace(GEOM_ARGS);
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_supported(1, GEOM_ARGS, engine_f16(engine))) {
    size_t b = tc_workspace(1, GEOM_ARGS, engine_f16(engine));
    return engine_forced_tc(engine) ? b : (a > b ? a : b);
  }
  return a;
}

int frcnn_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx,
                       int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                       int engine, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && w && dx, "conv2d_dgrad: null pointer");
  const bool f16 = engine_f16(engine);
  if (engine_forced_tc(engine) && !tc_supported(1, GEOM_ARGS, f16)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_dgrad: shape not supported by the tcgen05 engine");
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_supported(1, GEOM_ARGS, f16))
    return tc_conv2d_dgrad(dy, w, addend, dx, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), nullptr, nullptr, f16);
  return simt_conv2d_dgrad(dy, w, addend, dx, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream));
}

size_t frcnn_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  size_t a = simt_wgrad_workspace(GEOM_ARGS);
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_supported(2, GEOM_ARGS, engine_f16(engine))) {
    size_t b = tc_workspace(2, GEOM_ARGS, engine_f16(engine));
    return engine_forced_tc(engine) ? b : (a > b ? a : b);
  }
  return a;
}

int frcnn_conv2d_wgrad(const float *dy, const float *x, float *dw,
                       int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                       int engine, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && x && dw, "conv2d_wgrad: null pointer");
  const bool f16 = engine_f16(engine);
  if (engine_forced_tc(engine) && !tc_supported(2, GEOM_ARGS, f16)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_wgrad: shape not supported by the tcgen05 engine");
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_supported(2, GEOM_ARGS, f16))
    return tc_conv2d_wgrad(dy, x, dw, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), nullptr, nullptr, f16);
  return simt_conv2d_wgrad(dy, x, dw, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream));
}

/* ---- tf32 hi/lo operand splits shared between the passes of one step -------------------------------- */
int frcnn_conv2d_uses_tensor_cores(int pass, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  if (engine == FRCNN_ENGINE_SIMT_FP32 || pass < 0 || pass > 2) return 0;
  return tc_supported(pass, GEOM_ARGS, engine_f16(engine)) ? 1 : 0;
}

size_t frcnn_tf32_split_bytes(size_t count) { return tf32_split_bytes(count); }

void frcnn_debug_tc_trace(void *buf) { tc_set_trace(buf); }
int frcnn_debug_pair_max_active_clusters(void) { return tc_pair_max_active_clusters(); }

int frcnn_tf32_split(const float *x, size_t count, void *out, void *stream)
{
  FRCNN_REQUIRE(x && out && count > 0, "tf32_split: bad argument");
  return tf32_split(x, count, out, as_stream(stream));
}

int frcnn_conv2d_fwd_presplit(const float *x, const float *w, const void *x_split, const void *w_split, const float *scale, const float *bias,
                              const float *residual, float *y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                              void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(x && w && y, "conv2d_fwd_presplit: null pointer");
  if (!tc_supported(0, GEOM_ARGS, false)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_fwd_presplit: shape not supported by the tcgen05 engine");
  return tc_conv2d_fwd(x, w, scale, bias, residual, y, GEOM_ARGS, act, workspace, workspace_bytes, as_stream(stream), x_split, w_split, false);
}

int frcnn_conv2d_dgrad_presplit(const float *dy, const float *w, const void *dy_split, const void *w_split, const float *addend, float *dx,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                                void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && w && dx, "conv2d_dgrad_presplit: null pointer");
  if (!tc_supported(1, GEOM_ARGS, false)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_dgrad_presplit: shape not supported by the tcgen05 engine");
  return tc_conv2d_dgrad(dy, w, addend, dx, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), dy_split, w_split, false);
}

int frcnn_conv2d_wgrad_presplit(const float *dy, const float *x, const void *dy_split, const void *x_split, float *dw,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                                void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && x && dw, "conv2d_wgrad_presplit: null pointer");
  if (!tc_supported(2, GEOM_ARGS, false)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_wgrad_presplit: shape not supported by the tcgen05 engine");
  return tc_conv2d_wgrad(dy, x, dw, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), dy_split, x_split, false);
}

/* ---- fp16 engine (FRCNN_ENGINE_TC_3XF16): operand splits and the GEMM entry points that take them ---------------- */
size_t frcnn_f16_split_bytes(size_t count) { return f16_split_bytes(count); }

int frcnn_f16_split(const float *x, size_t count, void *out, void *stream)
{
  FRCNN_REQUIRE(x && out && count > 0, "f16_split: bad argument");
  FRCNN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 127) == 0, "f16_split: x must be 16-byte, out 128-byte aligned");
  return f16_split(x, count, out, as_stream(stream), nullptr, 0);
}

int frcnn_f16_split_carried(const float *x, size_t count, void *out, int ctas_per_sm, void *stream)
{
  FRCNN_REQUIRE(x && out && count > 0, "f16_split_carried: bad argument");
  FRCNN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 127) == 0, "f16_split_carried: x must be 16-byte, out 128-byte aligned");
  return f16_split_carried(x, count, out, ctas_per_sm, as_stream(stream));
}

int frcnn_f16_split_from_amax(const float *x, size_t count, const void *amax, int slots, void *out, void *stream)
{
  FRCNN_REQUIRE(x && out && amax && count > 0 && slots > 0 && slots <= 1000, "f16_split_from_amax: bad argument");
  FRCNN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 127) == 0, "f16_split_from_amax: x must be 16-byte, out 128-byte aligned");
  return f16_split(x, count, out, as_stream(stream), amax, slots);
}

int frcnn_conv2d_amax_slots(int pass, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad)
{
  return (pass == 0 || pass == 1) ? tc_amax_slots(pass, GEOM_ARGS, true) : 0;
}

int frcnn_conv2d_fwd_f16(const float *x, const float *w, const void *x_split, const void *w_split, const float *scale, const float *bias,
                         const float *residual, float *y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                         void *y_amax, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(x && w && y, "conv2d_fwd_f16: null pointer");
  if (!tc_supported(0, GEOM_ARGS, true)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_fwd_f16: shape not supported by the fp16 tcgen05 engine");
  return tc_conv2d_fwd(x, w, scale, bias, residual, y, GEOM_ARGS, act, workspace, workspace_bytes, as_stream(stream), x_split, w_split, true, y_amax);
}

int frcnn_conv2d_dgrad_f16(const float *dy, const float *w, const void *dy_split, const void *w_split, const float *addend, float *dx,
                           int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           void *dx_amax, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && w && dx, "conv2d_dgrad_f16: null pointer");
  if (!tc_supported(1, GEOM_ARGS, true)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_dgrad_f16: shape not supported by the fp16 tcgen05 engine");
  return tc_conv2d_dgrad(dy, w, addend, dx, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), dy_split, w_split, true, dx_amax);
}

int frcnn_conv2d_wgrad_f16(const float *dy, const float *x, const void *dy_split, const void *x_split, float *dw,
                           int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && x && dw, "conv2d_wgrad_f16: null pointer");
  if (!tc_supported(2, GEOM_ARGS, true)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_wgrad_f16: shape not supported by the fp16 tcgen05 engine");
  return tc_conv2d_wgrad(dy, x, dw, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), dy_split, x_split, true);
}

/* The whole backward of `y = [maxpool2x2] act(conv(x, w) + b)` on the fp16 engine in one call: (pooling + ReLU backward) -> fused
 * activation-backward / operand split / bias row-sum -> data gradient -> filter gradient.  Same kernels as the single entry points, launched
 * back to back from C: what it saves is four host round trips per layer. */
int frcnn_conv2d_bwd_f16(const float *dy, const float *y, int act, int pooled, const float *x, const void *x_split, const float *w, const void *w_split,
                         const void *dy_amax, int dy_amax_slots, float *dz_full, void *dz_split, float *dbias, float *dx, void *dx_amax, float *dw,
                         int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                         void *bias_workspace, size_t bias_workspace_bytes, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && x && w && dz_split && (dx || dw), "conv2d_bwd_f16: null pointer");
  FRCNN_REQUIRE(stride == 1, "conv2d_bwd_f16: stride-1 layers only");
  const int Ho = H + 2 * pad - KH + 1, Wo = W + 2 * pad - KW + 1;
  const size_t rows = (size_t)N * Ho * Wo;
  const float *g = dy;
  if (pooled) {                                          // dy is the gradient of the pooled map: route it to the window maxima under the ReLU mask
    FRCNN_REQUIRE(y && dz_full && act == FRCNN_ACT_RELU, "conv2d_bwd_f16: pooled layers need y, a full-resolution scratch map and a ReLU");
    int rc = frcnn_maxpool2x2_relu_bwd(dy, y, dz_full, N, Ho, Wo, Cout, stream);
    if (rc != FRCNN_OK) return rc;
    g = dz_full;
    act = FRCNN_ACT_NONE;
  }
  int rc = frcnn_act_bwd_fused_f16(g, y, act, nullptr, dz_split, dbias, rows, Cout, dy_amax, dy_amax_slots, bias_workspace, bias_workspace_bytes, stream);
  if (rc != FRCNN_OK) return rc;
  if (dx) {
    rc = frcnn_conv2d_dgrad_f16(g, w, dz_split, w_split, nullptr, dx, GEOM_ARGS, dx_amax, workspace, workspace_bytes, stream);
    if (rc != FRCNN_OK) return rc;
  }
  if (dw) rc = frcnn_conv2d_wgrad_f16(g, x, dz_split, x_split, dw, GEOM_ARGS, workspace, workspace_bytes, stream);
  return rc;
}

}  // extern "C"
