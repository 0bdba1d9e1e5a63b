// Engine dispatch for the GEMM-shaped entry points + library-level symbols.
#include "common.cuh"

namespace frcnn {

thread_local char g_last_error[512] = "";

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

// conv_tc.cu (tcgen05 engine)
bool tc_fwd_supported(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
size_t tc_fwd_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int tc_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual, float *y,
                  int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                  void *workspace, size_t workspace_bytes, cudaStream_t st, const void *x_split, const void *w_split);
size_t tf32_split_bytes(size_t count);
void tc_set_trace(void *buf);
int tf32_split(const float *x, size_t count, void *out, cudaStream_t st);
bool tc_dgrad_supported(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
size_t tc_dgrad_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int tc_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx,
                    int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                    void *workspace, size_t workspace_bytes, cudaStream_t st, const void *dy_split, const void *w_split);
bool tc_wgrad_supported(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
size_t tc_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int tc_conv2d_wgrad(const float *dy, const float *x, float *dw,
                    int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                    void *workspace, size_t workspace_bytes, cudaStream_t st, const void *dy_split, const void *x_split);

}  // namespace frcnn

using namespace frcnn;

#define GEOM_ARGS N, H, W, Cin, Cout, KH, KW, stride, pad

extern "C" {

int frcnn_version(void) { return 100; }

const char *frcnn_last_error_string(void) { return g_last_error; }

size_t frcnn_conv2d_fwd_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  size_t a = simt_fwd_workspace(GEOM_ARGS);
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_fwd_supported(GEOM_ARGS)) {
    size_t b = tc_fwd_workspace(GEOM_ARGS);
    return engine == FRCNN_ENGINE_TC_3XTF32 ? b : (a > b ? a : b);
  }
  return a;
}

int frcnn_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual,
                     float *y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                     int engine, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(x && w && y, "conv2d_fwd: null pointer");
  FRCNN_REQUIRE(act >= FRCNN_ACT_NONE && act <= FRCNN_ACT_SIGMOID, "conv2d_fwd: unknown activation");
  if (engine == FRCNN_ENGINE_TC_3XTF32 && !tc_fwd_supported(GEOM_ARGS)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_fwd: shape not supported by the tcgen05 engine");
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_fwd_supported(GEOM_ARGS))
    return tc_conv2d_fwd(x, w, scale, bias, residual, y, GEOM_ARGS, act, workspace, workspace_bytes, as_stream(stream), nullptr, nullptr);
  return simt_conv2d_fwd(x, w, scale, bias, residual, y, GEOM_ARGS, act, workspace, workspace_bytes, as_stream(stream));
}

size_t frcnn_conv2d_dgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  size_t a = simt_dgrad_workspace(GEOM_ARGS);
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_dgrad_supported(GEOM_ARGS)) {
    size_t b = tc_dgrad_workspace(GEOM_ARGS);
    return engine == FRCNN_ENGINE_TC_3XTF32 ? b : (a > b ? a : b);
  }
  return a;
}

int frcnn_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx,
                       int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                       int engine, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && w && dx, "conv2d_dgrad: null pointer");
  if (engine == FRCNN_ENGINE_TC_3XTF32 && !tc_dgrad_supported(GEOM_ARGS)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_dgrad: shape not supported by the tcgen05 engine");
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_dgrad_supported(GEOM_ARGS))
    return tc_conv2d_dgrad(dy, w, addend, dx, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), nullptr, nullptr);
  return simt_conv2d_dgrad(dy, w, addend, dx, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream));
}

size_t frcnn_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  size_t a = simt_wgrad_workspace(GEOM_ARGS);
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_wgrad_supported(GEOM_ARGS)) {
    size_t b = tc_wgrad_workspace(GEOM_ARGS);
    return engine == FRCNN_ENGINE_TC_3XTF32 ? b : (a > b ? a : b);
  }
  return a;
}

int frcnn_conv2d_wgrad(const float *dy, const float *x, float *dw,
                       int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                       int engine, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && x && dw, "conv2d_wgrad: null pointer");
  if (engine == FRCNN_ENGINE_TC_3XTF32 && !tc_wgrad_supported(GEOM_ARGS)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_wgrad: shape not supported by the tcgen05 engine");
  if (engine != FRCNN_ENGINE_SIMT_FP32 && tc_wgrad_supported(GEOM_ARGS))
    return tc_conv2d_wgrad(dy, x, dw, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), nullptr, nullptr);
  return simt_conv2d_wgrad(dy, x, dw, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream));
}

/* ---- tf32 hi/lo operand splits shared between the passes of one step -------------------------------- */
int frcnn_conv2d_uses_tensor_cores(int pass, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine)
{
  if (engine == FRCNN_ENGINE_SIMT_FP32) return 0;
  if (pass == 0) return tc_fwd_supported(GEOM_ARGS) ? 1 : 0;
  if (pass == 1) return tc_dgrad_supported(GEOM_ARGS) ? 1 : 0;
  return tc_wgrad_supported(GEOM_ARGS) ? 1 : 0;
}

size_t frcnn_tf32_split_bytes(size_t count) { return tf32_split_bytes(count); }

void frcnn_debug_tc_trace(void *buf) { tc_set_trace(buf); }

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
  if (!tc_fwd_supported(GEOM_ARGS)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_fwd_presplit: shape not supported by the tcgen05 engine");
  return tc_conv2d_fwd(x, w, scale, bias, residual, y, GEOM_ARGS, act, workspace, workspace_bytes, as_stream(stream), x_split, w_split);
}

int frcnn_conv2d_dgrad_presplit(const float *dy, const float *w, const void *dy_split, const void *w_split, const float *addend, float *dx,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                                void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && w && dx, "conv2d_dgrad_presplit: null pointer");
  if (!tc_dgrad_supported(GEOM_ARGS)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_dgrad_presplit: shape not supported by the tcgen05 engine");
  return tc_conv2d_dgrad(dy, w, addend, dx, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), dy_split, w_split);
}

int frcnn_conv2d_wgrad_presplit(const float *dy, const float *x, const void *dy_split, const void *x_split, float *dw,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                                void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy && x && dw, "conv2d_wgrad_presplit: null pointer");
  if (!tc_wgrad_supported(GEOM_ARGS)) return fail(FRCNN_E_UNSUPPORTED, "conv2d_wgrad_presplit: shape not supported by the tcgen05 engine");
  return tc_conv2d_wgrad(dy, x, dw, GEOM_ARGS, workspace, workspace_bytes, as_stream(stream), dy_split, x_split);
}

}  // extern "C"
