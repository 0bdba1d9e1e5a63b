// K7: RoI max pooling forward/backward (torchvision.ops.RoIPool semantics).
//
// Forward: one CTA per (RoI, 128-channel slab); the feature map is NHWC so the 32 lanes of a warp
// read 32 consecutive channels of one cell (128 B, fully coalesced).  Each thread scans its
// channel's 49 bins in the reference's row-major order with a strict '>' so that the argmax is
// the reference's; results are staged in shared memory as [channel][49] and written out as one
// contiguous (128 x 49) fp32 run per slab -- i.e. directly in the (K, C, 7, 7) order that fc1
// (models/vgg16.py:129) consumes, with 128-bit stores.
//
// Backward: deterministic, atomics-free.  A thread owns one (feature row h, channel c) line of
// the gradient map in shared memory, walks every (RoI, bin) of its channel in ascending order and
// accumulates the entries whose argmax falls on its line; the summation order per cell is
// therefore fixed (RoI ascending, bin ascending), unlike the atomicAdd scatter of the library op.
#include <float.h>
#include "common.cuh"

namespace frcnn {

constexpr int kSlab = 128;       // channels per CTA (forward)

struct RoiBins {
  int ys, xs, rh, rw;
  float bh, bw;
};

// proposals are (y1,x1,y2,x2); RoIPool sees (x1,y1,x2,y2) (models/detector.py:68-69)
__device__ __forceinline__ RoiBins roi_bins(const float *__restrict__ p, float scale, int PH, int PW)
{
  RoiBins r;
  r.xs = (int)roundf(__fmul_rn(p[1], scale));
  r.ys = (int)roundf(__fmul_rn(p[0], scale));
  int xe = (int)roundf(__fmul_rn(p[3], scale));
  int ye = (int)roundf(__fmul_rn(p[2], scale));
  r.rw = max(xe - r.xs + 1, 1);
  r.rh = max(ye - r.ys + 1, 1);
  r.bh = __fdiv_rn((float)r.rh, (float)PH);
  r.bw = __fdiv_rn((float)r.rw, (float)PW);
  return r;
}

__global__ void __launch_bounds__(kSlab)
roi_pool_fwd_kernel(const float *__restrict__ fm, int H, int W, int C, const float *__restrict__ proposals, int PH, int PW, float scale,
                    float *__restrict__ out, int32_t *__restrict__ argmax)
{
  extern __shared__ float smem[];                 // [kSlab][PH*PW] values then [kSlab][PH*PW] argmax
  const int bins = PH * PW;
  float *s_val = smem;
  int32_t *s_arg = reinterpret_cast<int32_t *>(smem + (size_t)kSlab * bins);
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kSlab;
  const int c = c0 + threadIdx.x;
  const RoiBins r = roi_bins(proposals + 4 * (size_t)n, scale, PH, PW);
  if (c < C) {
    for (int ph = 0; ph < PH; ph++) {
      int hs = (int)floorf(__fmul_rn((float)ph, r.bh)) + r.ys;
      int he = (int)ceilf(__fmul_rn((float)(ph + 1), r.bh)) + r.ys;
      hs = min(max(hs, 0), H); he = min(max(he, 0), H);
      for (int pw = 0; pw < PW; pw++) {
        int ws = (int)floorf(__fmul_rn((float)pw, r.bw)) + r.xs;
        int we = (int)ceilf(__fmul_rn((float)(pw + 1), r.bw)) + r.xs;
        ws = min(max(ws, 0), W); we = min(max(we, 0), W);
        bool empty = (he <= hs) || (we <= ws);
        float best = empty ? 0.f : -FLT_MAX;
        int besti = -1;
        for (int h = hs; h < he; h++) {
          const float *row = fm + ((size_t)h * W) * C + c;
          for (int w = ws; w < we; w++) {
            float v = __ldg(row + (size_t)w * C);
            if (v > best) { best = v; besti = h * W + w; }
          }
        }
        s_val[threadIdx.x * bins + ph * PW + pw] = best;
        s_arg[threadIdx.x * bins + ph * PW + pw] = besti;
      }
    }
  }
  __syncthreads();
  // contiguous write-out of the slab: out[(n*C + c0) * bins ...]
  int live = min(kSlab, C - c0);
  size_t base = ((size_t)n * C + c0) * bins;
  int total = live * bins;
  if (((base | (size_t)total) & 3) == 0) {
    float4 *o4 = reinterpret_cast<float4 *>(out + base);
    int4 *a4 = reinterpret_cast<int4 *>(argmax + base);
    const float4 *sv4 = reinterpret_cast<const float4 *>(s_val);
    const int4 *sa4 = reinterpret_cast<const int4 *>(s_arg);
    for (int e = threadIdx.x; e < total / 4; e += blockDim.x) {
      o4[e] = sv4[e];
      if (argmax) a4[e] = sa4[e];
    }
  } else {
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      out[base + e] = s_val[e];
      if (argmax) argmax[base + e] = s_arg[e];
    }
  }
}

// grid (C/32, ceil(H/8)); block (32 lanes = channels, 8 warps = rows).  A per-CTA table of the
// clipped row range [hs,he) of every (RoI, ph) lets a thread visit only the bins that can hold an
// argmax on its row (typically 7-14 of the 49), in ascending (RoI, bin) order.
__global__ void __launch_bounds__(256)
roi_pool_bwd_kernel(const float *__restrict__ dout, const int32_t *__restrict__ argmax, const float *__restrict__ proposals, float scale,
                    int K, int H, int W, int C, int PH, int PW, const float *__restrict__ addend, float *__restrict__ dfm)
{
  extern __shared__ float line[];                 // [W][256] floats, then the (K x PH) row-range table
  int32_t *rows = reinterpret_cast<int32_t *>(line + (size_t)W * 256);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int h = blockIdx.y * 8 + warp;
  const bool live = c < C && h < H;
  const int bins = PH * PW;
  for (int e = threadIdx.x; e < K * PH; e += blockDim.x) {
    int n = e / PH, ph = e - n * PH;
    const RoiBins r = roi_bins(proposals + 4 * (size_t)n, scale, PH, PW);
    int hs = (int)floorf(__fmul_rn((float)ph, r.bh)) + r.ys;
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), r.bh)) + r.ys;
    hs = min(max(hs, 0), H); he = min(max(he, 0), H);
    rows[e] = (hs << 16) | he;
  }
  for (int w = 0; w < W; w++) line[w * 256 + threadIdx.x] = 0.f;
  __syncthreads();
  if (live) {
    const int lo = h * W, hi = lo + W;
    for (int n = 0; n < K; n++) {
      const int32_t *a = argmax + ((size_t)n * C + c) * bins;
      const float *g = dout + ((size_t)n * C + c) * bins;
      for (int ph = 0; ph < PH; ph++) {
        const int packed = rows[n * PH + ph];
        if (h < (packed >> 16) || h >= (packed & 0xffff)) continue;
        for (int pw = 0; pw < PW; pw++) {
          int idx = __ldg(a + ph * PW + pw);
          if (idx >= lo && idx < hi) line[(idx - lo) * 256 + threadIdx.x] += __ldg(g + ph * PW + pw);
        }
      }
    }
    for (int w = 0; w < W; w++) {
      size_t o = ((size_t)h * W + w) * C + c;
      float v = line[w * 256 + threadIdx.x];
      if (addend) v += __ldg(addend + o);
      dfm[o] = v;
    }
  }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_roi_pool_fwd(const float *fm, int H, int W, int C, const float *proposals, int K, int PH, int PW, float spatial_scale,
                       float *out, int32_t *argmax, void *stream)
{
  FRCNN_REQUIRE(fm && proposals && out && H > 0 && W > 0 && C > 0 && K >= 0 && PH > 0 && PW > 0, "roi_pool_fwd: bad argument");
  if (K == 0) return FRCNN_OK;
  size_t smem = (size_t)kSlab * PH * PW * (sizeof(float) + sizeof(int32_t));
  FRCNN_REQUIRE(smem <= 200 * 1024, "roi_pool_fwd: pooled size too large");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(roi_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "roi_pool_fwd: smem attribute");
  }
  roi_pool_fwd_kernel<<<dim3(K, ceil_div(C, kSlab)), kSlab, smem, as_stream(stream)>>>(fm, H, W, C, proposals, PH, PW, spatial_scale, out, argmax);
  FRCNN_CHECK_LAUNCH("roi_pool_fwd_kernel");
  return FRCNN_OK;
}

int frcnn_roi_pool_bwd(const float *dout, const int32_t *argmax, const float *proposals, int K, int H, int W, int C, int PH, int PW,
                       float spatial_scale, const float *addend, float *dfm, void *stream)
{
  FRCNN_REQUIRE(dout && argmax && proposals && dfm && K >= 0 && H > 0 && W > 0 && C > 0 && PH > 0 && PW > 0, "roi_pool_bwd: bad argument");
  FRCNN_REQUIRE(H < 32768, "roi_pool_bwd: feature map too tall");
  size_t smem = (size_t)W * 256 * sizeof(float) + (size_t)K * PH * sizeof(int32_t);
  FRCNN_REQUIRE(smem <= 200 * 1024, "roi_pool_bwd: feature map too wide / too many RoIs for the shared-memory buffers");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(roi_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "roi_pool_bwd: smem attribute");
  }
  roi_pool_bwd_kernel<<<dim3(ceil_div(C, 32), ceil_div(H, 8)), 256, smem, as_stream(stream)>>>(dout, argmax, proposals, spatial_scale, K, H, W, C, PH, PW, addend, dfm);
  FRCNN_CHECK_LAUNCH("roi_pool_bwd_kernel");
  return FRCNN_OK;
}

}  // extern "C"
