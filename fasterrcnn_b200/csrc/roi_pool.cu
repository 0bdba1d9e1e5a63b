// K7: RoI max pooling forward/backward (torchvision.ops.RoIPool semantics).
//
// Forward (roi_pool_fwd_v4_kernel, C % 4 == 0): one CTA per (RoI, 128-channel slab), 8 warps; a warp takes every 8th bin and
// its 32 lanes read 4 consecutive channels each -- one 512-byte NHWC segment per warp-wide 128-bit load, so a bin of r x s cells
// is r*s independent vector loads per lane (the feature map is L2 resident; what has to be hidden is L2 latency, i.e. bytes in
// flight per thread).  Each bin is scanned in the reference's row-major order with a strict '>' per channel so that the argmax
// is the reference's; the values are staged in shared memory as [channel][49] and streamed out (st.global.cs, the output must not
// evict the feature map from L2) as one contiguous (128 x 49) fp32 run -- i.e. directly in the (K, C, 7, 7) order that fc1
// (models/vgg16.py:129) consumes -- with 128-bit stores.  The argmax is private to this file (forward writes it, backward reads
// it), so it is kept bin-major, (K, 49, C): the forward stores it straight from registers (a warp's four channels per lane are
// one coalesced 512-byte run per bin, no staging), the backward reads it with lane = channel, coalesced.  The HBM write of
// values + argmax is the algorithmic traffic.
// roi_pool_fwd_kernel (one channel per lane, 32-channel groups) remains for channel counts that are not a multiple of 4.
//
// Backward: deterministic, atomics-free.  A CTA owns one feature-map row h and 32 channels; its 8 warps split the RoIs into 8
// contiguous chunks, each warp accumulating into its own [W][32] line buffer in shared memory the gradient entries of its RoIs
// whose argmax falls on row h (a per-CTA table of clipped row ranges skips the bins that cannot), walking (RoI, bin) in
// ascending order; the 8 buffers are then combined in chunk order.  The summation order per cell is therefore fixed, unlike
// the atomicAdd scatter of the library op.
#include <float.h>
#include "common.cuh"

namespace frcnn {

constexpr int kRoiChannels = 32;  // channels per CTA (one warp-wide NHWC segment)
constexpr int kRoiWarps = 8;

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

__global__ void __launch_bounds__(kRoiWarps * 32)
roi_pool_fwd_kernel(const float *__restrict__ fm, int H, int W, int C, const float *__restrict__ proposals, int PH, int PW, float scale,
                    float *__restrict__ out, int32_t *__restrict__ argmax)
{
  pdl_enter();
  extern __shared__ float smem[];                 // [32][PH*PW] values
  const int bins = PH * PW;
  float *s_val = smem;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kRoiChannels;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane;
  const RoiBins r = roi_bins(proposals + 4 * (size_t)n, scale, PH, PW);
  if (c < C) {
    for (int b = warp; b < bins; b += kRoiWarps) {
      const int ph = b / PW, pw = b - ph * PW;
      int hs = (int)floorf(__fmul_rn((float)ph, r.bh)) + r.ys;
      int he = (int)ceilf(__fmul_rn((float)(ph + 1), r.bh)) + r.ys;
      hs = min(max(hs, 0), H); he = min(max(he, 0), H);
      int ws = (int)floorf(__fmul_rn((float)pw, r.bw)) + r.xs;
      int we = (int)ceilf(__fmul_rn((float)(pw + 1), r.bw)) + r.xs;
      ws = min(max(ws, 0), W); we = min(max(we, 0), W);
      const bool empty = (he <= hs) || (we <= ws);
      float best = empty ? 0.f : -FLT_MAX;
      int besti = -1;
      for (int h = hs; h < he; h++) {
        const float *row = fm + ((size_t)h * W) * C + c;
        for (int w = ws; w < we; w++) {
          const float v = __ldg(row + (size_t)w * C);
          if (v > best) { best = v; besti = h * W + w; }
        }
      }
      s_val[lane * bins + b] = best;
      if (argmax) argmax[((size_t)n * bins + b) * C + c] = besti;      // bin-major argmax: coalesced over the warp's channels
    }
  }
  __syncthreads();
  // contiguous write-out of the group: out[(n*C + c0) * bins ...]
  const int live = min(kRoiChannels, C - c0);
  const size_t base = ((size_t)n * C + c0) * bins;
  const int total = live * bins;
  if (((base | (size_t)total) & 3) == 0) {
    float4 *o4 = reinterpret_cast<float4 *>(out + base);
    const float4 *sv4 = reinterpret_cast<const float4 *>(s_val);
    for (int e = threadIdx.x; e < total / 4; e += blockDim.x) o4[e] = sv4[e];
  } else {
    for (int e = threadIdx.x; e < total; e += blockDim.x) out[base + e] = s_val[e];
  }
}


// volatile: keeps the four loads of a round back to back (ptxas otherwise sinks each one below the previous load's compares to
// save registers, which serialises the L2 round trips)
__device__ __forceinline__ float4 ldg_nc_v4(const float4 *p)
{
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ---- forward, 4 channels per lane ---------------------------------------------------------------------------------------
constexpr int kRoiSlab = 128;     // channels per CTA: 32 lanes x float4

__global__ void __launch_bounds__(kRoiWarps * 32)
roi_pool_fwd_v4_kernel(const float *__restrict__ fm, int H, int W, int C, const float *__restrict__ proposals, int PH, int PW, float scale,
                       float *__restrict__ out, int32_t *__restrict__ argmax)
{
  pdl_enter();
  extern __shared__ float smem[];                 // [128][PH*PW] values
  const int bins = PH * PW;
  float *s_val = smem;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * kRoiSlab;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + 4 * lane;
  const RoiBins r = roi_bins(proposals + 4 * (size_t)n, scale, PH, PW);
  if (c < C) {
    const float4 *fm4 = reinterpret_cast<const float4 *>(fm + c);
    const int C4 = C >> 2;
    for (int b = warp; b < bins; b += kRoiWarps) {
      const int ph = b / PW, pw = b - ph * PW;
      int hs = (int)floorf(__fmul_rn((float)ph, r.bh)) + r.ys;
      int he = (int)ceilf(__fmul_rn((float)(ph + 1), r.bh)) + r.ys;
      hs = min(max(hs, 0), H); he = min(max(he, 0), H);
      int ws = (int)floorf(__fmul_rn((float)pw, r.bw)) + r.xs;
      int we = (int)ceilf(__fmul_rn((float)(pw + 1), r.bw)) + r.xs;
      ws = min(max(ws, 0), W); we = min(max(we, 0), W);
      const bool empty = (he <= hs) || (we <= ws);
      const float init = empty ? 0.f : -FLT_MAX;
      float b0 = init, b1 = init, b2 = init, b3 = init;
      int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
      // the bin's cells in row-major order, four per round, software-pipelined by one round: the 128-bit loads of round r + 1 are
      // issued before the compares of round r, so at least four L2 round trips per lane overlap whatever ptxas interleaves inside
      // a round (a bin is typically 2-3 cells wide: too short for the w loop alone to keep loads in flight).  Slots past the end
      // re-read the last cell -- a second look at a cell never passes the strict '>' -- so the loads carry no predicate.
      const int nw = we - ws, total = empty ? 0 : (he - hs) * nw;
      if (total > 0) {
        int h = hs, w = ws, issued = 0;
        int cell[4], ncell[4];
        float4 v[4], nv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          cell[u] = h * W + w;
          if (++issued < total && ++w == we) { w = ws; h++; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = ldg_nc_v4(fm4 + (size_t)cell[u] * C4);
        for (int k = 0; k < total; k += 4) {
          const bool more = k + 4 < total;                           // warp-uniform
          if (more) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
              ncell[u] = h * W + w;
              if (++issued < total && ++w == we) { w = ws; h++; }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) nv[u] = ldg_nc_v4(fm4 + (size_t)ncell[u] * C4);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (v[u].x > b0) { b0 = v[u].x; i0 = cell[u]; }
            if (v[u].y > b1) { b1 = v[u].y; i1 = cell[u]; }
            if (v[u].z > b2) { b2 = v[u].z; i2 = cell[u]; }
            if (v[u].w > b3) { b3 = v[u].w; i3 = cell[u]; }
          }
          if (more) {
#pragma unroll
            for (int u = 0; u < 4; u++) { v[u] = nv[u]; cell[u] = ncell[u]; }
          }
        }
      }
      const int o = (4 * lane) * bins + b;
      s_val[o] = b0; s_val[o + bins] = b1; s_val[o + 2 * bins] = b2; s_val[o + 3 * bins] = b3;
      if (argmax) __stcs(reinterpret_cast<int4 *>(argmax + ((size_t)n * bins + b) * C + c), make_int4(i0, i1, i2, i3));   // bin-major, 512 B per warp
    }
  }
  __syncthreads();
  // contiguous write-out of the slab: out[(n*C + c0) * bins ...]; live * bins and the base are multiples of 4 (C % 4 == 0)
  const int live = min(kRoiSlab, C - c0);
  const size_t base = ((size_t)n * C + c0) * bins;
  const int total4 = (live * bins) >> 2;
  float4 *o4 = reinterpret_cast<float4 *>(out + base);
  const float4 *sv4 = reinterpret_cast<const float4 *>(s_val);
  for (int e = threadIdx.x; e < total4; e += blockDim.x) __stcs(o4 + e, sv4[e]);
}

// grid (ceil(C/32), H); block = 8 warps x 32 channel lanes.
__global__ void __launch_bounds__(kRoiWarps * 32)
roi_pool_bwd_kernel(const float *__restrict__ dout, const int32_t *__restrict__ argmax, const float *__restrict__ proposals, float scale,
                    int K, int H, int W, int C, int PH, int PW, const float *__restrict__ addend, float *__restrict__ dfm)
{
  pdl_enter();
  extern __shared__ float line[];                 // [8 chunks][W][32] floats, then the (K x PH) row-range table
  int32_t *rows = reinterpret_cast<int32_t *>(line + (size_t)kRoiWarps * W * 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32;
  const int c = c0 + lane;
  const int h = blockIdx.y;
  const int bins = PH * PW;
  for (int e = threadIdx.x; e < K * PH; e += blockDim.x) {
    int n = e / PH, ph = e - n * PH;
    const RoiBins r = roi_bins(proposals + 4 * (size_t)n, scale, PH, PW);
    int hs = (int)floorf(__fmul_rn((float)ph, r.bh)) + r.ys;
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), r.bh)) + r.ys;
    hs = min(max(hs, 0), H); he = min(max(he, 0), H);
    rows[e] = (hs << 16) | he;
  }
  float *mine = line + (size_t)warp * W * 32;
  for (int w = 0; w < W; w++) mine[w * 32 + lane] = 0.f;
  __syncthreads();
  if (c < C) {
    const int lo = h * W, hi = lo + W;
    const int per = (K + kRoiWarps - 1) / kRoiWarps;
    const int n_end = min(K, (warp + 1) * per);
    for (int n = warp * per; n < n_end; n++) {
      const int32_t *a = argmax + (size_t)n * bins * C + c;              // bin-major: a[bin * C], lanes = consecutive channels
      const float *g = dout + ((size_t)n * C + c) * bins;
      for (int ph = 0; ph < PH; ph++) {
        const int packed = rows[n * PH + ph];
        if (h < (packed >> 16) || h >= (packed & 0xffff)) continue;       // warp-uniform: the table is per (RoI, bin row)
        if (PW == 7) {
          // the bin row's 7 argmax words and 7 gradients are loaded before the first test (14 loads in flight per lane): with one
          // load per iteration the walk was a chain of L2 round trips (~120 us for 128 RoIs)
          int idx[7];
          float gv[7];
#pragma unroll
          for (int pw = 0; pw < 7; pw++) idx[pw] = __ldg(a + (size_t)(ph * 7 + pw) * C);
#pragma unroll
          for (int pw = 0; pw < 7; pw++) gv[pw] = __ldg(g + ph * 7 + pw);
#pragma unroll
          for (int pw = 0; pw < 7; pw++)
            if (idx[pw] >= lo && idx[pw] < hi) mine[(idx[pw] - lo) * 32 + lane] += gv[pw];
        } else {
          for (int pw = 0; pw < PW; pw++) {
            int idx = __ldg(a + (size_t)(ph * PW + pw) * C);
            if (idx >= lo && idx < hi) mine[(idx - lo) * 32 + lane] += __ldg(g + ph * PW + pw);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < W * 32; e += blockDim.x) {
    const int w = e >> 5, l = e & 31;
    if (c0 + l >= C) continue;
    float v = line[e];
#pragma unroll
    for (int k = 1; k < kRoiWarps; k++) v += line[(size_t)k * W * 32 + e];          // chunk order: RoIs ascending
    const size_t o = ((size_t)h * W + w) * C + c0 + l;
    if (addend) v += __ldg(addend + o);
    dfm[o] = v;
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
  if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(fm) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(argmax)) & 15) == 0) {
    size_t smem4 = (size_t)kRoiSlab * PH * PW * sizeof(float);
    FRCNN_REQUIRE(smem4 <= 200 * 1024, "roi_pool_fwd: pooled size too large");
    if (smem4 > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(roi_pool_fwd_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
      if (e != cudaSuccess) return cuda_fail(e, "roi_pool_fwd: smem attribute");
    }
    launch(roi_pool_fwd_v4_kernel, dim3(K, ceil_div(C, kRoiSlab)), kRoiWarps * 32, smem4, as_stream(stream), fm, H, W, C, proposals, PH, PW, spatial_scale, out, argmax);
    FRCNN_CHECK_LAUNCH("roi_pool_fwd_v4_kernel");
    return FRCNN_OK;
  }
  size_t smem = (size_t)kRoiChannels * PH * PW * sizeof(float);
  FRCNN_REQUIRE(smem <= 200 * 1024, "roi_pool_fwd: pooled size too large");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(roi_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "roi_pool_fwd: smem attribute");
  }
  launch(roi_pool_fwd_kernel, dim3(K, ceil_div(C, kRoiChannels)), kRoiWarps * 32, smem, as_stream(stream), fm, H, W, C, proposals, PH, PW, spatial_scale, out, argmax);
  FRCNN_CHECK_LAUNCH("roi_pool_fwd_kernel");
  return FRCNN_OK;
}

int frcnn_roi_pool_bwd(const float *dout, const int32_t *argmax, const float *proposals, int K, int H, int W, int C, int PH, int PW,
                       float spatial_scale, const float *addend, float *dfm, void *stream)
{
  FRCNN_REQUIRE(dout && argmax && proposals && dfm && K >= 0 && H > 0 && W > 0 && C > 0 && PH > 0 && PW > 0, "roi_pool_bwd: bad argument");
  FRCNN_REQUIRE(H < 32768, "roi_pool_bwd: feature map too tall");
  size_t smem = (size_t)kRoiWarps * W * 32 * sizeof(float) + (size_t)K * PH * sizeof(int32_t);
  FRCNN_REQUIRE(smem <= 200 * 1024, "roi_pool_bwd: feature map too wide / too many RoIs for the shared-memory buffers");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(roi_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "roi_pool_bwd: smem attribute");
  }
  launch(roi_pool_bwd_kernel, dim3(ceil_div(C, 32), H), kRoiWarps * 32, smem, as_stream(stream), dout, argmax, proposals, spatial_scale, K, H, W, C, PH, PW, addend, dfm);
  FRCNN_CHECK_LAUNCH("roi_pool_bwd_kernel");
  return FRCNN_OK;
}

}  // extern "C"
