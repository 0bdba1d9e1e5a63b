// RoIAlign forward / backward -- EXTENSION (BASELINE configs 3 and 5): the reference only has
// RoIPool; semantics follow torchvision.ops.roi_align (sampling_ratio > 0 fixed grid, `aligned` flag), which defines
// this op.  Feature map NHWC, proposals (y1,x1,y2,x2), output (K,C,PH,PW).
//
// Forward: one CTA per (RoI, 128-channel slab).  The PH*PW*S*S bilinear taps of the RoI (4 corner offsets + 4
// weights each) are computed ONCE into shared memory by the CTA and then reused by every warp.  C % 4 == 0
// (roi_align_fwd_v4_kernel): 8 warps, a warp takes every 8th bin, its 32 lanes read 4 consecutive channels each
// (one 512-byte NHWC segment per warp-wide 128-bit load; the S*S*4 corner loads of a bin are independent, so
// 16 vector loads per lane are in flight against the L2-resident map); otherwise one channel per thread.  Results
// are staged in shared memory and streamed out as one contiguous (128 x PH*PW) run, exactly like roi_pool.cu.
// Backward: deterministic and atomics-free -- a thread owns one (feature row, channel) line, rebuilds the tap table
// per RoI in shared memory and accumulates the taps that land on its row in ascending (RoI, bin, sample) order.
#include <math.h>
#include "common.cuh"

namespace frcnn {

constexpr int kAlignSlab = 128;

struct Tap {
  int o00, o01, o10, o11;          // cell indices (h*W + w) of the four corners, -1 if the sample is outside
  float w00, w01, w10, w11;
};

// torchvision bilinear_interpolate / pre_calc_for_bilinear_interpolate in fp32
__device__ __forceinline__ Tap make_tap(float y, float x, int H, int W, int *rows)
{
  Tap t;
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
    t.o00 = t.o01 = t.o10 = t.o11 = -1;
    t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
    rows[0] = rows[1] = -1;
    return t;
  }
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
  float ly = y - (float)y_low, lx = x - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
  t.o00 = y_low * W + x_low; t.o01 = y_low * W + x_high; t.o10 = y_high * W + x_low; t.o11 = y_high * W + x_high;
  t.w00 = __fmul_rn(hy, hx); t.w01 = __fmul_rn(hy, lx); t.w10 = __fmul_rn(ly, hx); t.w11 = __fmul_rn(ly, lx);
  rows[0] = y_low; rows[1] = y_high;
  return t;
}

struct RoiGeom {
  float start_h, start_w, bin_h, bin_w;
};

__device__ __forceinline__ RoiGeom roi_geom(const float *__restrict__ p, float scale, int PH, int PW, int aligned)
{
  const float offset = aligned ? 0.5f : 0.f;
  RoiGeom r;
  r.start_w = __fsub_rn(__fmul_rn(p[1], scale), offset);
  r.start_h = __fsub_rn(__fmul_rn(p[0], scale), offset);
  float end_w = __fsub_rn(__fmul_rn(p[3], scale), offset), end_h = __fsub_rn(__fmul_rn(p[2], scale), offset);
  float rw = __fsub_rn(end_w, r.start_w), rh = __fsub_rn(end_h, r.start_h);
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  r.bin_h = __fdiv_rn(rh, (float)PH);
  r.bin_w = __fdiv_rn(rw, (float)PW);
  return r;
}

__device__ __forceinline__ void sample_xy(const RoiGeom &r, int ph, int pw, int iy, int ix, int S, float *y, float *x)
{
  // y = start_h + ph*bin_h + (iy + .5) * bin_h / S   (same operation order as the library op)
  *y = __fadd_rn(__fadd_rn(r.start_h, __fmul_rn((float)ph, r.bin_h)), __fdiv_rn(__fmul_rn((float)iy + 0.5f, r.bin_h), (float)S));
  *x = __fadd_rn(__fadd_rn(r.start_w, __fmul_rn((float)pw, r.bin_w)), __fdiv_rn(__fmul_rn((float)ix + 0.5f, r.bin_w), (float)S));
}

__global__ void __launch_bounds__(kAlignSlab)
roi_align_fwd_kernel(const float *__restrict__ fm, int H, int W, int C, const float *__restrict__ proposals, int PH, int PW, int S, float scale,
                     int aligned, float *__restrict__ out)
{
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const int bins = PH * PW, taps = bins * S * S;
  Tap *tab = reinterpret_cast<Tap *>(smem_raw);
  float *s_val = reinterpret_cast<float *>(tab + taps);
  const int n = blockIdx.x, c0 = blockIdx.y * kAlignSlab, c = c0 + threadIdx.x;
  const RoiGeom r = roi_geom(proposals + 4 * (size_t)n, scale, PH, PW, aligned);
  for (int e = threadIdx.x; e < taps; e += blockDim.x) {
    int s = e % (S * S), b = e / (S * S);
    int rows[2];
    float y, x;
    sample_xy(r, b / PW, b % PW, s / S, s % S, S, &y, &x);
    tab[e] = make_tap(y, x, H, W, rows);
  }
  __syncthreads();
  if (c < C) {
    for (int b = 0; b < bins; b++) {
      float acc = 0.f;
      for (int s = 0; s < S * S; s++) {
        const Tap tp = tab[b * S * S + s];
        if (tp.o00 < 0) continue;
        float v = __fmul_rn(tp.w00, __ldg(fm + (size_t)tp.o00 * C + c));
        v = __fadd_rn(v, __fmul_rn(tp.w01, __ldg(fm + (size_t)tp.o01 * C + c)));
        v = __fadd_rn(v, __fmul_rn(tp.w10, __ldg(fm + (size_t)tp.o10 * C + c)));
        v = __fadd_rn(v, __fmul_rn(tp.w11, __ldg(fm + (size_t)tp.o11 * C + c)));
        acc = __fadd_rn(acc, v);
      }
      s_val[threadIdx.x * bins + b] = __fdiv_rn(acc, (float)(S * S));      // output_val /= count
    }
  }
  __syncthreads();
  int live = min(kAlignSlab, C - c0);
  size_t base = ((size_t)n * C + c0) * bins;
  for (int e = threadIdx.x; e < live * bins; e += blockDim.x) out[base + e] = s_val[e];
}

// ---- forward, 4 channels per lane; same operation order per channel as the scalar kernel ------------------------------
constexpr int kAlignWarps = 8;

// volatile: keeps the corner loads of a round back to back (see roi_pool.cu)
__device__ __forceinline__ float4 ldg_nc_v4(const float4 *p)
{
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void tap_fma4(float4 &v, float w, const float4 f, bool first)
{
  if (first) { v.x = __fmul_rn(w, f.x); v.y = __fmul_rn(w, f.y); v.z = __fmul_rn(w, f.z); v.w = __fmul_rn(w, f.w); }
  else {
    v.x = __fadd_rn(v.x, __fmul_rn(w, f.x)); v.y = __fadd_rn(v.y, __fmul_rn(w, f.y));
    v.z = __fadd_rn(v.z, __fmul_rn(w, f.z)); v.w = __fadd_rn(v.w, __fmul_rn(w, f.w));
  }
}

__global__ void __launch_bounds__(kAlignWarps * 32)
roi_align_fwd_v4_kernel(const float *__restrict__ fm, int H, int W, int C, const float *__restrict__ proposals, int PH, int PW, int S, float scale,
                        int aligned, float *__restrict__ out)
{
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const int bins = PH * PW, ss = S * S, taps = bins * ss;
  Tap *tab = reinterpret_cast<Tap *>(smem_raw);
  float *s_val = reinterpret_cast<float *>(tab + taps);
  const int n = blockIdx.x, c0 = blockIdx.y * kAlignSlab;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + 4 * lane;
  const RoiGeom r = roi_geom(proposals + 4 * (size_t)n, scale, PH, PW, aligned);
  for (int e = threadIdx.x; e < taps; e += blockDim.x) {
    int s = e % ss, b = e / ss;
    int rows[2];
    float y, x;
    sample_xy(r, b / PW, b % PW, s / S, s % S, S, &y, &x);
    tab[e] = make_tap(y, x, H, W, rows);
  }
  __syncthreads();
  if (c < C) {
    const float4 *fm4 = reinterpret_cast<const float4 *>(fm + c);
    const int C4 = C >> 2;
    const float count = (float)ss;
    for (int b = warp; b < bins; b += kAlignWarps) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // two samples (eight corner loads, issued back to back) per round; a sample outside the map keeps its place in the
      // sequence but reads cell 0 and is not accumulated (the table is per RoI, so the skip is warp-uniform)
      for (int s = 0; s < ss; s += 2) {
        Tap tp[2];
        float4 f[2][4];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          tp[u] = tab[b * ss + min(s + u, ss - 1)];
          if (s + u >= ss) tp[u].o00 = -1;
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const bool ok = tp[u].o00 >= 0;
          f[u][0] = ldg_nc_v4(fm4 + (size_t)(ok ? tp[u].o00 : 0) * C4);
          f[u][1] = ldg_nc_v4(fm4 + (size_t)(ok ? tp[u].o01 : 0) * C4);
          f[u][2] = ldg_nc_v4(fm4 + (size_t)(ok ? tp[u].o10 : 0) * C4);
          f[u][3] = ldg_nc_v4(fm4 + (size_t)(ok ? tp[u].o11 : 0) * C4);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          if (tp[u].o00 < 0) continue;
          float4 v;
          tap_fma4(v, tp[u].w00, f[u][0], true);
          tap_fma4(v, tp[u].w01, f[u][1], false);
          tap_fma4(v, tp[u].w10, f[u][2], false);
          tap_fma4(v, tp[u].w11, f[u][3], false);
          acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
      }
      const int o = (4 * lane) * bins + b;
      s_val[o] = __fdiv_rn(acc.x, count);                           // output_val /= count
      s_val[o + bins] = __fdiv_rn(acc.y, count);
      s_val[o + 2 * bins] = __fdiv_rn(acc.z, count);
      s_val[o + 3 * bins] = __fdiv_rn(acc.w, count);
    }
  }
  __syncthreads();
  const int live = min(kAlignSlab, C - c0);
  const size_t base = ((size_t)n * C + c0) * bins;
  const int total4 = (live * bins) >> 2;
  float4 *o4 = reinterpret_cast<float4 *>(out + base);
  const float4 *sv4 = reinterpret_cast<const float4 *>(s_val);
  for (int e = threadIdx.x; e < total4; e += blockDim.x) __stcs(o4 + e, sv4[e]);
}

__global__ void __launch_bounds__(256)
roi_align_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ proposals, float scale, int aligned, int K, int H, int W, int C,
                     int PH, int PW, int S, const float *__restrict__ addend, float *__restrict__ dfm)
{
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const int bins = PH * PW, taps = bins * S * S;
  float *line = reinterpret_cast<float *>(smem_raw);                      // [W][256]
  Tap *tab = reinterpret_cast<Tap *>(line + (size_t)W * 256);
  int *trow = reinterpret_cast<int *>(tab + taps);                         // [taps][2] rows touched by each tap
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane, h = blockIdx.y * 8 + warp;
  const bool live = c < C && h < H;
  for (int w = 0; w < W; w++) line[w * 256 + threadIdx.x] = 0.f;
  const float inv_count = 1.0f / (float)(S * S);
  for (int n = 0; n < K; n++) {
    __syncthreads();
    const RoiGeom r = roi_geom(proposals + 4 * (size_t)n, scale, PH, PW, aligned);
    for (int e = threadIdx.x; e < taps; e += blockDim.x) {
      int s = e % (S * S), b = e / (S * S);
      float y, x;
      sample_xy(r, b / PW, b % PW, s / S, s % S, S, &y, &x);
      tab[e] = make_tap(y, x, H, W, trow + 2 * e);
    }
    __syncthreads();
    if (!live) continue;
    const float *g = dout + ((size_t)n * C + c) * bins;
    for (int e = 0; e < taps; e++) {
      const int r0 = trow[2 * e], r1 = trow[2 * e + 1];
      if (r0 != h && r1 != h) continue;
      const Tap tp = tab[e];
      const float gv = __fmul_rn(__ldg(g + e / (S * S)), inv_count);
      if (r0 == h) {
        line[(tp.o00 - h * W) * 256 + threadIdx.x] += __fmul_rn(gv, tp.w00);
        line[(tp.o01 - h * W) * 256 + threadIdx.x] += __fmul_rn(gv, tp.w01);
      }
      if (r1 == h) {
        line[(tp.o10 - h * W) * 256 + threadIdx.x] += __fmul_rn(gv, tp.w10);
        line[(tp.o11 - h * W) * 256 + threadIdx.x] += __fmul_rn(gv, tp.w11);
      }
    }
  }
  if (live)
    for (int w = 0; w < W; w++) {
      size_t o = ((size_t)h * W + w) * C + c;
      float v = line[w * 256 + threadIdx.x];
      if (addend) v += __ldg(addend + o);
      dfm[o] = v;
    }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_roi_align_fwd(const float *fm, int H, int W, int C, const float *proposals, int K, int PH, int PW, float spatial_scale,
                        int sampling_ratio, int aligned, float *out, void *stream)
{
  FRCNN_REQUIRE(fm && proposals && out && H > 0 && W > 0 && C > 0 && K >= 0 && PH > 0 && PW > 0, "roi_align_fwd: bad argument");
  FRCNN_REQUIRE(sampling_ratio > 0 && sampling_ratio <= 4, "roi_align_fwd: sampling_ratio must be in 1..4 (adaptive grids are not implemented)");
  if (K == 0) return FRCNN_OK;
  const int taps = PH * PW * sampling_ratio * sampling_ratio;
  size_t smem = (size_t)taps * sizeof(Tap) + (size_t)kAlignSlab * PH * PW * sizeof(float);
  FRCNN_REQUIRE(smem <= 200 * 1024, "roi_align_fwd: pooled size too large");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(roi_align_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "roi_align_fwd: smem attribute");
  }
  if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(fm) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(roi_align_fwd_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "roi_align_fwd: smem attribute");
    }
    launch(roi_align_fwd_v4_kernel, dim3(K, ceil_div(C, kAlignSlab)), kAlignWarps * 32, smem, as_stream(stream), fm, H, W, C, proposals, PH, PW, sampling_ratio, spatial_scale, aligned, out);
    FRCNN_CHECK_LAUNCH("roi_align_fwd_v4_kernel");
    return FRCNN_OK;
  }
  launch(roi_align_fwd_kernel, dim3(K, ceil_div(C, kAlignSlab)), kAlignSlab, smem, as_stream(stream), fm, H, W, C, proposals, PH, PW, sampling_ratio, spatial_scale, aligned, out);
  FRCNN_CHECK_LAUNCH("roi_align_fwd_kernel");
  return FRCNN_OK;
}

int frcnn_roi_align_bwd(const float *dout, const float *proposals, int K, int H, int W, int C, int PH, int PW, float spatial_scale,
                        int sampling_ratio, int aligned, const float *addend, float *dfm, void *stream)
{
  FRCNN_REQUIRE(dout && proposals && dfm && K >= 0 && H > 0 && W > 0 && C > 0 && PH > 0 && PW > 0, "roi_align_bwd: bad argument");
  FRCNN_REQUIRE(sampling_ratio > 0 && sampling_ratio <= 4, "roi_align_bwd: sampling_ratio must be in 1..4");
  const int taps = PH * PW * sampling_ratio * sampling_ratio;
  size_t smem = (size_t)W * 256 * sizeof(float) + (size_t)taps * (sizeof(Tap) + 2 * sizeof(int));
  FRCNN_REQUIRE(smem <= 200 * 1024, "roi_align_bwd: feature map too wide for the shared-memory line buffer");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(roi_align_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "roi_align_bwd: smem attribute");
  }
  launch(roi_align_bwd_kernel, dim3(ceil_div(C, 32), ceil_div(H, 8)), 256, smem, as_stream(stream), dout, proposals, spatial_scale, aligned, K, H, W, C, PH, PW, sampling_ratio, addend, dfm);
  FRCNN_CHECK_LAUNCH("roi_align_bwd_kernel");
  return FRCNN_OK;
}

}  // extern "C"
