// Two narrow output heads on one shared input: the RPN's 1x1 convolutions (512 -> 9 sigmoid scores, 512 -> 36 box deltas,
// models/rpn.py:88-96) and the detector's linear heads (4096 -> 21 class logits, 4096 -> 80 box deltas, models/detector.py:76-78).
//
// These are GEMMs with a tiny N (45 / 101 columns in total) and are pure latency for either implicit-GEMM engine (a 128-wide tile
// is mostly padding, K = 21 data-gradient chains, scalar filter-gradient fallbacks).  Here both heads of a stage are ONE problem:
//   forward   y[m][n] = act_h(sum_k x[m][k] * w_h[n][k] + b_h[n])       one warp per input row, lanes split K, weights staged in
//                                                                        shared memory and shared by the CTA's 8 rows; K sliced over
//                                                                        blockIdx.z when M is small (deterministic second pass)
//   dgrad     dx[m][k] = sum_n dz[m][n] * w[n][k]                       one thread per (row, 4 consecutive k), N-long loop
//   wgrad     dw[n][k] = sum_m dz[m][n] * x[m][k], db[n] = sum_m dz[m][n]   one thread per (n, 4 k) and row slice, fixed-order
//                                                                        reduction over the slices
// with dz = dy for a linear head and dy * y * (1 - y) for a sigmoid head.  All sums are plain fp32 FMAs in a fixed order.
#include "common.cuh"

namespace frcnn {

constexpr int kHeadRows = 8;         // rows (warps) per forward CTA
constexpr int kHeadCols = 32;        // output columns per forward CTA
constexpr int kHeadChunk = 128;      // k values staged per iteration (one float4 per lane)

struct Heads {
  const float *w1, *b1, *w2, *b2;    // (N1, K), (N1), (N2, K), (N2)
  int N1, N2, act1, act2;
};

__device__ __forceinline__ float head_act(float v, int act)
{
  return act == FRCNN_ACT_SIGMOID ? 1.0f / (1.0f + expf(-v)) : (act == FRCNN_ACT_RELU ? (v > 0.f ? v : 0.f) : v);
}

// grid (ceil(M / 8), ceil(N / 32), kslices); block 256.  partial == nullptr: one K slice, bias + activation applied here.
__global__ void __launch_bounds__(256)
heads_fwd_kernel(const float *__restrict__ x, int M, int K, Heads h, float *__restrict__ y1, float *__restrict__ y2, float *__restrict__ partial, int k_per_slice)
{
  pdl_enter();
  __shared__ __align__(16) float ws[kHeadCols][kHeadChunk + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * kHeadRows + warp;
  const int n0 = blockIdx.y * kHeadCols;
  const int N = h.N1 + h.N2;
  const int k_begin = blockIdx.z * k_per_slice;
  const int k_end = min(K, k_begin + k_per_slice);
  float acc[kHeadCols];
#pragma unroll
  for (int j = 0; j < kHeadCols; j++) acc[j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += kHeadChunk) {
    __syncthreads();
    // stage w[n0 .. n0+32)[k0 .. k0+128): 32 rows x 32 float4, coalesced along k
    for (int e = threadIdx.x; e < kHeadCols * (kHeadChunk / 4); e += 256) {
      const int r = e / (kHeadChunk / 4), q = e - r * (kHeadChunk / 4);
      const int n = n0 + r, k = k0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N && k < k_end) {
        const float *row = n < h.N1 ? h.w1 + (size_t)n * K : h.w2 + (size_t)(n - h.N1) * K;
        v = __ldg(reinterpret_cast<const float4 *>(row + k));
      }
      *reinterpret_cast<float4 *>(&ws[r][q * 4]) = v;
    }
    __syncthreads();
    const int k = k0 + lane * 4;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M && k < k_end) xv = __ldg(reinterpret_cast<const float4 *>(x + (size_t)m * K + k));
#pragma unroll
    for (int j = 0; j < kHeadCols; j++) {
      const float4 wv = *reinterpret_cast<const float4 *>(&ws[j][lane * 4]);
      acc[j] = fmaf(xv.x, wv.x, acc[j]);
      acc[j] = fmaf(xv.y, wv.y, acc[j]);
      acc[j] = fmaf(xv.z, wv.z, acc[j]);
      acc[j] = fmaf(xv.w, wv.w, acc[j]);
    }
  }
  // warp reduction of every column (lanes hold disjoint k), lane j keeps column j
  float mine = 0.f;
#pragma unroll
  for (int j = 0; j < kHeadCols; j++) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == j) mine = v;
  }
  const int n = n0 + lane;
  if (m >= M || n >= N) return;
  if (partial) {
    partial[((size_t)blockIdx.z * M + m) * N + n] = mine;
    return;
  }
  if (n < h.N1) y1[(size_t)m * h.N1 + n] = head_act(mine + (h.b1 ? __ldg(h.b1 + n) : 0.f), h.act1);
  else y2[(size_t)m * h.N2 + (n - h.N1)] = head_act(mine + (h.b2 ? __ldg(h.b2 + n - h.N1) : 0.f), h.act2);
}

__global__ void heads_fwd_reduce_kernel(const float *__restrict__ partial, int M, Heads h, int kslices, float *__restrict__ y1, float *__restrict__ y2)
{
  pdl_enter();
  const int N = h.N1 + h.N2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * N) return;
  const int m = e / N, n = e - m * N;
  float v = 0.f;
  for (int s = 0; s < kslices; s++) v += partial[((size_t)s * M + m) * N + n];       // slice order: ascending k
  if (n < h.N1) y1[(size_t)m * h.N1 + n] = head_act(v + (h.b1 ? __ldg(h.b1 + n) : 0.f), h.act1);
  else y2[(size_t)m * h.N2 + (n - h.N1)] = head_act(v + (h.b2 ? __ldg(h.b2 + n - h.N1) : 0.f), h.act2);
}

// dz[m][n] for both heads into one (M, N1 + N2) buffer
__global__ void heads_dz_kernel(const float *__restrict__ dy1, const float *__restrict__ y1, const float *__restrict__ dy2, const float *__restrict__ y2,
                                int M, Heads h, float *__restrict__ dz)
{
  pdl_enter();
  const int N = h.N1 + h.N2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * N) return;
  const int m = e / N, n = e - m * N;
  float g, y;
  int act;
  if (n < h.N1) { g = dy1[(size_t)m * h.N1 + n]; y = y1 ? y1[(size_t)m * h.N1 + n] : 0.f; act = h.act1; }
  else { g = dy2[(size_t)m * h.N2 + n - h.N1]; y = y2 ? y2[(size_t)m * h.N2 + n - h.N1] : 0.f; act = h.act2; }
  if (act == FRCNN_ACT_SIGMOID) g = g * (1.f - y) * y;                                 // sigmoid_bwd_kernel's expression (torch's order)
  else if (act == FRCNN_ACT_RELU) g = y > 0.f ? g : 0.f;
  dz[e] = g;
}

// dx[m][k..k+4) = sum_n dz[m][n] * w[n][k..k+4)   (n ascending);  grid over M * K / 4 threads
__global__ void heads_dgrad_kernel(const float *__restrict__ dz, int M, int K, Heads h, const float *__restrict__ addend, float *__restrict__ dx)
{
  pdl_enter();
  const int N = h.N1 + h.N2;
  const int k4 = K / 4;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)M * k4) return;
  const int m = (int)(e / k4), q = (int)(e - (size_t)m * k4);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const float *dzr = dz + (size_t)m * N;
  for (int n = 0; n < N; n++) {
    const float g = __ldg(dzr + n);
    const float *row = n < h.N1 ? h.w1 + (size_t)n * K : h.w2 + (size_t)(n - h.N1) * K;
    const float4 wv = __ldg(reinterpret_cast<const float4 *>(row) + q);
    s.x = fmaf(g, wv.x, s.x); s.y = fmaf(g, wv.y, s.y); s.z = fmaf(g, wv.z, s.z); s.w = fmaf(g, wv.w, s.w);
  }
  if (addend) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(addend) + e);
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
  }
  reinterpret_cast<float4 *>(dx)[e] = s;
}

// partial[slice][n][k..k+4) = sum over the slice's rows of dz[m][n] * x[m][k..k+4);  grid (N, ceil(K/4 / 256), slices)
__global__ void __launch_bounds__(256)
heads_wgrad_kernel(const float *__restrict__ dz, const float *__restrict__ x, int M, int K, int N, int rows_per_slice,
                   float *__restrict__ partial, float *__restrict__ partial_b)
{
  pdl_enter();
  const int n = blockIdx.x;
  const int q = blockIdx.y * 256 + threadIdx.x;
  const int k4 = K / 4;
  const int m0 = blockIdx.z * rows_per_slice, m1 = min(M, m0 + rows_per_slice);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  float sb = 0.f;
  if (q < k4) {
    for (int m = m0; m < m1; m++) {
      const float g = __ldg(dz + (size_t)m * N + n);
      const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + (size_t)m * K) + q);
      s.x = fmaf(g, xv.x, s.x); s.y = fmaf(g, xv.y, s.y); s.z = fmaf(g, xv.z, s.z); s.w = fmaf(g, xv.w, s.w);
      sb += g;
    }
    reinterpret_cast<float4 *>(partial + ((size_t)blockIdx.z * N + n) * K)[q] = s;
    if (q == 0) partial_b[(size_t)blockIdx.z * N + n] = sb;
  }
}

__global__ void heads_wgrad_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ partial_b, int K, Heads h, int slices,
                                          float *__restrict__ dw1, float *__restrict__ db1, float *__restrict__ dw2, float *__restrict__ db2)
{
  pdl_enter();
  const int N = h.N1 + h.N2;
  const int k4 = K / 4;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (size_t)N * k4) {
    const int n = (int)(e / k4), q = (int)(e - (size_t)n * k4);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < slices; z++) {                                                  // slice order: ascending rows
      const float4 p = reinterpret_cast<const float4 *>(partial + ((size_t)z * N + n) * K)[q];
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    float *dst = n < h.N1 ? dw1 + (size_t)n * K : dw2 + (size_t)(n - h.N1) * K;
    reinterpret_cast<float4 *>(dst)[q] = s;
  }
  if (e < (size_t)N) {
    float sb = 0.f;
    for (int z = 0; z < slices; z++) sb += partial_b[(size_t)z * N + e];
    if ((int)e < h.N1) { if (db1) db1[e] = sb; }
    else if (db2) db2[e - h.N1] = sb;
  }
}

static void heads_plan(int M, int K, int N, int *kslices, int *k_per_slice, int *mslices, int *rows_per_slice)
{
  // forward: enough CTAs to cover the machine when there are few rows
  const int ctas = ceil_div(M, kHeadRows) * ceil_div(N, kHeadCols);
  int ks = 1;
  if (ctas < 2 * kNumSMs) ks = min(ceil_div(2 * kNumSMs, ctas), max(1, K / (4 * kHeadChunk)));
  int per = ceil_div(ceil_div(K, ks), kHeadChunk) * kHeadChunk;
  *kslices = ceil_div(K, per);
  *k_per_slice = per;
  // filter gradient: (n, k4-block) CTAs times row slices
  const int base = N * ceil_div(K / 4, 256);
  int ms = 1;
  if (base < 4 * kNumSMs) ms = min(ceil_div(4 * kNumSMs, base), max(1, M / 16));
  int rows = ceil_div(M, ms);
  *mslices = ceil_div(M, rows);
  *rows_per_slice = rows;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

size_t frcnn_heads_workspace_bytes(int M, int K, int N1, int N2)
{
  if (M <= 0 || K <= 0 || N1 <= 0 || N2 < 0) return 0;
  int ks, kper, ms, rows;
  const int N = N1 + N2;
  heads_plan(M, K, N, &ks, &kper, &ms, &rows);
  size_t fwd = ks > 1 ? (size_t)ks * M * N * sizeof(float) : 0;
  size_t bwd = (size_t)M * N * sizeof(float) + (size_t)ms * N * ((size_t)K + 1) * sizeof(float) + 256;
  return (fwd > bwd ? fwd : bwd) + 256;
}

int frcnn_heads_fwd(const float *x, int M, int K, const float *w1, const float *b1, int N1, int act1, const float *w2, const float *b2, int N2, int act2,
                    float *y1, float *y2, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(x && w1 && y1 && M > 0 && K > 0 && K % 4 == 0 && N1 > 0 && N2 >= 0 && (N2 == 0 || (w2 && y2)), "heads_fwd: bad argument");
  Heads h{w1, b1, w2, b2, N1, N2, act1, act2};
  const int N = N1 + N2;
  int ks, kper, ms, rows;
  heads_plan(M, K, N, &ks, &kper, &ms, &rows);
  cudaStream_t st = as_stream(stream);
  float *partial = nullptr;
  if (ks > 1) {
    if (workspace == nullptr || workspace_bytes < (size_t)ks * M * N * sizeof(float)) return fail(FRCNN_E_WORKSPACE, "heads_fwd: workspace too small");
    partial = reinterpret_cast<float *>(workspace);
  }
  launch(heads_fwd_kernel, dim3(ceil_div(M, kHeadRows), ceil_div(N, kHeadCols), ks), 256, 0, st, x, M, K, h, y1, y2, partial, kper);
  FRCNN_CHECK_LAUNCH("heads_fwd_kernel");
  if (ks > 1) {
    launch(heads_fwd_reduce_kernel, ceil_div(M * N, 256), 256, 0, st, partial, M, h, ks, y1, y2);
    FRCNN_CHECK_LAUNCH("heads_fwd_reduce_kernel");
  }
  return FRCNN_OK;
}

int frcnn_heads_bwd(const float *x, int M, int K, const float *w1, int N1, int act1, const float *y1, const float *dy1,
                    const float *w2, int N2, int act2, const float *y2, const float *dy2,
                    float *dx, float *dw1, float *db1, float *dw2, float *db2, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(x && w1 && dy1 && dw1 && M > 0 && K > 0 && K % 4 == 0 && N1 > 0 && N2 >= 0 && (N2 == 0 || (w2 && dy2 && dw2)), "heads_bwd: bad argument");
  FRCNN_REQUIRE((act1 == FRCNN_ACT_NONE || y1) && (act2 == FRCNN_ACT_NONE || N2 == 0 || y2), "heads_bwd: an activated head needs its forward output");
  Heads h{w1, nullptr, w2, nullptr, N1, N2, act1, act2};
  const int N = N1 + N2;
  int ks, kper, ms, rows;
  heads_plan(M, K, N, &ks, &kper, &ms, &rows);
  if (workspace == nullptr || workspace_bytes < frcnn_heads_workspace_bytes(M, K, N1, N2)) return fail(FRCNN_E_WORKSPACE, "heads_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  float *dz = reinterpret_cast<float *>(workspace);
  float *partial = dz + (((size_t)M * N + 63) / 64) * 64;
  float *partial_b = partial + (size_t)ms * N * K;
  launch(heads_dz_kernel, ceil_div(M * N, 256), 256, 0, st, dy1, y1, dy2, y2, M, h, dz);
  FRCNN_CHECK_LAUNCH("heads_dz_kernel");
  if (dx) {
    launch(heads_dgrad_kernel, (unsigned)ceil_div<size_t>((size_t)M * (K / 4), 256), 256, 0, st, dz, M, K, h, nullptr, dx);
    FRCNN_CHECK_LAUNCH("heads_dgrad_kernel");
  }
  launch(heads_wgrad_kernel, dim3(N, ceil_div(K / 4, 256), ms), 256, 0, st, dz, x, M, K, N, rows, partial, partial_b);
  FRCNN_CHECK_LAUNCH("heads_wgrad_kernel");
  launch(heads_wgrad_reduce_kernel, (unsigned)ceil_div<size_t>((size_t)N * (K / 4), 256), 256, 0, st, partial, partial_b, K, h, ms, dw1, db1, dw2, db2);
  FRCNN_CHECK_LAUNCH("heads_wgrad_reduce_kernel");
  return FRCNN_OK;
}

}  // extern "C"
