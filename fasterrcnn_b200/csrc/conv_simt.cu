// fp32 CUDA-core implicit-GEMM convolution: forward, data gradient, filter gradient.
// Engine FRCNN_ENGINE_SIMT_FP32: exact fp32 FMA accumulation.  It is (a) the numerically exact
// engine that the tensor-core engine (conv_tc.cu) is validated against and (b) the engine for
// shapes the tensor-core path does not take (Cin=3 stem, tiny heads).
//
// GEMM views (activations NHWC, filters OHWI):
//   FWD   C[M=N*Ho*Wo][Cout]      = A[M][K=KH*KW*Cin]  (gathered x)  * B[Cout][K]   (w, K-major)
//   DGRAD C[M=N*H*W][Cin]         = A[M][K=KH*KW*Cout] (gathered dy) * B[K][Cin]    (w, N-major)
//   WGRAD C[M=Cout][KH*KW*Cin]    = A[K=N*Ho*Wo][Cout] (dy, M-major) * B[K][KH*KW*Cin] (gathered x)
// Tiles: 128 x {128,64} x 16, 256 threads, 8x{8,4} register micro-tiles, double-buffered shared
// memory with register prefetch, optional deterministic split-K (partials -> reduce+epilogue).
#include "common.cuh"

namespace frcnn {

struct ConvGeom {
  int N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo;
};

enum { MODE_FWD = 0, MODE_DGRAD = 1 };

struct Epilogue {
  const float *scale;
  const float *bias;
  const float *residual;
  int act;
};

__device__ __forceinline__ float apply_act(float v, int act)
{
  if (act == FRCNN_ACT_RELU) return v > 0.0f ? v : 0.0f;
  if (act == FRCNN_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

__device__ __forceinline__ float epilogue_one(float v, int m, int n, int Nn, const Epilogue &e)
{
  if (e.scale) v *= e.scale[n];
  if (e.bias) v += e.bias[n];
  if (e.residual) v += e.residual[(size_t)m * Nn + n];
  return apply_act(v, e.act);
}

constexpr int BK = 16;

// ------------------------------------------------------------------------------------------
// FWD / DGRAD kernel
// ------------------------------------------------------------------------------------------
template <int MODE, int BM, int BN, bool VEC>
__global__ void __launch_bounds__(256, 2)
igemm_kernel(const float *__restrict__ src, const float *__restrict__ wgt, float *__restrict__ dst,
             Epilogue epi, ConvGeom g, int M, int Nn, int K, int C, int chunks_per_split, int splits)
{
  pdl_enter();
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_LD = BM / 64;                  // float4 loads per thread for the A tile
  constexpr int B_LD = BN / 64;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int t = threadIdx.x;
  const int bm0 = blockIdx.x * BM;
  const int bn0 = blockIdx.y * BN;
  const int z = blockIdx.z;
  const int total_chunks = (K + BK - 1) / BK;
  const int chunk_begin = z * chunks_per_split;
  int chunk_end = chunk_begin + chunks_per_split;
  if (chunk_end > total_chunks) chunk_end = total_chunks;
  const int KHW = g.KH * g.KW;

  // per-thread row decode for the A loads
  constexpr int A_ROWS = VEC ? A_LD : (BM * BK / 256);
  int row_n[A_ROWS], row_y[A_ROWS], row_x[A_ROWS];
  bool row_ok[A_ROWS];
#pragma unroll
  for (int i = 0; i < A_ROWS; i++) {
    int r = VEC ? ((t >> 2) + 64 * i) : ((t + 256 * i) / BK);
    int m = bm0 + r;
    row_ok[i] = m < M;
    int mm = row_ok[i] ? m : 0;
    if constexpr (MODE == MODE_FWD) {
      int hw = g.Ho * g.Wo;
      int n = mm / hw, rem = mm - n * hw;
      int oh = rem / g.Wo, ow = rem - oh * g.Wo;
      row_n[i] = n; row_y[i] = oh * g.stride - g.pad; row_x[i] = ow * g.stride - g.pad;
    } else {
      int hw = g.H * g.W;
      int n = mm / hw, rem = mm - n * hw;
      int ih = rem / g.W, iw = rem - ih * g.W;
      row_n[i] = n; row_y[i] = ih + g.pad; row_x[i] = iw + g.pad;
    }
  }

  // source offset (in floats) of row i for tap (kh,kw), or -1 when the tap falls outside
  auto src_offset = [&](int i, int kh, int kw) -> long long {
    if (!row_ok[i]) return -1;
    if constexpr (MODE == MODE_FWD) {
      int ih = row_y[i] + kh, iw = row_x[i] + kw;
      if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return -1;
      return ((long long)(row_n[i] * g.H + ih) * g.W + iw) * g.Cin;
    } else {
      int ty = row_y[i] - kh, tx = row_x[i] - kw;
      if (ty < 0 || tx < 0) return -1;
      int oh = ty / g.stride, ow = tx / g.stride;
      if (oh * g.stride != ty || ow * g.stride != tx || oh >= g.Ho || ow >= g.Wo) return -1;
      return ((long long)(row_n[i] * g.Ho + oh) * g.Wo + ow) * g.Cout;
    }
  };

  float4 a_pref[VEC ? A_LD : 1];
  float4 b_pref[VEC ? B_LD : 1];
  float a_s[VEC ? 1 : A_ROWS];
  float b_s[VEC ? 1 : (BN * BK / 256)];

  auto load_tile = [&](int chunk) {
    const int k0 = chunk * BK;
    if constexpr (VEC) {
      const int kq = (t & 3) * 4;
      const int k = k0 + kq;
      const bool k_ok = k < K;
      int tap = k_ok ? k / C : 0;
      int c = k - tap * C;
      int kh = tap / g.KW, kw = tap - kh * g.KW;
#pragma unroll
      for (int i = 0; i < A_LD; i++) {
        long long off = k_ok ? src_offset(i, kh, kw) : -1;
        a_pref[i] = off >= 0 ? __ldg(reinterpret_cast<const float4 *>(src + off + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if constexpr (MODE == MODE_FWD) {
#pragma unroll
        for (int i = 0; i < B_LD; i++) {
          int n = bn0 + (t >> 2) + 64 * i;
          b_pref[i] = (k_ok && n < Nn) ? __ldg(reinterpret_cast<const float4 *>(wgt + (size_t)n * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        // B is N-major: row k = (tap, co) lives at w[(co*KHW + tap)*Cin + n]
        constexpr int TPR = BN / 4;                // threads per k-row
        constexpr int RPP = 256 / TPR;             // k-rows per pass
#pragma unroll
        for (int i = 0; i < B_LD; i++) {
          int kk = t / TPR + RPP * i;
          int kb = k0 + kk;
          int n = bn0 + (t % TPR) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kb < K && n < Nn) {
            int tapb = kb / C, co = kb - tapb * C;
            v = __ldg(reinterpret_cast<const float4 *>(wgt + ((size_t)co * KHW + tapb) * g.Cin + n));
          }
          b_pref[i] = v;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_ROWS; i++) {
        int e = t + 256 * i;
        int kk = e % BK;
        int k = k0 + kk;
        float v = 0.f;
        if (k < K) {
          int tap = k / C, c = k - tap * C;
          int kh = tap / g.KW, kw = tap - kh * g.KW;
          long long off = src_offset(i, kh, kw);
          if (off >= 0) v = __ldg(src + off + c);
        }
        a_s[i] = v;
      }
#pragma unroll
      for (int i = 0; i < BN * BK / 256; i++) {
        int e = t + 256 * i;
        float v = 0.f;
        if constexpr (MODE == MODE_FWD) {
          int kk = e % BK, nn = e / BK;
          int k = k0 + kk, n = bn0 + nn;
          if (k < K && n < Nn) v = __ldg(wgt + (size_t)n * K + k);
        } else {
          int nn = e % BN, kk = e / BN;
          int k = k0 + kk, n = bn0 + nn;
          if (k < K && n < Nn) {
            int tapb = k / C, co = k - tapb * C;
            v = __ldg(wgt + ((size_t)co * KHW + tapb) * g.Cin + n);
          }
        }
        b_s[i] = v;
      }
    }
  };

  auto store_tile = [&](int buf) {
    if constexpr (VEC) {
      const int kq = (t & 3) * 4;
#pragma unroll
      for (int i = 0; i < A_LD; i++) {
        int r = (t >> 2) + 64 * i;
        As[buf][kq + 0][r] = a_pref[i].x; As[buf][kq + 1][r] = a_pref[i].y;
        As[buf][kq + 2][r] = a_pref[i].z; As[buf][kq + 3][r] = a_pref[i].w;
      }
      if constexpr (MODE == MODE_FWD) {
#pragma unroll
        for (int i = 0; i < B_LD; i++) {
          int r = (t >> 2) + 64 * i;
          Bs[buf][kq + 0][r] = b_pref[i].x; Bs[buf][kq + 1][r] = b_pref[i].y;
          Bs[buf][kq + 2][r] = b_pref[i].z; Bs[buf][kq + 3][r] = b_pref[i].w;
        }
      } else {
        constexpr int TPR = BN / 4;
        constexpr int RPP = 256 / TPR;
#pragma unroll
        for (int i = 0; i < B_LD; i++) {
          int kk = t / TPR + RPP * i;
          *reinterpret_cast<float4 *>(&Bs[buf][kk][(t % TPR) * 4]) = b_pref[i];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_ROWS; i++) {
        int e = t + 256 * i;
        As[buf][e % BK][e / BK] = a_s[i];
      }
#pragma unroll
      for (int i = 0; i < BN * BK / 256; i++) {
        int e = t + 256 * i;
        if constexpr (MODE == MODE_FWD) Bs[buf][e % BK][e / BK] = b_s[i];
        else Bs[buf][e / BN][e % BN] = b_s[i];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const int tx = t & 15, ty = t >> 4;

  if (chunk_begin < chunk_end) {
    load_tile(chunk_begin);
    store_tile(0);
  }
  __syncthreads();
  int cur = 0;
  for (int chunk = chunk_begin; chunk < chunk_end; chunk++) {
    const bool has_next = chunk + 1 < chunk_end;
    if (has_next) load_tile(chunk + 1);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[TM], b[TN];
      {
        float4 v = *reinterpret_cast<const float4 *>(&As[cur][kk][ty * 4]);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        if (TM == 8) {
          float4 u = *reinterpret_cast<const float4 *>(&As[cur][kk][BM / 2 + ty * 4]);
          a[TM - 4] = u.x; a[TM - 3] = u.y; a[TM - 2] = u.z; a[TM - 1] = u.w;
        }
      }
      {
        float4 v = *reinterpret_cast<const float4 *>(&Bs[cur][kk][tx * 4]);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        if (TN == 8) {
          float4 u = *reinterpret_cast<const float4 *>(&Bs[cur][kk][BN / 2 + tx * 4]);
          b[TN - 4] = u.x; b[TN - 3] = u.y; b[TN - 2] = u.z; b[TN - 1] = u.w;
        }
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) store_tile(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

  // epilogue / partial store
  float *out = dst;
  if (splits > 1) out = dst + (size_t)z * M * Nn;
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int m = bm0 + ((TM == 8 && i >= 4) ? (BM / 2 + ty * 4 + i - 4) : (ty * 4 + i));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < TN / 4; jh++) {
      int n = bn0 + (jh == 0 ? tx * 4 : BN / 2 + tx * 4);
      if (n >= Nn) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = acc[i][jh * 4 + j];
      if (splits == 1) {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (n + j < Nn) v[j] = epilogue_one(v[j], m, n + j, Nn, epi);
      }
      float *p = out + (size_t)m * Nn + n;
      if ((Nn & 3) == 0) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (n + j < Nn) p[j] = v[j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// WGRAD kernel (A = dy M-major, B = gathered x N-major, K = output pixels)
// ------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(256, 2)
wgrad_kernel(const float *__restrict__ dy, const float *__restrict__ x, float *__restrict__ dst,
             ConvGeom g, int M, int Nn, int Kpix, int chunks_per_split, int splits)
{
  pdl_enter();
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_TPR = BM / 4, A_RPP = 256 / A_TPR, A_LD = BK / A_RPP;
  constexpr int B_TPR = BN / 4, B_RPP = 256 / B_TPR, B_LD = BK / B_RPP;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int t = threadIdx.x;
  const int bm0 = blockIdx.x * BM;
  const int bn0 = blockIdx.y * BN;
  const int z = blockIdx.z;
  const int total_chunks = (Kpix + BK - 1) / BK;
  const int chunk_begin = z * chunks_per_split;
  int chunk_end = chunk_begin + chunks_per_split;
  if (chunk_end > total_chunks) chunk_end = total_chunks;

  // this thread's B column quad: n -> (tap, ci), fixed for the whole K loop
  const int nb = bn0 + (t % B_TPR) * 4;
  const bool nb_ok = nb < Nn;
  int b_kh = 0, b_kw = 0, b_ci = 0;
  if (nb_ok) {
    int tap = nb / g.Cin;
    b_ci = nb - tap * g.Cin;
    b_kh = tap / g.KW;
    b_kw = tap - b_kh * g.KW;
  }
  const int ma = bm0 + (t % A_TPR) * 4;
  const bool ma_ok = ma < M;
  const int HoWo = g.Ho * g.Wo;

  float4 a_pref[A_LD], b_pref[B_LD];
  auto load_tile = [&](int chunk) {
    const int k0 = chunk * BK;
#pragma unroll
    for (int i = 0; i < A_LD; i++) {
      int k = k0 + t / A_TPR + A_RPP * i;
      a_pref[i] = (ma_ok && k < Kpix) ? __ldg(reinterpret_cast<const float4 *>(dy + (size_t)k * M + ma)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_LD; i++) {
      int k = k0 + t / B_TPR + B_RPP * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nb_ok && k < Kpix) {
        int n = k / HoWo, rem = k - n * HoWo;
        int oh = rem / g.Wo, ow = rem - oh * g.Wo;
        int ih = oh * g.stride - g.pad + b_kh, iw = ow * g.stride - g.pad + b_kw;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
          v = __ldg(reinterpret_cast<const float4 *>(x + ((size_t)(n * g.H + ih) * g.W + iw) * g.Cin + b_ci));
      }
      b_pref[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; i++) *reinterpret_cast<float4 *>(&As[buf][t / A_TPR + A_RPP * i][(t % A_TPR) * 4]) = a_pref[i];
#pragma unroll
    for (int i = 0; i < B_LD; i++) *reinterpret_cast<float4 *>(&Bs[buf][t / B_TPR + B_RPP * i][(t % B_TPR) * 4]) = b_pref[i];
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;
  const int tx = t & 15, ty = t >> 4;

  if (chunk_begin < chunk_end) {
    load_tile(chunk_begin);
    store_tile(0);
  }
  __syncthreads();
  int cur = 0;
  for (int chunk = chunk_begin; chunk < chunk_end; chunk++) {
    const bool has_next = chunk + 1 < chunk_end;
    if (has_next) load_tile(chunk + 1);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[TM], b[TN];
      {
        float4 v = *reinterpret_cast<const float4 *>(&As[cur][kk][ty * 4]);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        if (TM == 8) {
          float4 u = *reinterpret_cast<const float4 *>(&As[cur][kk][BM / 2 + ty * 4]);
          a[TM - 4] = u.x; a[TM - 3] = u.y; a[TM - 2] = u.z; a[TM - 1] = u.w;
        }
      }
      {
        float4 v = *reinterpret_cast<const float4 *>(&Bs[cur][kk][tx * 4]);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        if (TN == 8) {
          float4 u = *reinterpret_cast<const float4 *>(&Bs[cur][kk][BN / 2 + tx * 4]);
          b[TN - 4] = u.x; b[TN - 3] = u.y; b[TN - 2] = u.z; b[TN - 1] = u.w;
        }
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) store_tile(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

  float *out = dst + (splits > 1 ? (size_t)z * M * Nn : 0);
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int m = bm0 + ((TM == 8 && i >= 4) ? (BM / 2 + ty * 4 + i - 4) : (ty * 4 + i));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < TN / 4; jh++) {
      int n = bn0 + (jh == 0 ? tx * 4 : BN / 2 + tx * 4);
      if (n >= Nn) continue;
      *reinterpret_cast<float4 *>(out + (size_t)m * Nn + n) =
          make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
    }
  }
}

// scalar filter gradient for shapes the vector kernel does not take (Cin or Cout not a multiple of
// 4: the RGB stem, the 9/21-wide heads): one thread per filter element and K slice (blockIdx.y),
// fixed summation order inside a slice, slices combined by the deterministic split-K reduce.
__global__ void wgrad_scalar_kernel(const float *__restrict__ dy, const float *__restrict__ x, float *__restrict__ dst, ConvGeom g, int pix_per_slice)
{
  pdl_enter();
  const size_t total = (size_t)g.Cout * g.KH * g.KW * g.Cin;
  const int npix = g.N * g.Ho * g.Wo;
  const int p0 = blockIdx.y * pix_per_slice;
  const int p1 = p0 + pix_per_slice < npix ? p0 + pix_per_slice : npix;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int ci = (int)(e % g.Cin);
    size_t r = e / g.Cin;
    int kw = (int)(r % g.KW); r /= g.KW;
    int kh = (int)(r % g.KH);
    int co = (int)(r / g.KH);
    float s = 0.f;
    for (int p = p0; p < p1; p++) {
      int n = p / (g.Ho * g.Wo), rem = p - n * g.Ho * g.Wo;
      int oh = rem / g.Wo, ow = rem - oh * g.Wo;
      int ih = oh * g.stride - g.pad + kh, iw = ow * g.stride - g.pad + kw;
      if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
      s = fmaf(__ldg(dy + (size_t)p * g.Cout + co), __ldg(x + ((size_t)(n * g.H + ih) * g.W + iw) * g.Cin + ci), s);
    }
    dst[(size_t)blockIdx.y * total + e] = s;
  }
}

static int scalar_wgrad_slices(const ConvGeom &g)
{
  int npix = g.N * g.Ho * g.Wo;
  int slices = npix / 64;
  if (slices > 64) slices = 64;
  if (slices < 1) slices = 1;
  return slices;
}

// split-K second stage: out = epilogue(sum_z partial[z]) in fixed z order (deterministic)
__global__ void splitk_reduce_kernel(const float *__restrict__ partial, float *__restrict__ out, int M, int Nn, int splits, Epilogue epi)
{
  pdl_enter();
  size_t total = (size_t)M * Nn;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; z++) s += partial[(size_t)z * total + e];
    int m = (int)(e / Nn), n = (int)(e - (size_t)m * Nn);
    out[e] = epilogue_one(s, m, n, Nn, epi);
  }
}

int launch_splitk_reduce(const float *partial, float *out, int M, int Nn, int splits, const Epilogue &epi, cudaStream_t st)
{
  size_t total = (size_t)M * Nn;
  launch(splitk_reduce_kernel, elementwise_grid(total, 256), 256, 0, st, partial, out, M, Nn, splits, epi);
  FRCNN_CHECK_LAUNCH("splitk_reduce_kernel");
  return FRCNN_OK;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Plan {
  int BN;        // 128 or 64
  int splits;
  int chunks_per_split;
  int grid_m, grid_n;
};

static Plan make_plan(int M, int Nn, int K)
{
  Plan p;
  p.BN = Nn <= 64 ? 64 : 128;
  p.grid_m = ceil_div(M, 128);
  p.grid_n = ceil_div(Nn, p.BN);
  int chunks = ceil_div(K, BK);
  int tiles = p.grid_m * p.grid_n;
  int splits = 1;
  if (tiles < kNumSMs) {
    splits = ceil_div(2 * kNumSMs, tiles);
    int max_splits = chunks / 8;             // at least 8 chunks (128 k) per split
    if (splits > max_splits) splits = max_splits;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
  }
  p.chunks_per_split = ceil_div(chunks, splits);
  p.splits = ceil_div(chunks, p.chunks_per_split);
  return p;
}

static bool geom_ok(const ConvGeom &g)
{
  return g.N > 0 && g.H > 0 && g.W > 0 && g.Cin > 0 && g.Cout > 0 && g.KH > 0 && g.KW > 0 && g.stride > 0 && g.pad >= 0 && g.Ho > 0 && g.Wo > 0;
}

static ConvGeom make_geom(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad)
{
  ConvGeom g{N, H, W, Cin, Cout, KH, KW, stride, pad, 0, 0};
  if (stride > 0) {
    g.Ho = (H + 2 * pad - KH) / stride + 1;
    g.Wo = (W + 2 * pad - KW) / stride + 1;
  }
  return g;
}

template <int MODE>
static int launch_igemm(const float *src, const float *wgt, float *dst, const Epilogue &epi, const ConvGeom &g,
                        int M, int Nn, int K, int C, void *workspace, size_t workspace_bytes, cudaStream_t st)
{
  Plan p = make_plan(M, Nn, K);
  float *target = dst;
  if (p.splits > 1) {
    size_t need = (size_t)p.splits * M * Nn * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) return fail(FRCNN_E_WORKSPACE, "conv2d: workspace too small for split-K partials");
    target = reinterpret_cast<float *>(workspace);
  }
  dim3 grid(p.grid_m, p.grid_n, p.splits);
  bool vec = (C % 4 == 0) && (K % 4 == 0);
  if (MODE == MODE_DGRAD) vec = vec && (g.Cin % 4 == 0);   // N-major weight rows are read as float4 along Cin
  Epilogue kernel_epi = epi;
#define LAUNCH(BNV, VECV) launch(igemm_kernel<MODE, 128, BNV, VECV>, grid, 256, 0, st, src, wgt, target, kernel_epi, g, M, Nn, K, C, p.chunks_per_split, p.splits)
  if (p.BN == 128) { if (vec) LAUNCH(128, true); else LAUNCH(128, false); }
  else { if (vec) LAUNCH(64, true); else LAUNCH(64, false); }
#undef LAUNCH
  FRCNN_CHECK_LAUNCH("igemm_kernel");
  if (p.splits > 1) {
    size_t total = (size_t)M * Nn;
    launch(splitk_reduce_kernel, elementwise_grid(total, 256), 256, 0, st, target, dst, M, Nn, p.splits, epi);
    FRCNN_CHECK_LAUNCH("splitk_reduce_kernel");
  }
  return FRCNN_OK;
}

size_t simt_fwd_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad)
{
  ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
  if (!geom_ok(g)) return 0;
  int M = N * g.Ho * g.Wo;
  Plan p = make_plan(M, Cout, KH * KW * Cin);
  return p.splits > 1 ? (size_t)p.splits * M * Cout * sizeof(float) : 0;
}

size_t simt_dgrad_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad)
{
  ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
  if (!geom_ok(g)) return 0;
  int M = N * H * W;
  Plan p = make_plan(M, Cin, KH * KW * Cout);
  return p.splits > 1 ? (size_t)p.splits * M * Cin * sizeof(float) : 0;
}

static Plan make_wgrad_plan(const ConvGeom &g)
{
  return make_plan(g.Cout, g.KH * g.KW * g.Cin, g.N * g.Ho * g.Wo);
}

size_t simt_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad)
{
  ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
  if (!geom_ok(g)) return 0;
  if ((Cin % 4) != 0 || (Cout % 4) != 0) {
    int slices = scalar_wgrad_slices(g);
    return slices > 1 ? (size_t)slices * Cout * KH * KW * Cin * sizeof(float) : 0;
  }
  Plan p = make_wgrad_plan(g);
  return p.splits > 1 ? (size_t)p.splits * Cout * KH * KW * Cin * sizeof(float) : 0;
}


// ---- RGB stem: 3x3 stride-1 "same" convolution with Cin = 3 (vgg16.py:32 block1_conv1) ----------------------------------
// K = 27 is too short for either GEMM engine (one or two k-steps per tile, all prologue/epilogue), so the stem gets a direct
// kernel: a CTA owns a 16 x 8 pixel patch and 64 output channels; the haloed input patch (10 x 18 x 3 floats) and the 27 x 64
// filter slice sit in shared memory.  A half-warp owns 8 consecutive pixels of one row, lane = channel quad, so every thread
// carries 8 x 4 accumulators, reads its input values as shared-memory broadcasts and its weights as conflict-free 128-bit
// loads (39 loads per 288 FMAs), and a pixel's 64 channels leave as one 256-byte row (full 128-bit coalescing).
constexpr int kStemTW = 16, kStemTH = 8;

__global__ void __launch_bounds__(256, 2)
stem3x3_kernel(const float *__restrict__ x, const float *__restrict__ w, float *__restrict__ y, Epilogue epi, int N, int H, int W, int Cout)
{
  pdl_enter();
  __shared__ float xs[kStemTH + 2][(kStemTW + 2) * 3];
  __shared__ __align__(16) float ws[27][64];
  const int tiles_w = (W + kStemTW - 1) / kStemTW, tiles_h = (H + kStemTH - 1) / kStemTH;
  int tile = blockIdx.x;
  const int n = tile / (tiles_w * tiles_h);
  tile -= n * tiles_w * tiles_h;
  const int oh0 = (tile / tiles_w) * kStemTH, ow0 = (tile % tiles_w) * kStemTW;
  const int co0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  for (int e = t; e < (kStemTH + 2) * (kStemTW + 2) * 3; e += 256) {
    const int r = e / ((kStemTW + 2) * 3), q = e - r * (kStemTW + 2) * 3;
    const int ih = oh0 + r - 1, iw = ow0 + q / 3 - 1;
    xs[r][q] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(x + (((size_t)n * H + ih) * W + iw) * 3 + q % 3) : 0.f;
  }
  for (int e = t; e < 27 * 64; e += 256) {
    const int co = e / 27, k = e - co * 27;                           // OHWI: 27 contiguous taps per output channel
    ws[k][co] = __ldg(w + (size_t)(co0 + co) * 27 + k);
  }
  __syncthreads();
  const int half = t >> 4, cq = t & 15;                               // 16 half-warps: row = half / 2, 8-pixel segment = half % 2
  const int row = half >> 1, px0 = (half & 1) * 8;
  float acc[8][4];
#pragma unroll
  for (int p = 0; p < 8; p++) acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; kh++) {
    float xv[30];
#pragma unroll
    for (int i = 0; i < 30; i++) xv[i] = xs[row + kh][px0 * 3 + i];   // 10 pixels x 3 channels, broadcast within the half-warp
#pragma unroll
    for (int kw = 0; kw < 3; kw++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 wv = *reinterpret_cast<const float4 *>(&ws[(kh * 3 + kw) * 3 + c][cq * 4]);
#pragma unroll
        for (int p = 0; p < 8; p++) {
          const float v = xv[(p + kw) * 3 + c];
          acc[p][0] = fmaf(v, wv.x, acc[p][0]);
          acc[p][1] = fmaf(v, wv.y, acc[p][1]);
          acc[p][2] = fmaf(v, wv.z, acc[p][2]);
          acc[p][3] = fmaf(v, wv.w, acc[p][3]);
        }
      }
  }
  const int oh = oh0 + row;
  if (oh >= H) return;
  const int c = co0 + cq * 4;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bi = make_float4(0.f, 0.f, 0.f, 0.f);
  if (epi.scale) sc = __ldg(reinterpret_cast<const float4 *>(epi.scale + c));
  if (epi.bias) bi = __ldg(reinterpret_cast<const float4 *>(epi.bias + c));
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int ow = ow0 + px0 + p;
    if (ow >= W) break;
    float4 o = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
    if (epi.scale) { o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w; }
    o.x = apply_act(o.x + bi.x, epi.act); o.y = apply_act(o.y + bi.y, epi.act);
    o.z = apply_act(o.z + bi.z, epi.act); o.w = apply_act(o.w + bi.w, epi.act);
    *reinterpret_cast<float4 *>(y + (((size_t)n * H + oh) * W + ow) * Cout + c) = o;
  }
}

int simt_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual, float *y,
                    int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                    void *workspace, size_t workspace_bytes, cudaStream_t st)
{
  ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
  if (!geom_ok(g)) return fail(FRCNN_E_BADARG, "conv2d_fwd: bad geometry");
  Epilogue epi{scale, bias, residual, act};
  if (Cin == 3 && KH == 3 && KW == 3 && stride == 1 && pad == 1 && Cout % 64 == 0 && residual == nullptr) {
    const int tiles = N * ceil_div(H, kStemTH) * ceil_div(W, kStemTW);
    launch(stem3x3_kernel, dim3(tiles, Cout / 64), 256, 0, st, x, w, y, epi, N, H, W, Cout);
    FRCNN_CHECK_LAUNCH("stem3x3_kernel");
    return FRCNN_OK;
  }
  return launch_igemm<MODE_FWD>(x, w, y, epi, g, N * g.Ho * g.Wo, Cout, KH * KW * Cin, Cin, workspace, workspace_bytes, st);
}

int simt_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx,
                      int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                      void *workspace, size_t workspace_bytes, cudaStream_t st)
{
  ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
  if (!geom_ok(g)) return fail(FRCNN_E_BADARG, "conv2d_dgrad: bad geometry");
  Epilogue epi{nullptr, nullptr, addend, FRCNN_ACT_NONE};
  return launch_igemm<MODE_DGRAD>(dy, w, dx, epi, g, N * H * W, Cin, KH * KW * Cout, Cout, workspace, workspace_bytes, st);
}

int simt_conv2d_wgrad(const float *dy, const float *x, float *dw,
                      int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                      void *workspace, size_t workspace_bytes, cudaStream_t st)
{
  ConvGeom g = make_geom(N, H, W, Cin, Cout, KH, KW, stride, pad);
  if (!geom_ok(g)) return fail(FRCNN_E_BADARG, "conv2d_wgrad: bad geometry");
  if ((Cin % 4) != 0 || (Cout % 4) != 0) {
    size_t total = (size_t)Cout * KH * KW * Cin;
    int slices = scalar_wgrad_slices(g);
    int npix = N * g.Ho * g.Wo;
    int per = ceil_div(npix, slices);
    slices = ceil_div(npix, per);
    float *target = dw;
    if (slices > 1) {
      if (workspace == nullptr || workspace_bytes < (size_t)slices * total * sizeof(float)) return fail(FRCNN_E_WORKSPACE, "conv2d_wgrad: workspace too small for the K slices");
      target = reinterpret_cast<float *>(workspace);
    }
    launch(wgrad_scalar_kernel, dim3(elementwise_grid(total, 128, 2), slices), 128, 0, st, dy, x, target, g, per);
    FRCNN_CHECK_LAUNCH("wgrad_scalar_kernel");
    if (slices > 1) {
      Epilogue none{nullptr, nullptr, nullptr, FRCNN_ACT_NONE};
      return launch_splitk_reduce(target, dw, Cout, KH * KW * Cin, slices, none, st);
    }
    return FRCNN_OK;
  }
  const int M = Cout, Nn = KH * KW * Cin, Kpix = N * g.Ho * g.Wo;
  Plan p = make_wgrad_plan(g);
  float *target = dw;
  if (p.splits > 1) {
    size_t need = (size_t)p.splits * M * Nn * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) return fail(FRCNN_E_WORKSPACE, "conv2d_wgrad: workspace too small for split-K partials");
    target = reinterpret_cast<float *>(workspace);
  }
  dim3 grid(p.grid_m, p.grid_n, p.splits);
  if (p.BN == 128) launch(wgrad_kernel<128, 128>, grid, 256, 0, st, dy, x, target, g, M, Nn, Kpix, p.chunks_per_split, p.splits);
  else launch(wgrad_kernel<128, 64>, grid, 256, 0, st, dy, x, target, g, M, Nn, Kpix, p.chunks_per_split, p.splits);
  FRCNN_CHECK_LAUNCH("wgrad_kernel");
  if (p.splits > 1) {
    Epilogue none{nullptr, nullptr, nullptr, FRCNN_ACT_NONE};
    size_t total = (size_t)M * Nn;
    launch(splitk_reduce_kernel, elementwise_grid(total, 256), 256, 0, st, target, dw, M, Nn, p.splits, none);
    FRCNN_CHECK_LAUNCH("splitk_reduce_kernel");
  }
  return FRCNN_OK;
}

}  // namespace frcnn
