// HBM-bound elementwise / pooling / optimizer kernels.  All are grid-stride with grids sized in
// multiples of the SM count, 128-bit accesses where the channel count allows, read-only loads
// through the non-coherent path.
#include "common.cuh"
#include "f16_split.cuh"

namespace frcnn {

// ---- layout: (N, C, HW) <-> (N, HW, C) through a 32x33 shared tile -------------------------
__global__ void transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols, int tiles_r, int tiles_c, int batch)
{
  pdl_enter();
  // src: (batch, rows, cols) -> dst: (batch, cols, rows)
  __shared__ float tile[32][33];
  const int total = batch * tiles_r * tiles_c;
  for (int tidx = blockIdx.x; tidx < total; tidx += gridDim.x) {
    int b = tidx / (tiles_r * tiles_c);
    int rem = tidx - b * tiles_r * tiles_c;
    int tr = rem / tiles_c, tc = rem - tr * tiles_c;
    const float *s = src + (size_t)b * rows * cols;
    float *d = dst + (size_t)b * rows * cols;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int r = tr * 32 + i, c = tc * 32 + threadIdx.x;
      tile[i][threadIdx.x] = (r < rows && c < cols) ? __ldg(s + (size_t)r * cols + c) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = tc * 32 + i, r = tr * 32 + threadIdx.x;
      if (r < rows && c < cols) d[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
    __syncthreads();
  }
}

static int launch_transpose(const float *src, float *dst, int batch, int rows, int cols, cudaStream_t st)
{
  int tiles_r = ceil_div(rows, 32), tiles_c = ceil_div(cols, 32);
  long long total = (long long)batch * tiles_r * tiles_c;
  int grid = (int)(total < (long long)kNumSMs * 16 ? total : (long long)kNumSMs * 16);
  if (grid < 1) grid = 1;
  launch(transpose_kernel, grid, dim3(32, 8), 0, st, src, dst, rows, cols, tiles_r, tiles_c, batch);
  FRCNN_CHECK_LAUNCH("transpose_kernel");
  return FRCNN_OK;
}

// ---- relu backward / add / sgd ---------------------------------------------------------------
// Backward of y = relu(conv * s[c] + shift[c] (+ residual)) (frozen BatchNorm as a per-channel epilogue factor, models/resnet.py:56-77):
// dz = dy * (y > 0) is the gradient of the residual branch, dzs = dz * s[c] is what the convolution's dgrad / wgrad consume.  One pass
// over (rows, C) row-major data; y == nullptr: no activation mask; dz == nullptr: only dzs is wanted.  C % 4 == 0.
__global__ void act_bwd_scale_kernel(const float *__restrict__ dy, const float *__restrict__ y, const float *__restrict__ scale,
                                     float *__restrict__ dz, float *__restrict__ dzs, size_t count4, int C4)
{
  pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count4; i += stride) {
    float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + i);
    if (y) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(y) + i);
      g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f;
      g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
    }
    if (dz) reinterpret_cast<float4 *>(dz)[i] = g;
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + (i % C4));
    reinterpret_cast<float4 *>(dzs)[i] = make_float4(__fmul_rn(g.x, sc.x), __fmul_rn(g.y, sc.y), __fmul_rn(g.z, sc.z), __fmul_rn(g.w, sc.w));
  }
}

__global__ void relu_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ y, float *__restrict__ dz, size_t count)
{
  pdl_enter();
  size_t n4 = count / 4;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + i);
    float4 v = __ldg(reinterpret_cast<const float4 *>(y) + i);
    g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f;
    g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
    reinterpret_cast<float4 *>(dz)[i] = g;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride)
    dz[i] = y[i] > 0.f ? dy[i] : 0.f;
}

__global__ void add_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, size_t count)
{
  pdl_enter();
  size_t n4 = count / 4;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 u = __ldg(reinterpret_cast<const float4 *>(a) + i);
    float4 v = __ldg(reinterpret_cast<const float4 *>(b) + i);
    reinterpret_cast<float4 *>(out)[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride) out[i] = a[i] + b[i];
}

__device__ __forceinline__ float fused_tf32_rna(float x)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// torch.optim.SGD (momentum, dampening 0, no nesterov, L2 weight decay folded into the gradient):
// separate mul/add roundings as torch's foreach kernels produce them (no FMA contraction).
__device__ __forceinline__ float sgd_one(float p, float g, float &buf, float lr, float mom, float wd, float gs, int first)
{
  float gg = __fmul_rn(g, gs);
  if (wd != 0.f) gg = __fadd_rn(gg, __fmul_rn(wd, p));
  float b = first ? gg : __fadd_rn(__fmul_rn(buf, mom), gg);
  buf = b;
  return __fadd_rn(p, __fmul_rn(-lr, b));
}

// hi / lo (optional): the updated weights' tf32 operand split for the next step's tcgen05 GEMMs, written in the same pass
// hi16 / lo16 / e16 (optional): the fp16 engine's split of the updated weights instead, with the exponent the buffer already holds
__global__ void sgd_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ buf, size_t count,
                           float lr, float mom, float wd, float gs, int first, float *__restrict__ hi, float *__restrict__ lo,
                           __half *__restrict__ hi16, __half *__restrict__ lo16, const int *__restrict__ e16)
{
  pdl_enter();
  const float s16 = hi16 ? pow2i(__ldg(e16)) : 1.0f;
  size_t n4 = count / 4;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4 *>(p)[i];
    float4 gv = __ldg(reinterpret_cast<const float4 *>(g) + i);
    float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4 *>(buf)[i];
    pv.x = sgd_one(pv.x, gv.x, bv.x, lr, mom, wd, gs, first);
    pv.y = sgd_one(pv.y, gv.y, bv.y, lr, mom, wd, gs, first);
    pv.z = sgd_one(pv.z, gv.z, bv.z, lr, mom, wd, gs, first);
    pv.w = sgd_one(pv.w, gv.w, bv.w, lr, mom, wd, gs, first);
    reinterpret_cast<float4 *>(p)[i] = pv;
    reinterpret_cast<float4 *>(buf)[i] = bv;
    if (hi) {
      float4 h, l;
      h.x = fused_tf32_rna(pv.x); h.y = fused_tf32_rna(pv.y); h.z = fused_tf32_rna(pv.z); h.w = fused_tf32_rna(pv.w);
      l.x = pv.x - h.x; l.y = pv.y - h.y; l.z = pv.z - h.z; l.w = pv.w - h.w;
      reinterpret_cast<float4 *>(hi)[i] = h;
      reinterpret_cast<float4 *>(lo)[i] = l;
    }
    if (hi16) {
      uint2 h, l;
      split16x4(pv, s16, h, l);
      reinterpret_cast<uint2 *>(hi16)[i] = h;
      reinterpret_cast<uint2 *>(lo16)[i] = l;
    }
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride) {
    float b = first ? 0.f : buf[i];
    const float v = sgd_one(p[i], g[i], b, lr, mom, wd, gs, first);
    p[i] = v;
    buf[i] = b;
    if (hi) { const float h = fused_tf32_rna(v); hi[i] = h; lo[i] = v - h; }
    if (hi16) split16(v, s16, hi16[i], lo16[i]);
  }
}

// out[r][c] = x[r][c] * scale[r]  (per-row scale: folds a frozen BatchNorm's gamma/sqrt(var+eps) into the
// filter rows, and un-folds it from the filter gradient)
__global__ void scale_rows_kernel(const float *__restrict__ x, const float *__restrict__ scale, float *__restrict__ out, size_t rows, size_t row_len)
{
  pdl_enter();
  size_t total = rows * row_len;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
    out[e] = x[e] * __ldg(scale + e / row_len);
}

// ---- bias gradient: column sums of (rows, C), two deterministic stages ----------------------
constexpr int kBiasRowsPerBlock = 64;

__global__ void bias_grad_stage1(const float *__restrict__ dz, float *__restrict__ partial, size_t rows, int C)
{
  pdl_enter();
  // block b sums rows [b*64, b*64+64) for every channel; 4 independent row streams per thread keep
  // enough loads in flight, combined in a fixed order (deterministic)
  size_t r0 = (size_t)blockIdx.x * kBiasRowsPerBlock;
  size_t r1 = r0 + kBiasRowsPerBlock < rows ? r0 + kBiasRowsPerBlock : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    size_t r = r0;
    for (; r + 4 <= r1; r += 4) {
      s0 += __ldg(dz + r * C + c);
      s1 += __ldg(dz + (r + 1) * C + c);
      s2 += __ldg(dz + (r + 2) * C + c);
      s3 += __ldg(dz + (r + 3) * C + c);
    }
    for (; r < r1; r++) s0 += __ldg(dz + r * C + c);
    partial[(size_t)blockIdx.x * C + c] = (s0 + s1) + (s2 + s3);
  }
}

__global__ void bias_grad_stage2(const float *__restrict__ partial, float *__restrict__ dbias, int blocks, int C)
{
  pdl_enter();
  // one warp per channel group of 32: lanes = channels (coalesced), 8 row-slices per CTA combined through shared memory
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < C)
    for (int b = slice; b < blocks; b += 8) s += partial[(size_t)b * C + c];
  red[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && c < C) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) tot += red[k][lane];
    dbias[c] = tot;
  }
}


// ---- fused activation backward: dz (+ its tf32 hi/lo split) (+ per-channel sums for the bias gradient) in one pass -------
// Replaces relu_bwd -> split_hi_lo -> bias_grad_stage1 (three passes over dz) for the layers whose gradient GEMMs run on the
// tcgen05 engine.  A CTA owns a slab of <= 1024 channels (blockIdx.y) and a block of rows (blockIdx.x); threads are laid out
// [row lane][float4 channel group] so every access is a full 128-bit coalesced row segment; column sums are kept per thread
// and combined over the row lanes through shared memory in a fixed order (deterministic), then reduced by bias_grad_stage2.

template <int MODE>     // 0: dz = dy   1: dz = y > 0 ? dy : 0
__global__ void __launch_bounds__(256)
act_bwd_fused_kernel(const float *__restrict__ dy, const float *__restrict__ y, float *__restrict__ dz, float *__restrict__ hi, float *__restrict__ lo,
                     float *__restrict__ partial, size_t rows, int C, int slab_c, int rows_per_block,
                     unsigned *__restrict__ hdr16, const unsigned *__restrict__ partials16, int G16, __half *__restrict__ hi16, __half *__restrict__ lo16)
{
  pdl_enter();
  __shared__ float4 red[256];
  // fp16 engine: exponent from the partial maxima of |dy| (the amax pass ran just before; |dz| <= |dy| under the ReLU mask)
  float s16 = 1.0f;
  if (hi16) {
    const unsigned amax = f16_reduce_partials(partials16, G16, reinterpret_cast<unsigned *>(red));
    const int e = f16_exponent(amax);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { hdr16[0] = amax; hdr16[1] = (unsigned)e; }
    s16 = pow2i(e);
  }
  const int groups = slab_c / 4;                       // float4 groups per row inside this CTA's channel slab (divides 256)
  const int lanes = 256 / groups;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  const size_t row_f4 = (size_t)C / 4;
  const size_t col_f4 = (size_t)blockIdx.y * groups + cg;
  const size_t r0 = (size_t)blockIdx.x * rows_per_block;
  const size_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t r = r0 + rl; r < r1; r += lanes) {
    const size_t e = r * row_f4 + col_f4;
    float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + e);
    if (MODE == 1) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(y) + e);
      g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f;
      g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
    }
    if (dz) reinterpret_cast<float4 *>(dz)[e] = g;
    if (hi) {
      float4 h, l;
      h.x = fused_tf32_rna(g.x); h.y = fused_tf32_rna(g.y); h.z = fused_tf32_rna(g.z); h.w = fused_tf32_rna(g.w);
      l.x = g.x - h.x; l.y = g.y - h.y; l.z = g.z - h.z; l.w = g.w - h.w;
      reinterpret_cast<float4 *>(hi)[e] = h;
      reinterpret_cast<float4 *>(lo)[e] = l;
    }
    if (hi16) {
      uint2 h, l;
      split16x4(g, s16, h, l);
      reinterpret_cast<uint2 *>(hi16)[e] = h;
      reinterpret_cast<uint2 *>(lo16)[e] = l;
    }
    s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
  }
  if (partial) {
    red[threadIdx.x] = s;
    __syncthreads();
    if (rl == 0) {
      float4 tot = red[cg];
      for (int k = 1; k < lanes; k++) {
        const float4 o = red[k * groups + cg];
        tot.x += o.x; tot.y += o.y; tot.z += o.z; tot.w += o.w;
      }
      reinterpret_cast<float4 *>(partial + (size_t)blockIdx.x * C)[col_f4] = tot;
    }
  }
}

static bool act_bwd_fused_plan(size_t rows, int C, int *slab_c, int *blocks_x, int *rows_per_block)
{
  if (C < 64 || rows == 0) return false;
  int sc = C >= 1024 ? 1024 : C;
  if (C % sc != 0 || (sc & (sc - 1)) != 0) return false;          // slabs of 64..1024 channels, power of two (4 * divisor of 256)
  const int lanes = 256 / (sc / 4);
  const int slabs = C / sc;
  size_t want = ceil_div<size_t>(rows, (size_t)lanes);
  size_t cap = (size_t)(4 * kNumSMs / slabs > 1 ? 4 * kNumSMs / slabs : 1);
  if (want > cap) want = cap;
  size_t per = ceil_div<size_t>(rows, want);
  per = ceil_div<size_t>(per, (size_t)lanes) * lanes;
  *slab_c = sc;
  *rows_per_block = (int)per;
  *blocks_x = (int)ceil_div<size_t>(rows, per);
  return true;
}

// ---- pooling ---------------------------------------------------------------------------------
template <int VEC>
__global__ void maxpool2x2_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C)
{
  pdl_enter();
  const int Ho = H / 2, Wo = W / 2, Cv = C / VEC;
  size_t total = (size_t)N * Ho * Wo * Cv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int cv = (int)(e % Cv);
    size_t r = e / Cv;
    int ow = (int)(r % Wo); r /= Wo;
    int oh = (int)(r % Ho);
    int n = (int)(r / Ho);
    const float *base = x + (((size_t)n * H + 2 * oh) * W + 2 * ow) * C + cv * VEC;
    if (VEC == 4) {
      float4 a = __ldg(reinterpret_cast<const float4 *>(base));
      float4 b = __ldg(reinterpret_cast<const float4 *>(base + C));
      float4 c = __ldg(reinterpret_cast<const float4 *>(base + (size_t)W * C));
      float4 d = __ldg(reinterpret_cast<const float4 *>(base + (size_t)W * C + C));
      float4 o;
      o.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)); o.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
      o.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)); o.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
      *reinterpret_cast<float4 *>(y + (((size_t)n * Ho + oh) * Wo + ow) * C + cv * 4) = o;
    } else {
      float o = fmaxf(fmaxf(__ldg(base), __ldg(base + C)), fmaxf(__ldg(base + (size_t)W * C), __ldg(base + (size_t)W * C + C)));
      y[(((size_t)n * Ho + oh) * Wo + ow) * C + cv] = o;
    }
  }
}

// first maximum of the 2x2 window in (kh,kw) scan order with strict '>' (torch max_pool2d),
// gradient passes only where the pre-pool activation is > 0 (ReLU backward of the producer).
__device__ __forceinline__ float pool_relu_grad(float v00, float v01, float v10, float v11, int pos, float g)
{
  int best = 0; float bv = v00;
  if (v01 > bv) { bv = v01; best = 1; }
  if (v10 > bv) { bv = v10; best = 2; }
  if (v11 > bv) { bv = v11; best = 3; }
  return (best == pos && bv > 0.f) ? g : 0.f;
}

__global__ void maxpool2x2_relu_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, float *__restrict__ dz, int N, int H, int W, int C)
{
  pdl_enter();
  // one thread per (n, window-or-edge cell, channel); windows cover rows/cols < 2*Ho / 2*Wo, the
  // odd trailing row/column (floor mode) receives zero gradient.
  const int Ho = H / 2, Wo = W / 2;
  const int Hc = (H + 1) / 2, Wc = (W + 1) / 2;
  size_t total = (size_t)N * Hc * Wc * C;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    size_t r = e / C;
    int cw = (int)(r % Wc); r /= Wc;
    int ch = (int)(r % Hc);
    int n = (int)(r / Hc);
    int h0 = 2 * ch, w0 = 2 * cw;
    size_t base = (((size_t)n * H + h0) * W + w0) * C + c;
    if (ch < Ho && cw < Wo) {
      float v00 = __ldg(x + base), v01 = __ldg(x + base + C);
      float v10 = __ldg(x + base + (size_t)W * C), v11 = __ldg(x + base + (size_t)W * C + C);
      float g = __ldg(dy + (((size_t)n * Ho + ch) * Wo + cw) * C + c);
      dz[base] = pool_relu_grad(v00, v01, v10, v11, 0, g);
      dz[base + C] = pool_relu_grad(v00, v01, v10, v11, 1, g);
      dz[base + (size_t)W * C] = pool_relu_grad(v00, v01, v10, v11, 2, g);
      dz[base + (size_t)W * C + C] = pool_relu_grad(v00, v01, v10, v11, 3, g);
    } else {
      // edge cells outside every window
      dz[base] = 0.f;
      if (w0 + 1 < W) dz[base + C] = 0.f;
      if (h0 + 1 < H) {
        dz[base + (size_t)W * C] = 0.f;
        if (w0 + 1 < W) dz[base + (size_t)W * C + C] = 0.f;
      }
    }
  }
}

__global__ void maxpool3x3s2_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C, int Ho, int Wo)
{
  pdl_enter();
  size_t total = (size_t)N * Ho * Wo * C;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    size_t r = e / C;
    int ow = (int)(r % Wo); r /= Wo;
    int oh = (int)(r % Ho);
    int n = (int)(r / Ho);
    float best = -INFINITY;
    for (int kh = 0; kh < 3; kh++) {
      int ih = oh * 2 - 1 + kh;
      if (ih < 0 || ih >= H) continue;
      for (int kw = 0; kw < 3; kw++) {
        int iw = ow * 2 - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        best = fmaxf(best, __ldg(x + (((size_t)n * H + ih) * W + iw) * C + c));
      }
    }
    y[e] = best;
  }
}

// ---- stride-2 helpers (models/resnet.py:79-81,109-118: the 3x3 / 1x1 stride-2 convolutions of layer2-4) -----------------------------
// A stride-2 convolution reads / produces every other pixel.  The tcgen05 engine is stride-1, so resnet.py runs
//   1x1 s2:  conv1x1_s1(subsample2(x));   3x3 s2 pad 1:  subsample2(conv3x3_s1(x))
// and, backwards, the stride-1 dgrad / wgrad on upsample2_zero(dy) -- the same sums of the same products as the strided form
// (zeros contribute exact zeros).  V = floats per thread (4 when C % 4 == 0).
template <int V>
__global__ void subsample2_kernel(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C, int Ho, int Wo)
{
  pdl_enter();
  const int Cv = C / V;
  const size_t total = (size_t)N * Ho * Wo * Cv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cv);
    size_t r = e / Cv;
    const int ow = (int)(r % Wo); r /= Wo;
    const int oh = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const size_t src = ((((size_t)n * H + 2 * oh) * W + 2 * ow) * Cv + c) * V;
    if (V == 4) *reinterpret_cast<float4 *>(y + e * 4) = __ldg(reinterpret_cast<const float4 *>(x + src));
    else y[e] = __ldg(x + src);
  }
}

// y (N, H, W, C) <- x (N, Ho, Wo, C): y[n, 2i, 2j, :] = x[n, i, j, :], every other pixel zero
template <int V>
__global__ void upsample2_zero_kernel(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C, int Ho, int Wo)
{
  pdl_enter();
  const int Cv = C / V;
  const size_t total = (size_t)N * H * W * Cv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cv);
    size_t r = e / Cv;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const bool live = !((h | w) & 1);
    const size_t src = ((((size_t)n * Ho + (h >> 1)) * Wo + (w >> 1)) * Cv + c) * V;
    if (V == 4) *reinterpret_cast<float4 *>(y + e * 4) = live ? __ldg(reinterpret_cast<const float4 *>(x + src)) : make_float4(0.f, 0.f, 0.f, 0.f);
    else y[e] = live ? __ldg(x + src) : 0.f;
  }
}

__global__ void spatial_mean_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int N, int H, int W, int C)
{
  pdl_enter();
  // y.mean(-1).mean(-1) (models/resnet.py:117): mean over W for every row, then mean of the row means
  size_t total = (size_t)N * C;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    int n = (int)(e / C);
    float acc = 0.f;
    for (int h = 0; h < H; h++) {
      float s = 0.f;
      for (int w = 0; w < W; w++) s += __ldg(x + (((size_t)n * H + h) * W + w) * C + c);
      acc += s / (float)W;
    }
    y[e] = acc / (float)H;
  }
}

__global__ void spatial_mean_bwd_kernel(const float *__restrict__ dy, float *__restrict__ dx, int N, int HW, int C)
{
  pdl_enter();
  size_t total = (size_t)N * HW * C;
  const float inv = 1.0f / (float)HW;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    int n = (int)(e / ((size_t)HW * C));
    dx[e] = __ldg(dy + (size_t)n * C + c) * inv;
  }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_nchw_to_nhwc(const float *src, float *dst, int N, int C, int H, int W, void *stream)
{
  FRCNN_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad argument");
  return launch_transpose(src, dst, N, C, H * W, as_stream(stream));
}

int frcnn_nhwc_to_nchw(const float *src, float *dst, int N, int C, int H, int W, void *stream)
{
  FRCNN_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad argument");
  return launch_transpose(src, dst, N, H * W, C, as_stream(stream));
}

int frcnn_relu_bwd(const float *dy, const float *y, float *dz, size_t count, void *stream)
{
  FRCNN_REQUIRE(dy && y && dz, "relu_bwd: null pointer");
  if (count == 0) return FRCNN_OK;
  launch(relu_bwd_kernel, elementwise_grid(count / 4 + 1, 256), 256, 0, as_stream(stream), dy, y, dz, count);
  FRCNN_CHECK_LAUNCH("relu_bwd_kernel");
  return FRCNN_OK;
}

int frcnn_act_bwd_scale(const float *dy, const float *y, const float *scale, float *dz, float *dzs, size_t rows, int C, void *stream)
{
  FRCNN_REQUIRE(dy && scale && dzs && C > 0 && C % 4 == 0, "act_bwd_scale: bad argument (C must be a multiple of 4)");
  FRCNN_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(dz) |
                  reinterpret_cast<uintptr_t>(dzs)) & 15) == 0, "act_bwd_scale: pointers must be 16-byte aligned");
  if (rows == 0) return FRCNN_OK;
  const size_t count4 = rows * (size_t)(C / 4);
  launch(act_bwd_scale_kernel, elementwise_grid(count4, 256), 256, 0, as_stream(stream), dy, y, scale, dz, dzs, count4, C / 4);
  FRCNN_CHECK_LAUNCH("act_bwd_scale_kernel");
  return FRCNN_OK;
}

int frcnn_add(const float *a, const float *b, float *out, size_t count, void *stream)
{
  FRCNN_REQUIRE(a && b && out, "add: null pointer");
  if (count == 0) return FRCNN_OK;
  launch(add_kernel, elementwise_grid(count / 4 + 1, 256), 256, 0, as_stream(stream), a, b, out, count);
  FRCNN_CHECK_LAUNCH("add_kernel");
  return FRCNN_OK;
}

// ctas_per_sm: cap of the grid-stride launch (default 8 x 148 CTAs, each looping over its share).  A large value (>= count / 1024 / 148)
// makes the launch non-persistent -- one short-lived CTA per 1024 elements -- which is what a side-stream launch next to the persistent
// GEMM kernels wants: 256 threads x 32 registers fit beside a GEMM CTA, and a short CTA never keeps an SM from the next GEMM launch.
static int sgd_launch(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                      float grad_scale, int first_step, float *hi, float *lo, __half *hi16, __half *lo16, const int *e16, int ctas_per_sm, void *stream)
{
  const size_t want = ceil_div<size_t>(count / 4 + 1, 256);
  size_t cap = (size_t)kNumSMs * (size_t)(ctas_per_sm > 0 ? ctas_per_sm : 8);
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  launch(sgd_kernel, grid, 256, 0, as_stream(stream), param, grad, momentum_buf, count, lr, momentum, weight_decay, grad_scale, first_step, hi, lo, hi16, lo16, e16);
  FRCNN_CHECK_LAUNCH("sgd_kernel");
  return FRCNN_OK;
}

static int sgd_step_split_tf32(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                               float grad_scale, int first_step, void *param_split, int ctas_per_sm, void *stream)
{
  FRCNN_REQUIRE(param && grad && momentum_buf, "sgd_step: null pointer");
  if (count == 0) return FRCNN_OK;
  float *hi = nullptr, *lo = nullptr;
  if (param_split) {
    hi = reinterpret_cast<float *>(param_split);
    lo = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(param_split) + (count * 4 + 1023) / 1024 * 1024);   // frcnn_tf32_split layout
  }
  return sgd_launch(param, grad, momentum_buf, count, lr, momentum, weight_decay, grad_scale, first_step, hi, lo, nullptr, nullptr, nullptr, ctas_per_sm, stream);
}

static int sgd_step_split_f16(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                              float grad_scale, int first_step, void *param_split, int ctas_per_sm, void *stream)
{
  FRCNN_REQUIRE(param && grad && momentum_buf && param_split && count > 0, "sgd_step_split_f16: bad argument");
  FRCNN_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(momentum_buf) | reinterpret_cast<uintptr_t>(param_split)) & 15) == 0,
                "sgd_step_split_f16: pointers must be 16-byte aligned");
  uint8_t *o = reinterpret_cast<uint8_t *>(param_split);
  __half *hi = reinterpret_cast<__half *>(o + kF16Header);
  __half *lo = reinterpret_cast<__half *>(o + kF16Header + f16_half_bytes(count));
  return sgd_launch(param, grad, momentum_buf, count, lr, momentum, weight_decay, grad_scale, first_step, nullptr, nullptr, hi, lo, reinterpret_cast<const int *>(o) + 1, ctas_per_sm, stream);
}

int frcnn_sgd_step_split(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                         float grad_scale, int first_step, void *param_split, void *stream)
{
  return sgd_step_split_tf32(param, grad, momentum_buf, count, lr, momentum, weight_decay, grad_scale, first_step, param_split, 0, stream);
}

int frcnn_sgd_step_split_f16(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                             float grad_scale, int first_step, void *param_split, void *stream)
{
  return sgd_step_split_f16(param, grad, momentum_buf, count, lr, momentum, weight_decay, grad_scale, first_step, param_split, 0, stream);
}

int frcnn_sgd_step_multi_ex(int n, float *const *params, const float *const *grads, float *const *momentum_bufs, const size_t *counts,
                            const float *lrs, const float *momenta, const float *weight_decays, const int *first_steps, void *const *param_splits,
                            int split_format, float grad_scale, int ctas_per_sm, void *stream)
{
  FRCNN_REQUIRE(n >= 0 && (n == 0 || (params && grads && momentum_bufs && counts && lrs && momenta && weight_decays && first_steps)), "sgd_step_multi: bad argument");
  FRCNN_REQUIRE(split_format >= 0 && split_format <= 2, "sgd_step_multi: split_format must be 0 (none), 1 (tf32) or 2 (fp16)");
  for (int i = 0; i < n; i++) {
    void *split = (param_splits && split_format != 0) ? param_splits[i] : nullptr;
    int rc;
    if (split && split_format == 2)
      rc = sgd_step_split_f16(params[i], grads[i], momentum_bufs[i], counts[i], lrs[i], momenta[i], weight_decays[i], grad_scale, first_steps[i], split, ctas_per_sm, stream);
    else
      rc = sgd_step_split_tf32(params[i], grads[i], momentum_bufs[i], counts[i], lrs[i], momenta[i], weight_decays[i], grad_scale, first_steps[i], split, ctas_per_sm, stream);
    if (rc != FRCNN_OK) return rc;
  }
  return FRCNN_OK;
}

int frcnn_sgd_step_multi(int n, float *const *params, const float *const *grads, float *const *momentum_bufs, const size_t *counts,
                         const float *lrs, const float *momenta, const float *weight_decays, const int *first_steps, void *const *param_splits,
                         int split_format, float grad_scale, void *stream)
{
  return frcnn_sgd_step_multi_ex(n, params, grads, momentum_bufs, counts, lrs, momenta, weight_decays, first_steps, param_splits, split_format, grad_scale, 0, stream);
}

int frcnn_sgd_step(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                   float grad_scale, int first_step, void *stream)
{
  return frcnn_sgd_step_split(param, grad, momentum_buf, count, lr, momentum, weight_decay, grad_scale, first_step, nullptr, stream);
}

int frcnn_scale_rows(const float *x, const float *scale, float *out, size_t rows, size_t row_len, void *stream)
{
  FRCNN_REQUIRE(x && scale && out, "scale_rows: null pointer");
  if (rows * row_len == 0) return FRCNN_OK;
  launch(scale_rows_kernel, elementwise_grid(rows * row_len, 256), 256, 0, as_stream(stream), x, scale, out, rows, row_len);
  FRCNN_CHECK_LAUNCH("scale_rows_kernel");
  return FRCNN_OK;
}

size_t frcnn_bias_grad_workspace_bytes(size_t rows, int C)
{
  size_t blocks = ceil_div<size_t>(rows, kBiasRowsPerBlock);
  return blocks * (size_t)C * sizeof(float);
}

int frcnn_bias_grad(const float *dz, float *dbias, size_t rows, int C, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dz && dbias && rows > 0 && C > 0, "bias_grad: bad argument");
  if (workspace == nullptr || workspace_bytes < frcnn_bias_grad_workspace_bytes(rows, C)) return fail(FRCNN_E_WORKSPACE, "bias_grad: workspace too small");
  int blocks = (int)ceil_div<size_t>(rows, kBiasRowsPerBlock);
  float *partial = reinterpret_cast<float *>(workspace);
  launch(bias_grad_stage1, blocks, C < 256 ? ((C + 31) / 32) * 32 : 256, 0, as_stream(stream), dz, partial, rows, C);
  FRCNN_CHECK_LAUNCH("bias_grad_stage1");
  launch(bias_grad_stage2, ceil_div(C, 32), 256, 0, as_stream(stream), partial, dbias, blocks, C);
  FRCNN_CHECK_LAUNCH("bias_grad_stage2");
  return FRCNN_OK;
}

int frcnn_act_bwd_fused_supported(size_t rows, int C)
{
  int a, b, c;
  return act_bwd_fused_plan(rows, C, &a, &b, &c) ? 1 : 0;
}

size_t frcnn_act_bwd_fused_workspace_bytes(size_t rows, int C)
{
  int sc, bx, per;
  if (!act_bwd_fused_plan(rows, C, &sc, &bx, &per)) return 0;
  return (size_t)bx * C * sizeof(float);
}

static int act_bwd_fused_impl(const float *dy, const float *y, int act, float *dz, void *dz_split, float *dbias, size_t rows, int C,
                              void *workspace, size_t workspace_bytes, void *stream, bool f16, const void *dy_amax = nullptr, int dy_amax_slots = 0)
{
  FRCNN_REQUIRE(dy && rows > 0 && C > 0, "act_bwd_fused: bad argument");
  FRCNN_REQUIRE(act == FRCNN_ACT_NONE || (act == FRCNN_ACT_RELU && y), "act_bwd_fused: activation must be NONE or RELU (with y)");
  FRCNN_REQUIRE(dz || dz_split || dbias, "act_bwd_fused: nothing to produce");
  int sc, bx, per;
  if (!act_bwd_fused_plan(rows, C, &sc, &bx, &per)) return fail(FRCNN_E_UNSUPPORTED, "act_bwd_fused: channel count not a multiple of a 64..1024 power-of-two slab");
  float *partial = nullptr;
  if (dbias) {
    if (workspace == nullptr || workspace_bytes < (size_t)bx * C * sizeof(float)) return fail(FRCNN_E_WORKSPACE, "act_bwd_fused: workspace too small");
    partial = reinterpret_cast<float *>(workspace);
  }
  cudaStream_t st = as_stream(stream);
  const size_t count = rows * (size_t)C;
  float *hi = nullptr, *lo = nullptr;
  __half *hi16 = nullptr, *lo16 = nullptr;
  unsigned *hdr16 = nullptr;
  const unsigned *partials16 = nullptr;
  int G16 = 0;
  if (dz_split && !f16) {
    hi = reinterpret_cast<float *>(dz_split);
    lo = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(dz_split) + (count * 4 + 1023) / 1024 * 1024);     // frcnn_tf32_split layout
  } else if (dz_split) {
    uint8_t *o = reinterpret_cast<uint8_t *>(dz_split);                                                             // frcnn_f16_split layout
    FRCNN_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0, "act_bwd_fused_f16: dz_split must be 16-byte aligned");
    hdr16 = reinterpret_cast<unsigned *>(o);
    hi16 = reinterpret_cast<__half *>(o + kF16Header);
    lo16 = reinterpret_cast<__half *>(o + kF16Header + f16_half_bytes(count));
    if (dy_amax) {                                               // partial maxima of dy left behind by the kernel that produced it
      partials16 = reinterpret_cast<const unsigned *>(dy_amax);
      G16 = dy_amax_slots;
    } else {
      G16 = f16_launch_amax(dy, count, o, st);                   // max |dy| bounds max |dz|: the exponent is known before dz is formed
      FRCNN_CHECK_LAUNCH("f16_amax_partials_kernel");
      partials16 = hdr16;
    }
  }
  dim3 grid(bx, C / sc);
  if (act == FRCNN_ACT_RELU) launch(act_bwd_fused_kernel<1>, grid, 256, 0, st, dy, y, dz, hi, lo, partial, rows, C, sc, per, hdr16, partials16, G16, hi16, lo16);
  else launch(act_bwd_fused_kernel<0>, grid, 256, 0, st, dy, y, dz, hi, lo, partial, rows, C, sc, per, hdr16, partials16, G16, hi16, lo16);
  FRCNN_CHECK_LAUNCH("act_bwd_fused_kernel");
  if (dbias) {
    launch(bias_grad_stage2, ceil_div(C, 32), 256, 0, st, partial, dbias, bx, C);
    FRCNN_CHECK_LAUNCH("bias_grad_stage2");
  }
  return FRCNN_OK;
}

int frcnn_act_bwd_fused(const float *dy, const float *y, int act, float *dz, void *dz_split, float *dbias, size_t rows, int C,
                        void *workspace, size_t workspace_bytes, void *stream)
{
  return act_bwd_fused_impl(dy, y, act, dz, dz_split, dbias, rows, C, workspace, workspace_bytes, stream, false);
}

int frcnn_act_bwd_fused_f16(const float *dy, const float *y, int act, float *dz, void *dz_split, float *dbias, size_t rows, int C,
                            const void *dy_amax, int dy_amax_slots, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(dy_amax == nullptr || (dy_amax_slots > 0 && dy_amax_slots <= 1000), "act_bwd_fused_f16: bad amax slot count");
  return act_bwd_fused_impl(dy, y, act, dz, dz_split, dbias, rows, C, workspace, workspace_bytes, stream, true, dy_amax, dy_amax_slots);
}

int frcnn_maxpool2x2_fwd(const float *x, float *y, int N, int H, int W, int C, void *stream)
{
  FRCNN_REQUIRE(x && y && N > 0 && H >= 2 && W >= 2 && C > 0, "maxpool2x2_fwd: bad argument");
  size_t total = (size_t)N * (H / 2) * (W / 2) * C;
  if (C % 4 == 0) launch(maxpool2x2_fwd_kernel<4>, elementwise_grid(total / 4, 256), 256, 0, as_stream(stream), x, y, N, H, W, C);
  else launch(maxpool2x2_fwd_kernel<1>, elementwise_grid(total, 256), 256, 0, as_stream(stream), x, y, N, H, W, C);
  FRCNN_CHECK_LAUNCH("maxpool2x2_fwd_kernel");
  return FRCNN_OK;
}

int frcnn_maxpool2x2_relu_bwd(const float *dy, const float *x, float *dz, int N, int H, int W, int C, void *stream)
{
  FRCNN_REQUIRE(dy && x && dz && N > 0 && H >= 2 && W >= 2 && C > 0, "maxpool2x2_relu_bwd: bad argument");
  size_t total = (size_t)N * ((H + 1) / 2) * ((W + 1) / 2) * C;
  launch(maxpool2x2_relu_bwd_kernel, elementwise_grid(total, 256), 256, 0, as_stream(stream), dy, x, dz, N, H, W, C);
  FRCNN_CHECK_LAUNCH("maxpool2x2_relu_bwd_kernel");
  return FRCNN_OK;
}

int frcnn_maxpool3x3s2_fwd(const float *x, float *y, int N, int H, int W, int C, void *stream)
{
  FRCNN_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "maxpool3x3s2_fwd: bad argument");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  size_t total = (size_t)N * Ho * Wo * C;
  launch(maxpool3x3s2_fwd_kernel, elementwise_grid(total, 256), 256, 0, as_stream(stream), x, y, N, H, W, C, Ho, Wo);
  FRCNN_CHECK_LAUNCH("maxpool3x3s2_fwd_kernel");
  return FRCNN_OK;
}

int frcnn_subsample2(const float *x, float *y, int N, int H, int W, int C, void *stream)
{
  FRCNN_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "subsample2: bad argument");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const bool v4 = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const size_t total = (size_t)N * Ho * Wo * C;
  if (v4) launch(subsample2_kernel<4>, elementwise_grid(total / 4, 256), 256, 0, as_stream(stream), x, y, N, H, W, C, Ho, Wo);
  else launch(subsample2_kernel<1>, elementwise_grid(total, 256), 256, 0, as_stream(stream), x, y, N, H, W, C, Ho, Wo);
  FRCNN_CHECK_LAUNCH("subsample2_kernel");
  return FRCNN_OK;
}

int frcnn_upsample2_zero(const float *x, float *y, int N, int H, int W, int C, void *stream)
{
  FRCNN_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "upsample2_zero: bad argument");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const bool v4 = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const size_t total = (size_t)N * H * W * C;
  if (v4) launch(upsample2_zero_kernel<4>, elementwise_grid(total / 4, 256), 256, 0, as_stream(stream), x, y, N, H, W, C, Ho, Wo);
  else launch(upsample2_zero_kernel<1>, elementwise_grid(total, 256), 256, 0, as_stream(stream), x, y, N, H, W, C, Ho, Wo);
  FRCNN_CHECK_LAUNCH("upsample2_zero_kernel");
  return FRCNN_OK;
}

int frcnn_spatial_mean_fwd(const float *x, float *y, int N, int H, int W, int C, void *stream)
{
  FRCNN_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "spatial_mean_fwd: bad argument");
  launch(spatial_mean_fwd_kernel, elementwise_grid((size_t)N * C, 256), 256, 0, as_stream(stream), x, y, N, H, W, C);
  FRCNN_CHECK_LAUNCH("spatial_mean_fwd_kernel");
  return FRCNN_OK;
}

int frcnn_spatial_mean_bwd(const float *dy, float *dx, int N, int HW, int C, void *stream)
{
  FRCNN_REQUIRE(dy && dx && N > 0 && HW > 0 && C > 0, "spatial_mean_bwd: bad argument");
  launch(spatial_mean_bwd_kernel, elementwise_grid((size_t)N * HW * C, 256), 256, 0, as_stream(stream), dy, dx, N, HW, C);
  FRCNN_CHECK_LAUNCH("spatial_mean_bwd_kernel");
  return FRCNN_OK;
}

}  // extern "C"
