// tcgen05 engine (FRCNN_ENGINE_TC_3XTF32): implicit-GEMM convolution on the 5th-gen tensor cores.
//
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) straight from the NHWC fp32
//     activation: for filter tap (kh,kw) the A tile of a (tile_h x tile_w) output patch is the
//     input patch shifted by (kh-pad, kw-pad); the TMA unit zero-fills the out-of-image part, so
//     im2col and padding are folded into the shared-memory staging and never exist in HBM;
//   * tcgen05.mma kind::tf32 issued by one thread, fp32 accumulators in TMEM (128 lanes x BN cols);
//   * fp32-grade accuracy through the error-compensated split x = hi + lo (hi = the 11-bit tf32
//     truncation the tensor core applies itself, lo = x - hi): D += A_hi*B_lo + A_lo*B_hi + A_hi*B_hi;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 =
//     epilogue (tcgen05.ld -> scale/bias/residual/activation -> 128-bit stores); a STAGES-deep
//     mbarrier ring (full/empty) between producer and issuer, tcgen05.commit frees slots.
#include "common.cuh"
#include "tc_common.cuh"

namespace frcnn {

using namespace tc;

// conv_simt.cu
struct Epilogue {
  const float *scale;
  const float *bias;
  const float *residual;
  int act;
};
int launch_splitk_reduce(const float *partial, float *out, int M, int Nn, int splits, const Epilogue &epi, cudaStream_t st);

// ---- lo = x - tf32_trunc(x) ----------------------------------------------------------------------
__global__ void split_lo_kernel(const float *__restrict__ x, float *__restrict__ lo, size_t count)
{
  size_t n4 = count / 4;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
    float4 r;
    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    reinterpret_cast<float4 *>(lo)[i] = r;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride)
    lo[i] = x[i] - __uint_as_float(__float_as_uint(x[i]) & 0xffffe000u);
}

// ---- kernel ---------------------------------------------------------------------------------------
struct TcGeom {
  int Cin, Cout, KH, KW, pad;
  int Ho, Wo;                 // output spatial size (== input size for the stride-1 "same" convs)
  int tile_w, tile_h;         // output patch of one CTA: tile_w * tile_h == 128
  int tiles_w, tiles_h;       // patches per image
  int kb_per_split, total_kb; // k-blocks (32 channels of one tap each)
};

constexpr int kTcThreads = 192;
constexpr int kBK = 32;                       // fp32 elements per 128-byte swizzle row
constexpr int kABytes = 128 * kBK * 4;        // 16 KB: one 128-row A tile

__device__ __forceinline__ float tc_act(float v, int act)
{
  if (act == FRCNN_ACT_RELU) return v > 0.0f ? v : 0.0f;
  if (act == FRCNN_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_conv_fwd_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   TcGeom g, float *__restrict__ out, float *__restrict__ partial, Epilogue epi)
{
  constexpr int kBBytes = BN * kBK * 4;
  constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * kStageBytes);
  uint64_t *empty = full + STAGES;
  uint64_t *tmem_full = empty + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  const int tiles_per_img = g.tiles_w * g.tiles_h;
  const int img = blockIdx.x / tiles_per_img;
  const int trem = blockIdx.x - img * tiles_per_img;
  const int oh0 = (trem / g.tiles_w) * g.tile_h;
  const int ow0 = (trem % g.tiles_w) * g.tile_w;
  const int n0 = blockIdx.y * BN;
  const int kb_begin = blockIdx.z * g.kb_per_split;
  int kb_end = kb_begin + g.kb_per_split;
  if (kb_end > g.total_kb) kb_end = g.total_kb;
  const int nkb = kb_end - kb_begin;
  const int cblocks = g.Cin / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int i = 0; i < nkb; i++) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int kb = kb_begin + i;
        const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * kBK;
        const int kh = tap / g.KW, kw = tap - kh * g.KW;
        uint8_t *st = smem + s * kStageBytes;
        mbar_expect_tx(&full[s], kStageBytes);
        tma_load_4d(st, &map_a_hi, &full[s], c0, ow0 + kw - g.pad, oh0 + kh - g.pad, img);
        tma_load_4d(st + kABytes, &map_a_lo, &full[s], c0, ow0 + kw - g.pad, oh0 + kh - g.pad, img);
        tma_load_2d(st + 2 * kABytes, &map_b_hi, &full[s], tap * g.Cin + c0, n0);
        tma_load_2d(st + 2 * kABytes + kBBytes, &map_b_lo, &full[s], tap * g.Cin + c0, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
      uint32_t accumulate = 0;
      for (int i = 0; i < nkb; i++) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * kStageBytes);
        const uint32_t a_lo = a_hi + kABytes;
        const uint32_t b_hi = a_hi + 2 * kABytes;
        const uint32_t b_lo = b_hi + kBBytes;
#pragma unroll
        for (int k = 0; k < kBK / 8; k++) {
          const uint64_t da_hi = make_smem_desc(a_hi + k * 32, 16, 1024);
          const uint64_t da_lo = make_smem_desc(a_lo + k * 32, 16, 1024);
          const uint64_t db_hi = make_smem_desc(b_hi + k * 32, 16, 1024);
          const uint64_t db_lo = make_smem_desc(b_lo + k * 32, 16, 1024);
          umma_tf32(tmem_base, da_hi, db_lo, idesc, accumulate);     // small terms first
          umma_tf32(tmem_base, da_lo, db_hi, idesc, 1);
          umma_tf32(tmem_base, da_hi, db_hi, idesc, 1);
          accumulate = 1;
        }
        umma_commit(&empty[s]);                                      // frees the slot when these MMAs retire
      }
      umma_commit(tmem_full);
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int oh = oh0 + row / g.tile_w, ow = ow0 + row % g.tile_w;
    const bool valid = oh < g.Ho && ow < g.Wo;
    const size_t pix = ((size_t)img * g.Ho + oh) * g.Wo + ow;
    const bool raw = gridDim.z > 1;
    // split-K: raw partial sums go to partial[z][pixel][Cout]; the reduce kernel applies the epilogue
    float *dst = raw ? partial + (size_t)blockIdx.z * ((size_t)(gridDim.x / tiles_per_img) * g.Ho * g.Wo) * g.Cout : out;
    for (int c = 0; c < BN / 32; c++) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
      if (nkb <= 0) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0.f;
      }
      if (valid) {
        const int col0 = n0 + c * 32;
        float *p = dst + pix * g.Cout + col0;
        if (!raw) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            float x = v[j];
            if (epi.scale) x *= __ldg(epi.scale + col0 + j);
            if (epi.bias) x += __ldg(epi.bias + col0 + j);
            if (epi.residual) x += __ldg(epi.residual + pix * g.Cout + col0 + j);
            v[j] = tc_act(x, epi.act);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// activation (N,H,W,C) fp32 as a 4-D tensor (C, W, H, N); box {32, box_w, box_h, 1}
static bool make_act_map(CUtensorMap *m, const float *base, int N, int H, int W, int C, int box_w, int box_h)
{
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  EncodeTiledFn f = encode_fn();
  if (!f) return false;
  return f(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// matrix (rows, K) fp32 row-major as a 2-D tensor (K, rows); box {32, box_rows}
static bool make_mat_map(CUtensorMap *m, const float *base, int rows, int K, int box_rows)
{
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  EncodeTiledFn f = encode_fn();
  if (!f) return false;
  return f(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct TcPlan {
  // geometry after folding nn.Linear (H=W=1) into a 1 x rows "image"
  int N, H, W;
  int BN, stages;
  int tile_w, tile_h, tiles_w, tiles_h;
  int total_kb, splits, kb_per_split;
  size_t x_lo_off, w_lo_off, partial_off, total_bytes;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static bool make_tc_plan(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, TcPlan *p)
{
  if (stride != 1 || (Cin % 32) != 0 || (Cout % 64) != 0) return false;
  if (KH != KW || 2 * pad != KH - 1) return false;                  // "same" convs and 1x1 / linear
  if (KH == 1 && H == 1 && W == 1) { p->N = 1; p->H = 1; p->W = N; }  // linear: rows become the W axis
  else { p->N = N; p->H = H; p->W = W; }
  if ((long long)p->N * p->H * p->W < 64) return false;             // tiny problems stay on the CUDA-core engine
  p->BN = (Cout % 128 == 0) ? 128 : 64;
  p->stages = p->BN == 128 ? 3 : 4;
  // output patch shape with the least padding waste (ties -> wider)
  long long best = -1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    int th = 128 / tw;
    long long area = (long long)ceil_div(p->W, tw) * tw * ceil_div(p->H, th) * th;
    if (best < 0 || area < best) { best = area; p->tile_w = tw; p->tile_h = th; }
  }
  p->tiles_w = ceil_div(p->W, p->tile_w);
  p->tiles_h = ceil_div(p->H, p->tile_h);
  p->total_kb = KH * KW * (Cin / 32);
  int ctas = p->N * p->tiles_w * p->tiles_h * (Cout / p->BN);
  int splits = 1;
  if (ctas < kNumSMs) {
    splits = ceil_div(kNumSMs, ctas);
    int max_splits = p->total_kb / 8;
    if (splits > max_splits) splits = max_splits;
    if (splits > 16) splits = 16;
    if (splits < 1) splits = 1;
  }
  p->kb_per_split = ceil_div(p->total_kb, splits);
  p->splits = ceil_div(p->total_kb, p->kb_per_split);
  size_t x_bytes = (size_t)p->N * p->H * p->W * Cin * 4;
  size_t w_bytes = (size_t)Cout * KH * KW * Cin * 4;
  p->x_lo_off = 0;
  p->w_lo_off = align_up(x_bytes, 1024);
  p->partial_off = p->w_lo_off + align_up(w_bytes, 1024);
  p->total_bytes = p->partial_off + (p->splits > 1 ? (size_t)p->splits * p->N * p->H * p->W * Cout * 4 : 0);
  return true;
}

bool tc_fwd_supported(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad)
{
  TcPlan p;
  return make_tc_plan(N, H, W, Cin, Cout, KH, KW, stride, pad, &p);
}

size_t tc_fwd_workspace(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad)
{
  TcPlan p;
  if (!make_tc_plan(N, H, W, Cin, Cout, KH, KW, stride, pad, &p)) return 0;
  return p.total_bytes;
}

template <int BN, int STAGES>
static int launch_fwd(const CUtensorMap &ma_hi, const CUtensorMap &ma_lo, const CUtensorMap &mb_hi, const CUtensorMap &mb_lo,
                      const TcGeom &g, const TcPlan &p, float *out, float *partial, const Epilogue &epi, int Cout, cudaStream_t st)
{
  constexpr int smem = STAGES * (2 * kABytes + 2 * BN * kBK * 4) + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(tc_conv_fwd_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "tc_conv_fwd_kernel: smem attribute");
  dim3 grid(p.N * p.tiles_w * p.tiles_h, Cout / BN, p.splits);
  tc_conv_fwd_kernel<BN, STAGES><<<grid, kTcThreads, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, g, out, partial, epi);
  FRCNN_CHECK_LAUNCH("tc_conv_fwd_kernel");
  return FRCNN_OK;
}

int tc_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual, float *y,
                  int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                  void *workspace, size_t workspace_bytes, cudaStream_t st)
{
  TcPlan p;
  if (!make_tc_plan(N, H, W, Cin, Cout, KH, KW, stride, pad, &p)) return fail(FRCNN_E_UNSUPPORTED, "tc_conv2d_fwd: unsupported shape");
  if (workspace == nullptr || workspace_bytes < p.total_bytes) return fail(FRCNN_E_WORKSPACE, "tc_conv2d_fwd: workspace too small");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return fail(FRCNN_E_BADARG, "tc_conv2d_fwd: operands and workspace must be 16-byte aligned");
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  float *x_lo = reinterpret_cast<float *>(ws + p.x_lo_off);
  float *w_lo = reinterpret_cast<float *>(ws + p.w_lo_off);
  float *partial = reinterpret_cast<float *>(ws + p.partial_off);
  const size_t x_count = (size_t)p.N * p.H * p.W * Cin, w_count = (size_t)Cout * KH * KW * Cin;
  split_lo_kernel<<<elementwise_grid(x_count / 4 + 1, 256), 256, 0, st>>>(x, x_lo, x_count);
  FRCNN_CHECK_LAUNCH("split_lo_kernel(x)");
  split_lo_kernel<<<elementwise_grid(w_count / 4 + 1, 256), 256, 0, st>>>(w, w_lo, w_count);
  FRCNN_CHECK_LAUNCH("split_lo_kernel(w)");

  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  const int K = KH * KW * Cin;
  bool ok = make_act_map(&ma_hi, x, p.N, p.H, p.W, Cin, p.tile_w, p.tile_h) && make_act_map(&ma_lo, x_lo, p.N, p.H, p.W, Cin, p.tile_w, p.tile_h) &&
            make_mat_map(&mb_hi, w, Cout, K, p.BN) && make_mat_map(&mb_lo, w_lo, Cout, K, p.BN);
  if (!ok) return fail(FRCNN_E_BADARG, "tc_conv2d_fwd: cuTensorMapEncodeTiled failed");

  TcGeom g{Cin, Cout, KH, KW, pad, p.H, p.W, p.tile_w, p.tile_h, p.tiles_w, p.tiles_h, p.kb_per_split, p.total_kb};
  Epilogue epi{scale, bias, residual, act};
  int rc;
  if (p.BN == 128) rc = launch_fwd<128, 3>(ma_hi, ma_lo, mb_hi, mb_lo, g, p, y, partial, epi, Cout, st);
  else rc = launch_fwd<64, 4>(ma_hi, ma_lo, mb_hi, mb_lo, g, p, y, partial, epi, Cout, st);
  if (rc != FRCNN_OK) return rc;
  if (p.splits > 1) return launch_splitk_reduce(partial, y, p.N * p.H * p.W, Cout, p.splits, epi, st);
  return FRCNN_OK;
}

// data / filter gradients on the tensor cores: next increment (MN-major operand variants); until
// then FRCNN_ENGINE_AUTO routes them to the fp32 CUDA-core engine.
bool tc_dgrad_supported(int, int, int, int, int, int, int, int, int) { return false; }
size_t tc_dgrad_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
int tc_conv2d_dgrad(const float *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int, void *, size_t, cudaStream_t)
{ return fail(FRCNN_E_UNSUPPORTED, "tcgen05 engine: dgrad not built"); }
bool tc_wgrad_supported(int, int, int, int, int, int, int, int, int) { return false; }
size_t tc_wgrad_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
int tc_conv2d_wgrad(const float *, const float *, float *, int, int, int, int, int, int, int, int, int, void *, size_t, cudaStream_t)
{ return fail(FRCNN_E_UNSUPPORTED, "tcgen05 engine: wgrad not built"); }

}  // namespace frcnn
