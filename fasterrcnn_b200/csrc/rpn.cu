// RPN proposal path: fused anchor generation + delta decode + clip + min-size flag (K5),
// top-N ordering, order-preserving compaction, greedy NMS (K6).
//
// Parity discipline: every fp32 operation that the reference performs as a separate torch
// kernel is an explicit __f*_rn intrinsic here so that ptxas cannot contract mul+add into FMA;
// anchors are built in fp64 from the 9x2 size table and rounded to fp32 exactly where
// models/anchors.py:118-135 rounds them.
#include <math.h>
#include "common.cuh"

namespace frcnn {

struct AnchorSizes {
  double h[9];
  double w[9];
};

// models/anchors.py:25-41: k = area*3 + aspect; w = sqrt(area/aspect); h = aspect*w  (fp64)
static AnchorSizes make_anchor_sizes()
{
  AnchorSizes s;
  const double areas[3] = {128.0 * 128.0, 256.0 * 256.0, 512.0 * 512.0};
  const double aspects[3] = {0.5, 1.0, 2.0};
  int k = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double w = sqrt(areas[i] / aspects[j]);
      s.w[k] = w;
      s.h[k] = aspects[j] * w;
      k++;
    }
  return s;
}

// ---- K5 ---------------------------------------------------------------------------------------
__global__ void rpn_decode_kernel(const float *__restrict__ deltas, const float *__restrict__ anchors_in, int fh, int fw, double feature_pixels, float img_h, float img_w,
                                  float min_size, AnchorSizes sizes, float *__restrict__ boxes, uint8_t *__restrict__ size_ok,
                                  float *__restrict__ anchors_out, float *__restrict__ valid_out)
{
  pdl_enter();
  const int A = fh * fw * 9;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < A; a += gridDim.x * blockDim.x) {
    int k = a % 9;
    int cell = a / 9;
    int x = cell % fw, y = cell / fw;
    // cell centre: fp64 value rounded to fp32 (anchors.py:104,118), then fp64 template added
    double cy = (double)__double2float_rn((double)y * feature_pixels + 0.5 * feature_pixels);
    double cx = (double)__double2float_rn((double)x * feature_pixels + 0.5 * feature_pixels);
    double y1 = cy + (-0.5 * sizes.h[k]), x1 = cx + (-0.5 * sizes.w[k]);
    double y2 = cy + (0.5 * sizes.h[k]), x2 = cx + (0.5 * sizes.w[k]);
    float acy = __double2float_rn(0.5 * (y1 + y2));
    float acx = __double2float_rn(0.5 * (x1 + x2));
    float ah = __double2float_rn(y2 - y1);
    float aw = __double2float_rn(x2 - x1);
    if (anchors_in) {
      float4 u = __ldg(reinterpret_cast<const float4 *>(anchors_in) + a);
      acy = u.x; acx = u.y; ah = u.z; aw = u.w;
    }
    if (anchors_out) reinterpret_cast<float4 *>(anchors_out)[a] = make_float4(acy, acx, ah, aw);
    if (valid_out) valid_out[a] = (y1 >= 0.0 && x1 >= 0.0 && y2 <= (double)img_h && x2 <= (double)img_w) ? 1.0f : 0.0f;

    // models/math_utils.py:122-127 in fp32: c = a_hw*t_yx + a_yx ; s = a_hw*exp(t_hw) ; box = c -/+ 0.5 s
    float4 d = __ldg(reinterpret_cast<const float4 *>(deltas) + a);
    float c_y = __fadd_rn(__fmul_rn(ah, d.x), acy);
    float c_x = __fadd_rn(__fmul_rn(aw, d.y), acx);
    float s_h = __fmul_rn(ah, __double2float_rn(exp((double)d.z)));   // correctly rounded expf
    float s_w = __fmul_rn(aw, __double2float_rn(exp((double)d.w)));
    float b0 = __fsub_rn(c_y, __fmul_rn(0.5f, s_h));
    float b1 = __fsub_rn(c_x, __fmul_rn(0.5f, s_w));
    float b2 = __fadd_rn(c_y, __fmul_rn(0.5f, s_h));
    float b3 = __fadd_rn(c_x, __fmul_rn(0.5f, s_w));
    // models/rpn.py:135-137: clamp y1,x1 >= 0; y2 <= H; x2 <= W
    b0 = b0 < 0.f ? 0.f : b0;
    b1 = b1 < 0.f ? 0.f : b1;
    b2 = b2 > img_h ? img_h : b2;
    b3 = b3 > img_w ? img_w : b3;
    reinterpret_cast<float4 *>(boxes)[a] = make_float4(b0, b1, b2, b3);
    // models/rpn.py:140-142
    size_ok[a] = (__fsub_rn(b2, b0) >= min_size && __fsub_rn(b3, b1) >= min_size) ? 1 : 0;
  }
}

// ---- top-N ordering by rank counting ------------------------------------------------------------
// rank(i) = #{j : s_j > s_i} + #{j : s_j == s_i and j > i}  (descending, ties -> higher index
// first).  The j range is split across blockIdx.y; partial ranks are integer atomics (exact,
// order independent).  N is ~2e4, so the N^2 = 4e8 compares run in a few microseconds on 148
// SMs and give a stable, deterministic order without a multi-pass radix sort.
constexpr int kRankTile = 1024;

// blockIdx.z = batch entry (scores / rank advance by n / 1 + n per entry).  LOW_FIRST selects the tie rule: false = higher index
// first (argsort ascending + flip, models/rpn.py:129-130), true = lower index first (stable descending, torchvision.ops.nms).
template <bool LOW_FIRST>
__global__ void rank_count_kernel(const float *__restrict__ scores, const uint8_t *__restrict__ keep_mask, int n, int j_per_slice, int32_t *__restrict__ rank)
{
  pdl_enter();
  __shared__ __align__(16) float tile[kRankTile];
  scores += (size_t)blockIdx.z * n;
  rank += (size_t)blockIdx.z * (n + 1);
  if (keep_mask) keep_mask += (size_t)blockIdx.z * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n && (keep_mask == nullptr || keep_mask[i] != 0);
  const float si = live ? scores[i] : 0.f;
  int j0 = blockIdx.y * j_per_slice;
  int j1 = j0 + j_per_slice < n ? j0 + j_per_slice : n;
  int cnt = 0;
  for (int base = j0; base < j1; base += kRankTile) {
    __syncthreads();
    for (int q = threadIdx.x; q < kRankTile; q += blockDim.x) {
      int j = base + q;
      // entries that do not take part are NaN: they compare false both ways
      tile[q] = (j < j1 && (keep_mask == nullptr || keep_mask[j] != 0)) ? scores[j] : __int_as_float(0x7fc00000);
    }
    __syncthreads();
    if (live) {
      int lim = j1 - base < kRankTile ? j1 - base : kRankTile;
      int q = 0;
      for (; q + 4 <= lim; q += 4) {
        float4 v = *reinterpret_cast<const float4 *>(&tile[q]);
        int j = base + q;
        cnt += (v.x > si) || (v.x == si && (LOW_FIRST ? j + 0 < i : j + 0 > i));
        cnt += (v.y > si) || (v.y == si && (LOW_FIRST ? j + 1 < i : j + 1 > i));
        cnt += (v.z > si) || (v.z == si && (LOW_FIRST ? j + 2 < i : j + 2 > i));
        cnt += (v.w > si) || (v.w == si && (LOW_FIRST ? j + 3 < i : j + 3 > i));
      }
      for (; q < lim; q++) {
        float v = tile[q];
        cnt += (v > si) || (v == si && (LOW_FIRST ? base + q < i : base + q > i));
      }
    }
  }
  if (live && cnt) atomicAdd(&rank[i], cnt);
}

// blockIdx.y = batch entry: rank / count_out advance by 1 + n, order by order_stride
__global__ void rank_scatter_kernel(const uint8_t *__restrict__ keep_mask, int n, int top_n, const int32_t *__restrict__ rank,
                                    int32_t *__restrict__ order, int32_t *__restrict__ count_out, int order_stride)
{
  pdl_enter();
  rank += (size_t)blockIdx.y * (n + 1);
  count_out += (size_t)blockIdx.y * (n + 1);
  order += (size_t)blockIdx.y * order_stride;
  if (keep_mask) keep_mask += (size_t)blockIdx.y * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = i < n && (keep_mask == nullptr || keep_mask[i] != 0);
  bool hit = false;
  if (live) {
    int r = rank[i];
    if (r < top_n) { order[r] = i; hit = true; }
  }
  unsigned m = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count_out, __popc(m));
}

// ---- order-preserving compaction of the ranked, size-filtered boxes (single CTA) ---------------
__global__ void __launch_bounds__(1024)
gather_filtered_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, const uint8_t *__restrict__ size_ok,
                       const int32_t *__restrict__ order, const int32_t *__restrict__ count, int capacity,
                       float *__restrict__ boxes_out, float *__restrict__ scores_out, int32_t *__restrict__ count_out)
{
  pdl_enter();
  __shared__ int warp_sums[32];
  __shared__ int carry;
  int n = *count;
  if (n > capacity) n = capacity;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    int r = base + threadIdx.x;
    int idx = r < n ? order[r] : -1;
    int flag = (idx >= 0 && size_ok[idx]) ? 1 : 0;
    // inclusive warp scan
    int v = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      warp_sums[lane] = w;                                   // inclusive sums over warps
    }
    __syncthreads();
    int offset = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + v - flag;
    if (flag) {
      reinterpret_cast<float4 *>(boxes_out)[offset] = __ldg(reinterpret_cast<const float4 *>(boxes) + idx);
      scores_out[offset] = scores[idx];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *count_out = carry;
}

// ---- K6: NMS ------------------------------------------------------------------------------------
// Stage 1: 64x64 suppression bit tiles, upper triangle only.  Arithmetic is the torchvision CPU
// kernel's, in fp32 with explicit roundings: area = (b2-b0)*(b3-b1); w = max(0, min-max) ...;
// ovr = inter / (area_i + area_j - inter); suppress iff ovr > thr (thr_f is the largest float
// <= the double threshold, which makes the float compare identical to the op's double compare).
__device__ __forceinline__ bool iou_exceeds(const float4 &a, float area_a, const float4 &b, float area_b, float thr_f)
{
  float t0 = fmaxf(a.x, b.x), t1 = fmaxf(a.y, b.y);
  float t2 = fminf(a.z, b.z), t3 = fminf(a.w, b.w);
  float w = fmaxf(0.f, __fsub_rn(t2, t0));
  float h = fmaxf(0.f, __fsub_rn(t3, t1));
  float inter = __fmul_rn(w, h);
  const float den = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  // the IEEE division costs ~8x the rest of the pair; an approximate quotient (2 ulp) decides every pair that is not within
  // 1e-5 of the threshold, the exact one only the rest -- the decision is the reference's bit for bit either way
  const float q = __fdividef(inter, den);
  if (fabsf(__fsub_rn(q, thr_f)) > 1e-5f) return q > thr_f;  // NaN / inf quotients fall through to the exact path
  float ovr = __fdiv_rn(inter, den);
  return ovr > thr_f;                                        // NaN (0/0) -> false, as in the reference op
}

__global__ void __launch_bounds__(64)
nms_mask_kernel(const float *__restrict__ boxes, const int32_t *__restrict__ count, int count_stride, int capacity, float thr_f,
                unsigned long long *__restrict__ mask, int col_blocks)
{
  pdl_enter();
  // blockIdx.z = batch entry (class): boxes / mask advance by one capacity-sized block, count by count_stride
  boxes += (size_t)blockIdx.z * capacity * 4;
  mask += (size_t)blockIdx.z * capacity * col_blocks;
  int n = count[(size_t)blockIdx.z * count_stride];
  if (n > capacity) n = capacity;
  const int row_b = blockIdx.y, col_b = blockIdx.x;
  if (col_b < row_b) return;
  if (row_b * 64 >= n || col_b * 64 >= n) return;
  __shared__ float4 cb[64];
  __shared__ float ca[64];
  const int t = threadIdx.x;
  {
    int j = col_b * 64 + t;
    float4 b = j < n ? __ldg(reinterpret_cast<const float4 *>(boxes) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    cb[t] = b;
    ca[t] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  }
  __syncthreads();
  int i = row_b * 64 + t;
  if (i < n) {
    float4 a = __ldg(reinterpret_cast<const float4 *>(boxes) + i);
    float area = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    int cols = n - col_b * 64 < 64 ? n - col_b * 64 : 64;
    int start = (row_b == col_b) ? t + 1 : 0;
    unsigned long long bits = 0;
    for (int q = start; q < cols; q++)
      if (iou_exceeds(a, area, cb[q], ca[q], thr_f)) bits |= 1ull << q;
    mask[(size_t)i * col_blocks + col_b] = bits;
  }
}

// fp64 twin of nms_mask_kernel for the per-class call site (models/faster_rcnn.py:216-220: float64 boxes, torchvision's kernel
// instantiated for double): same tiles, every operation an explicit IEEE double operation, exact division, strict compare.
__global__ void __launch_bounds__(64)
nms_mask_kernel_f64(const double *__restrict__ boxes, const int32_t *__restrict__ count, int capacity, double thr, unsigned long long *__restrict__ mask, int col_blocks)
{
  pdl_enter();
  int n = *count;
  if (n > capacity) n = capacity;
  const int row_b = blockIdx.y, col_b = blockIdx.x;
  if (col_b < row_b) return;
  if (row_b * 64 >= n || col_b * 64 >= n) return;
  __shared__ double cb[64][4];
  __shared__ double ca[64];
  const int t = threadIdx.x;
  {
    const int j = col_b * 64 + t;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    if (j < n) { b0 = boxes[(size_t)j * 4]; b1 = boxes[(size_t)j * 4 + 1]; b2 = boxes[(size_t)j * 4 + 2]; b3 = boxes[(size_t)j * 4 + 3]; }
    cb[t][0] = b0; cb[t][1] = b1; cb[t][2] = b2; cb[t][3] = b3;
    ca[t] = __dmul_rn(__dsub_rn(b2, b0), __dsub_rn(b3, b1));
  }
  __syncthreads();
  const int i = row_b * 64 + t;
  if (i < n) {
    const double a0 = boxes[(size_t)i * 4], a1 = boxes[(size_t)i * 4 + 1], a2 = boxes[(size_t)i * 4 + 2], a3 = boxes[(size_t)i * 4 + 3];
    const double area = __dmul_rn(__dsub_rn(a2, a0), __dsub_rn(a3, a1));
    const int cols = n - col_b * 64 < 64 ? n - col_b * 64 : 64;
    const int start = (row_b == col_b) ? t + 1 : 0;
    unsigned long long bits = 0;
    for (int q = start; q < cols; q++) {
      const double w = fmax(0.0, __dsub_rn(fmin(a2, cb[q][2]), fmax(a0, cb[q][0])));
      const double h = fmax(0.0, __dsub_rn(fmin(a3, cb[q][3]), fmax(a1, cb[q][1])));
      const double inter = __dmul_rn(w, h);
      const double ovr = __ddiv_rn(inter, __dsub_rn(__dadd_rn(area, ca[q]), inter));
      if (ovr > thr) bits |= 1ull << q;                        // NaN (0/0) -> false, as in the reference op
    }
    mask[(size_t)i * col_blocks + col_b] = bits;
  }
}

// Stage 2: the greedy scan, one CTA, software-pipelined over the 64-box blocks so that no global-memory round trip sits on
// the serial chain.  Block c may be resolved once its "removed" word holds the rows of every box kept in blocks < c:
//   * kept boxes of blocks < c-2: a column gather (one 8-byte load per kept box) ISSUED at iteration c-2 into registers and
//     folded at iteration c-1 -- a full iteration of slack for the L2 latency;
//   * kept boxes of blocks c-2 and c-1: their rows at columns +2 / +1, OR-ed in right after those blocks are resolved;
//   * the diagonal tile and those two row slices of every block depend on nothing and stream in three blocks ahead
//     (cp.async, 4-slot ring).
// Per iteration the chain is: warp 0 visits, in order, the surviving boxes of the block that suppress something inside it (ffs + one
// register shuffle per such box; boxes with an empty diagonal row need no visit), writes the kept positions in parallel, barrier,
// <=64 shared-memory atomics, barrier.  Stops at max_keep.
constexpr int kScanThreads = 1024;
constexpr int kScanGather = kScanThreads - 32;               // threads of warps 1..31 do the gathers
constexpr int kScanDefer = 3;                                // gather loads kept in flight per thread (covers max_keep <= 2976)

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, bool valid)
{
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 8 : 0;                                    // src-size 0 -> the 8 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long *__restrict__ mask, const int32_t *__restrict__ count, int count_stride, int capacity, int col_blocks,
                int max_keep, int32_t *__restrict__ keep_out, int keep_stride, int32_t *__restrict__ kept_count_out)
{
  pdl_enter();
  // blockIdx.x = batch entry (class): one CTA walks one greedy chain
  mask += (size_t)blockIdx.x * capacity * col_blocks;
  count += (size_t)blockIdx.x * count_stride;
  keep_out += (size_t)blockIdx.x * keep_stride;
  kept_count_out += blockIdx.x;
  extern __shared__ int32_t kept_idx[];                      // max_keep entries: positions of the kept boxes
  __shared__ unsigned long long ring[4][3][64];              // per block slot: [0] diagonal tile, [1] rows at column +1, [2] at column +2
  __shared__ unsigned long long removed[4];                  // removed[c & 3]: suppression word of block c, accumulated ahead of time
  __shared__ unsigned long long kept_bits_s;
  __shared__ int kept_total[2];                              // double-buffered by block parity: thread 0 writes [b + 1] while slower threads may still read [b]
  int n = *count;
  if (n > capacity) n = capacity;
  const int nblocks = (n + 63) / 64;
  const int t = threadIdx.x;
  if (t == 0) { kept_total[0] = kept_total[1] = 0; kept_bits_s = 0ull; removed[0] = removed[1] = removed[2] = removed[3] = 0ull; }

  // stream block `blk`'s three 64-word slices into ring slot blk & 3 (threads 64..255); one cp.async group per block
  auto prefetch_block = [&](int blk) {
    if (t >= 64 && t < 256 && blk < nblocks) {
      const int which = (t - 64) >> 6, j = (t - 64) & 63;
      const int i = blk * 64 + j, col = blk + which;
      const bool ok = i < n && col < nblocks;
      cp_async8(&ring[blk & 3][which][j], mask + (ok ? (size_t)i * col_blocks + col : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch_block(0);
  prefetch_block(1);
  prefetch_block(2);
  asm volatile("cp.async.wait_group 2;" ::: "memory");        // block 0 landed
  __syncthreads();

  unsigned long long pend[kScanDefer] = {0ull, 0ull, 0ull};  // gather issued last iteration, not yet folded
  unsigned long long pend_sync = 0ull;
  int pend_col = -1;
  int b = 0;
  for (; b < nblocks; b++) {
    const int kept_before = kept_total[b & 1];
    if (kept_before >= max_keep) break;                      // uniform (read after a barrier)
    if (t < 32) {
      // warp 0 resolves the block.  Only boxes that suppress something inside the block (non-zero diagonal row) have to be visited
      // in order; all the others are kept iff they are still alive at the end.  Lane l holds rows l and l + 32 of the diagonal tile.
      const int lim = n - b * 64 < 64 ? n - b * 64 : 64;
      const unsigned long long valid = lim == 64 ? ~0ull : ((1ull << lim) - 1ull);
      const unsigned long long *diag = ring[b & 3][0];
      const unsigned long long r0 = diag[t], r1 = diag[t + 32];
      const unsigned long long nz = (unsigned long long)__ballot_sync(0xffffffffu, r0 != 0ull) | ((unsigned long long)__ballot_sync(0xffffffffu, r1 != 0ull) << 32);
      unsigned long long alive = ~removed[b & 3] & valid;
      unsigned long long pending = alive & nz;
      while (pending) {                                        // ascending position = descending score: the greedy order
        const int q = __ffsll((long long)pending) - 1;
        const unsigned long long lo_row = __shfl_sync(0xffffffffu, r0, q & 31), hi_row = __shfl_sync(0xffffffffu, r1, q & 31);
        alive &= ~(q < 32 ? lo_row : hi_row);                  // a row only holds later positions (upper triangle)
        pending &= alive & ~(1ull << q);
      }
      unsigned long long kept = alive;
      const int room = max_keep - kept_before;
      while (__popcll(kept) > room) kept &= ~(1ull << (63 - __clzll((long long)kept)));      // only in the block that reaches max_keep
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int bit = t + 32 * half;
        if ((kept >> bit) & 1ull) {
          const int pos = kept_before + __popcll(kept & ((1ull << bit) - 1ull));
          kept_idx[pos] = b * 64 + bit;
          keep_out[pos] = b * 64 + bit;
        }
      }
      __syncwarp();
      if (t == 0) {
        kept_bits_s = kept;
        kept_total[(b + 1) & 1] = kept_before + __popcll(kept);
        removed[b & 3] = 0ull;                                 // slot is reused by block b + 4 (first written at iteration b + 2)
      }
    } else if (t >= 32) {
      if (pend_col >= 0) {                                   // fold the gather issued one iteration ago (column b + 1)
        unsigned long long acc = pend[0] | pend[1] | pend[2] | pend_sync;
        unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)acc);
        unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(acc >> 32));
        if ((t & 31) == 0) {
          unsigned long long w = ((unsigned long long)hi << 32) | lo;
          if (w) atomicOr(&removed[pend_col & 3], w);
        }
      }
      if (b + 2 < nblocks) {                                 // issue the gather for column b + 2 over the boxes kept in blocks < b
        const int g = t - 32;
#pragma unroll
        for (int u = 0; u < kScanDefer; u++) {
          const int k = g + u * kScanGather;
          pend[u] = k < kept_before ? mask[(size_t)kept_idx[k] * col_blocks + (b + 2)] : 0ull;
        }
        pend_sync = 0ull;
        for (int k = g + kScanDefer * kScanGather; k < kept_before; k += kScanGather) pend_sync |= mask[(size_t)kept_idx[k] * col_blocks + (b + 2)];
        pend_col = b + 2;
      } else {
        pend_col = -1;
      }
    }
    prefetch_block(b + 3);                                   // slot (b + 3) & 3 was last used by block b - 1
    asm volatile("cp.async.wait_group 2;" ::: "memory");      // groups up to block b + 1 complete (own copies; barrier publishes)
    __syncthreads();
    if (t < 64 && ((kept_bits_s >> t) & 1ull)) {
      const unsigned long long r1 = ring[b & 3][1][t], r2 = ring[b & 3][2][t];
      if (r1) atomicOr(&removed[(b + 1) & 3], r1);
      if (r2) atomicOr(&removed[(b + 2) & 3], r2);
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (t == 0) *kept_count_out = kept_total[b & 1];              // (break at block b: the count it started with; ran out of blocks: what block nblocks - 1 left)
}

// batched NMS glue: sorted[z][r] = boxes[z][order[z][r]] (r < n), and keep[z][r] = order[z][keep_pos[z][r]] (r < kept[z])
__global__ void nms_batched_sort_boxes_kernel(const float *__restrict__ boxes, const int32_t *__restrict__ order, int n, float *__restrict__ sorted)
{
  pdl_enter();
  const size_t zoff = (size_t)blockIdx.y * n;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    reinterpret_cast<float4 *>(sorted)[zoff + r] = __ldg(reinterpret_cast<const float4 *>(boxes) + zoff + order[zoff + r]);
}

__global__ void nms_batched_finish_kernel(const int32_t *__restrict__ order, const int32_t *__restrict__ keep_pos, const int32_t *__restrict__ kept, int n,
                                          int max_keep, int32_t *__restrict__ keep_out)
{
  pdl_enter();
  const int z = blockIdx.y;
  const int k = kept[z];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < max_keep; r += gridDim.x * blockDim.x)
    keep_out[(size_t)z * max_keep + r] = r < k ? order[(size_t)z * n + keep_pos[(size_t)z * max_keep + r]] : -1;
}

// dst rows [*dst_count, *dst_count + m) <- src rows: appends the ground-truth boxes behind the device-counted proposals
// (faster_rcnn.py:467) without bringing the count to the host
__global__ void append_rows_kernel(float *__restrict__ dst, const int32_t *__restrict__ dst_count, int dst_capacity_rows, int row_floats,
                                   const float *__restrict__ src, int m)
{
  pdl_enter();
  int n = __ldcg(dst_count);        // coherent, see gather_rows_kernel
  if (n < 0) n = 0;
  const int total = m * row_floats;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int r = e / row_floats;
    if (n + r < dst_capacity_rows) dst[(size_t)(n + r) * row_floats + (e - r * row_floats)] = src[e];
  }
}

__global__ void gather_rows_kernel(const float *__restrict__ src, int row_floats, const int32_t *__restrict__ index, const int32_t *__restrict__ count,
                                   int capacity, float *__restrict__ dst)
{
  pdl_enter();
  int n = __ldcg(count);            // a coherent load: ptxas may move a non-coherent (ld.global.nc) one above griddepcontrol.wait, and the count is the previous kernel's output
  if (n > capacity) n = capacity;
  size_t total = (size_t)n * row_floats;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(e / row_floats), c = (int)(e - (size_t)r * row_floats);
    dst[e] = src[(size_t)index[r] * row_floats + c];
  }
}


// ---- RPN ground-truth map (models/anchors.py:137-262) ------------------------------------------------
// IoU in fp64 between the anchor corners (formed in fp32 from the (cy,cx,h,w) map, then widened, as
// anchors.py:186-187 does) and the fp32 GT boxes; invalid anchors count as IoU -1.  Explicit fp64
// roundings (no FMA contraction) so that both kernels below reproduce identical IoU bits.
__device__ __forceinline__ double anchor_gt_iou(float ay1, float ax1, float ay2, float ax2, const float4 &g)
{
  double a0 = (double)ay1, a1 = (double)ax1, a2 = (double)ay2, a3 = (double)ax2;
  double g0 = (double)g.x, g1 = (double)g.y, g2 = (double)g.z, g3 = (double)g.w;
  double t0 = fmax(a0, g0), t1 = fmax(a1, g1), b0 = fmin(a2, g2), b1 = fmin(a3, g3);
  double inter = (t0 < b0 && t1 < b1) ? __dmul_rn(__dsub_rn(b0, t0), __dsub_rn(b1, t1)) : 0.0;
  double area_a = __dmul_rn(__dsub_rn(a2, a0), __dsub_rn(a3, a1));
  double area_g = __dmul_rn(__dsub_rn(g2, g0), __dsub_rn(g3, g1));
  double uni = __dsub_rn(__dadd_rn(area_a, area_g), inter);
  return __ddiv_rn(inter, __dadd_rn(uni, 1e-7));
}

__device__ __forceinline__ void anchor_corners(const float4 &a, float &y1, float &x1, float &y2, float &x2)
{
  y1 = __fsub_rn(a.x, __fmul_rn(0.5f, a.z)); x1 = __fsub_rn(a.y, __fmul_rn(0.5f, a.w));
  y2 = __fadd_rn(a.x, __fmul_rn(0.5f, a.z)); x2 = __fadd_rn(a.y, __fmul_rn(0.5f, a.w));
}

// pass 1: per-GT maximum IoU over all anchors (atomicMax on the int64 image of the double: order
// preserving for values >= 0, and -1 sorts below every non-negative value)
__global__ void rpn_targets_pass1(const float *__restrict__ anchors, const float *__restrict__ valid, int A, const float *__restrict__ gt, int M,
                                  long long *__restrict__ gt_max_bits)
{
  pdl_enter();
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < A; a += gridDim.x * blockDim.x) {
    float4 an = __ldg(reinterpret_cast<const float4 *>(anchors) + a);
    float y1, x1, y2, x2;
    anchor_corners(an, y1, x1, y2, x2);
    bool ok = valid[a] != 0.f;
    for (int m = 0; m < M; m++) {
      double iou = ok ? anchor_gt_iou(y1, x1, y2, x2, __ldg(reinterpret_cast<const float4 *>(gt) + m)) : -1.0;
      atomicMax(&gt_max_bits[m], __double_as_longlong(iou));
    }
  }
}

__global__ void rpn_targets_pass2(const float *__restrict__ anchors, const float *__restrict__ valid, int A, const float *__restrict__ gt, int M,
                                  const long long *__restrict__ gt_max_bits, double object_thr, double background_thr, float *__restrict__ rpn_map)
{
  pdl_enter();
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < A; a += gridDim.x * blockDim.x) {
    float4 an = __ldg(reinterpret_cast<const float4 *>(anchors) + a);
    float y1, x1, y2, x2;
    anchor_corners(an, y1, x1, y2, x2);
    bool ok = valid[a] != 0.f;
    double best = -INFINITY;
    int which = 0;
    bool top = false;
    for (int m = 0; m < M; m++) {
      double iou = ok ? anchor_gt_iou(y1, x1, y2, x2, __ldg(reinterpret_cast<const float4 *>(gt) + m)) : -1.0;
      if (iou > best) { best = iou; which = m; }                     // np.argmax: first maximum
      if (__double_as_longlong(iou) == gt_max_bits[m]) top = true;   // ious == max_iou_per_gt_box
    }
    int label = -1;
    if (best < background_thr) label = 0;
    if (best >= object_thr) label = 1;
    if (top) label = 1;
    float enable = label >= 0 ? 1.f : 0.f;
    if (label < 0) label = 0;
    // regression targets in fp32 (anchors.py:236-238 operate on float32 arrays)
    float4 g = __ldg(reinterpret_cast<const float4 *>(gt) + which);
    float gcy = __fmul_rn(0.5f, __fadd_rn(g.x, g.z)), gcx = __fmul_rn(0.5f, __fadd_rn(g.y, g.w));
    float gh = __fsub_rn(g.z, g.x), gw = __fsub_rn(g.w, g.y);
    float *o = rpn_map + (size_t)a * 6;
    o[0] = __fmul_rn(valid[a], enable);
    o[1] = (float)label;
    o[2] = __fdiv_rn(__fsub_rn(gcy, an.x), an.z);
    o[3] = __fdiv_rn(__fsub_rn(gcx, an.y), an.w);
    o[4] = __double2float_rn(log((double)__fdiv_rn(gh, an.z)));
    o[5] = __double2float_rn(log((double)__fdiv_rn(gw, an.w)));
  }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_rpn_decode(const float *deltas, const float *anchors_in, int fh, int fw, int feature_pixels, int img_h, int img_w, float min_size,
                     float *boxes, uint8_t *size_ok, float *anchors_out, float *valid_out, void *stream)
{
  FRCNN_REQUIRE(deltas && boxes && size_ok && fh > 0 && fw > 0 && feature_pixels > 0 && img_h > 0 && img_w > 0, "rpn_decode: bad argument");
  const int A = fh * fw * 9;
  static const AnchorSizes sizes = make_anchor_sizes();
  launch(rpn_decode_kernel, elementwise_grid(A, 128, 8), 128, 0, as_stream(stream), deltas, anchors_in, fh, fw, (double)feature_pixels, (float)img_h, (float)img_w, min_size, sizes, boxes, size_ok, anchors_out, valid_out);
  FRCNN_CHECK_LAUNCH("rpn_decode_kernel");
  return FRCNN_OK;
}

int frcnn_topk_order(const float *scores, const uint8_t *keep_mask, int n, int top_n, int32_t *order, int32_t *count_out, void *stream)
{
  FRCNN_REQUIRE(scores && order && count_out && n > 0 && top_n > 0, "topk_order: bad argument");
  cudaStream_t st = as_stream(stream);
  // the per-element rank array is count_out[1..n] (the caller allocates 1 + n int32)
  int32_t *rank = count_out + 1;
  cudaError_t e = cudaMemsetAsync(count_out, 0, (size_t)(1 + n) * sizeof(int32_t), st);
  if (e != cudaSuccess) return cuda_fail(e, "topk_order: memset");
  const int threads = 256;
  int gx = ceil_div(n, threads);
  int slices = ceil_div(4 * kNumSMs, gx);
  if (slices < 1) slices = 1;
  int j_per_slice = ceil_div(ceil_div(n, slices), kRankTile) * kRankTile;
  slices = ceil_div(n, j_per_slice);
  launch(rank_count_kernel<false>, dim3(gx, slices), threads, 0, st, scores, keep_mask, n, j_per_slice, rank);
  FRCNN_CHECK_LAUNCH("rank_count_kernel");
  launch(rank_scatter_kernel, gx, threads, 0, st, keep_mask, n, top_n, rank, order, count_out, 0);
  FRCNN_CHECK_LAUNCH("rank_scatter_kernel");
  return FRCNN_OK;
}

int frcnn_gather_filtered(const float *boxes, const float *scores, const uint8_t *size_ok, const int32_t *order,
                          const int32_t *count, int capacity, float *boxes_out, float *scores_out, int32_t *count_out, void *stream)
{
  FRCNN_REQUIRE(boxes && scores && size_ok && order && count && boxes_out && scores_out && count_out && capacity > 0, "gather_filtered: bad argument");
  launch(gather_filtered_kernel, 1, 1024, 0, as_stream(stream), boxes, scores, size_ok, order, count, capacity, boxes_out, scores_out, count_out);
  FRCNN_CHECK_LAUNCH("gather_filtered_kernel");
  return FRCNN_OK;
}

size_t frcnn_nms_workspace_bytes(int capacity)
{
  if (capacity <= 0) return 0;
  size_t col_blocks = (size_t)ceil_div(capacity, 64);
  return (size_t)capacity * col_blocks * sizeof(unsigned long long);
}

int frcnn_nms_sorted_f32(const float *boxes, const int32_t *count, int capacity, double iou_threshold, int max_keep,
                         int32_t *keep_out, int32_t *kept_count_out, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(boxes && count && keep_out && kept_count_out && capacity > 0 && max_keep > 0, "nms_sorted_f32: bad argument");
  if (workspace == nullptr || workspace_bytes < frcnn_nms_workspace_bytes(capacity)) return fail(FRCNN_E_WORKSPACE, "nms_sorted_f32: workspace too small");
  const int col_blocks = ceil_div(capacity, 64);
  float thr_f = (float)iou_threshold;
  if ((double)thr_f > iou_threshold) thr_f = nextafterf(thr_f, -INFINITY);
  cudaStream_t st = as_stream(stream);
  unsigned long long *mask = reinterpret_cast<unsigned long long *>(workspace);
  launch(nms_mask_kernel, dim3(col_blocks, col_blocks), 64, 0, st, boxes, count, 0, capacity, thr_f, mask, col_blocks);
  FRCNN_CHECK_LAUNCH("nms_mask_kernel");
  const int keep_cap = max_keep < capacity ? max_keep : capacity;
  size_t smem = (size_t)keep_cap * sizeof(int32_t);
  FRCNN_REQUIRE(smem <= 200 * 1024, "nms_sorted_f32: max_keep too large for the scan's kept list");
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "nms_scan_kernel: smem attribute");
  }
  launch(nms_scan_kernel, 1, kScanThreads, smem, st, mask, count, 0, capacity, col_blocks, keep_cap, keep_out, 0, kept_count_out);
  FRCNN_CHECK_LAUNCH("nms_scan_kernel");
  return FRCNN_OK;
}

int frcnn_nms_sorted_f64(const double *boxes, const int32_t *count, int capacity, double iou_threshold, int max_keep,
                         int32_t *keep_out, int32_t *kept_count_out, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(boxes && count && keep_out && kept_count_out && capacity > 0 && max_keep > 0, "nms_sorted_f64: bad argument");
  if (workspace == nullptr || workspace_bytes < frcnn_nms_workspace_bytes(capacity)) return fail(FRCNN_E_WORKSPACE, "nms_sorted_f64: workspace too small");
  const int col_blocks = ceil_div(capacity, 64);
  cudaStream_t st = as_stream(stream);
  unsigned long long *mask = reinterpret_cast<unsigned long long *>(workspace);
  launch(nms_mask_kernel_f64, dim3(col_blocks, col_blocks), 64, 0, st, boxes, count, capacity, iou_threshold, mask, col_blocks);
  FRCNN_CHECK_LAUNCH("nms_mask_kernel_f64");
  const int keep_cap = max_keep < capacity ? max_keep : capacity;
  size_t smem = (size_t)keep_cap * sizeof(int32_t);
  FRCNN_REQUIRE(smem <= 200 * 1024, "nms_sorted_f64: max_keep too large for the scan's kept list");
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "nms_scan_kernel: smem attribute");
  }
  launch(nms_scan_kernel, 1, kScanThreads, smem, st, mask, count, 0, capacity, col_blocks, keep_cap, keep_out, 0, kept_count_out);
  FRCNN_CHECK_LAUNCH("nms_scan_kernel");
  return FRCNN_OK;
}

// workspace layout of the batched NMS (each block 256-byte aligned): cnt (B, 1 + n) i32 | order (B, n) i32 | sorted (B, n, 4) f32 |
// keep_pos (B, max_keep) i32 | mask (B, n, ceil(n / 64)) u64
static size_t nms_batched_layout(int B, int n, int max_keep, size_t off[5])
{
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o = 0;
  off[0] = o; o += up((size_t)B * (1 + n) * 4);
  off[1] = o; o += up((size_t)B * n * 4);
  off[2] = o; o += up((size_t)B * n * 16);
  off[3] = o; o += up((size_t)B * max_keep * 4);
  off[4] = o; o += up((size_t)B * n * ceil_div(n, 64) * 8);
  return o;
}

size_t frcnn_nms_batched_workspace_bytes(int B, int n, int max_keep)
{
  if (B <= 0 || n <= 0 || max_keep <= 0) return 0;
  size_t off[5];
  return nms_batched_layout(B, n, max_keep < n ? max_keep : n, off);
}

int frcnn_nms_batched_f32(const float *boxes, const float *scores, int B, int n, double iou_threshold, int max_keep,
                          int32_t *keep_out, int32_t *kept_count_out, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(boxes && scores && keep_out && kept_count_out && B > 0 && B <= 65535 && n > 0 && max_keep > 0, "nms_batched_f32: bad argument");
  const int keep_cap = max_keep < n ? max_keep : n;
  size_t off[5];
  const size_t need = nms_batched_layout(B, n, keep_cap, off);
  if (workspace == nullptr || workspace_bytes < need) return fail(FRCNN_E_WORKSPACE, "nms_batched_f32: workspace too small");
  FRCNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "nms_batched_f32: boxes and workspace must be 16-byte aligned");
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  int32_t *cnt = reinterpret_cast<int32_t *>(ws + off[0]);
  int32_t *order = reinterpret_cast<int32_t *>(ws + off[1]);
  float *sorted = reinterpret_cast<float *>(ws + off[2]);
  int32_t *keep_pos = reinterpret_cast<int32_t *>(ws + off[3]);
  unsigned long long *mask = reinterpret_cast<unsigned long long *>(ws + off[4]);
  cudaStream_t st = as_stream(stream);
  float thr_f = (float)iou_threshold;
  if ((double)thr_f > iou_threshold) thr_f = nextafterf(thr_f, -INFINITY);
  cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)B * (1 + n) * sizeof(int32_t), st);
  if (e != cudaSuccess) return cuda_fail(e, "nms_batched_f32: memset");
  // 1. stable descending order per entry (rank counting, ties -> lower index first)
  const int threads = 256;
  const int gx = ceil_div(n, threads);
  int slices = ceil_div(4 * kNumSMs, gx * B);
  if (slices < 1) slices = 1;
  int j_per_slice = ceil_div(ceil_div(n, slices), kRankTile) * kRankTile;
  slices = ceil_div(n, j_per_slice);
  launch(rank_count_kernel<true>, dim3(gx, slices, B), threads, 0, st, scores, nullptr, n, j_per_slice, cnt + 1);
  FRCNN_CHECK_LAUNCH("rank_count_kernel");
  launch(rank_scatter_kernel, dim3(gx, B), threads, 0, st, nullptr, n, n, cnt + 1, order, cnt, n);
  FRCNN_CHECK_LAUNCH("rank_scatter_kernel");
  launch(nms_batched_sort_boxes_kernel, dim3(ceil_div(n, 256), B), 256, 0, st, boxes, order, n, sorted);
  FRCNN_CHECK_LAUNCH("nms_batched_sort_boxes_kernel");
  // 2. suppression bit tiles of every entry in one launch, 3. one greedy-scan CTA per entry
  const int col_blocks = ceil_div(n, 64);
  launch(nms_mask_kernel, dim3(col_blocks, col_blocks, B), 64, 0, st, sorted, cnt, 1 + n, n, thr_f, mask, col_blocks);
  FRCNN_CHECK_LAUNCH("nms_mask_kernel");
  size_t smem = (size_t)keep_cap * sizeof(int32_t);
  FRCNN_REQUIRE(smem <= 200 * 1024, "nms_batched_f32: max_keep too large for the scan's kept list");
  if (smem > 40 * 1024) {
    e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "nms_scan_kernel: smem attribute");
  }
  launch(nms_scan_kernel, B, kScanThreads, smem, st, mask, cnt, 1 + n, n, col_blocks, keep_cap, keep_pos, keep_cap, kept_count_out);
  FRCNN_CHECK_LAUNCH("nms_scan_kernel");
  // 4. positions in the sorted list -> original indices
  launch(nms_batched_finish_kernel, dim3(ceil_div(keep_cap, 256), B), 256, 0, st, order, keep_pos, kept_count_out, n, keep_cap, keep_out);
  FRCNN_CHECK_LAUNCH("nms_batched_finish_kernel");
  return FRCNN_OK;
}

int frcnn_gather_rows_f32(const float *src, int row_floats, const int32_t *index, const int32_t *count, int capacity, float *dst, void *stream)
{
  FRCNN_REQUIRE(src && index && count && dst && row_floats > 0 && capacity > 0, "gather_rows_f32: bad argument");
  launch(gather_rows_kernel, elementwise_grid((size_t)capacity * row_floats, 256), 256, 0, as_stream(stream), src, row_floats, index, count, capacity, dst);
  FRCNN_CHECK_LAUNCH("gather_rows_kernel");
  return FRCNN_OK;
}

int frcnn_append_rows_f32(float *dst, const int32_t *dst_count, int dst_capacity_rows, int row_floats, const float *src, int m, void *stream)
{
  FRCNN_REQUIRE(dst && dst_count && src && dst_capacity_rows > 0 && row_floats > 0 && m > 0, "append_rows_f32: bad argument");
  launch(append_rows_kernel, ceil_div(m * row_floats, 128), 128, 0, as_stream(stream), dst, dst_count, dst_capacity_rows, row_floats, src, m);
  FRCNN_CHECK_LAUNCH("append_rows_kernel");
  return FRCNN_OK;
}

int frcnn_rpn_targets(const float *anchors, const float *valid, int A, const float *gt_boxes, int M, double object_iou_threshold,
                      double background_iou_threshold, float *rpn_map, void *workspace, size_t workspace_bytes, void *stream)
{
  FRCNN_REQUIRE(anchors && valid && gt_boxes && rpn_map && A > 0 && M > 0, "rpn_targets: bad argument");
  if (workspace == nullptr || workspace_bytes < (size_t)M * sizeof(long long)) return fail(FRCNN_E_WORKSPACE, "rpn_targets: workspace too small (need 8*M bytes)");
  cudaStream_t st = as_stream(stream);
  long long *gt_max = reinterpret_cast<long long *>(workspace);
  // int64 image of -1.0 in every slot: 0xBFF0000000000000 -> byte pattern is not uniform, so fill with a kernel-free trick:
  // memset to 0x80 gives a large-magnitude negative int64, below the image of -1.0 and of every IoU >= 0
  cudaError_t e = cudaMemsetAsync(gt_max, 0x80, (size_t)M * sizeof(long long), st);
  if (e != cudaSuccess) return cuda_fail(e, "rpn_targets: memset");
  launch(rpn_targets_pass1, elementwise_grid(A, 128, 2), 128, 0, st, anchors, valid, A, gt_boxes, M, gt_max);
  FRCNN_CHECK_LAUNCH("rpn_targets_pass1");
  launch(rpn_targets_pass2, elementwise_grid(A, 128, 2), 128, 0, st, anchors, valid, A, gt_boxes, M, gt_max, object_iou_threshold, background_iou_threshold, rpn_map);
  FRCNN_CHECK_LAUNCH("rpn_targets_pass2");
  return FRCNN_OK;
}

}  // extern "C"
