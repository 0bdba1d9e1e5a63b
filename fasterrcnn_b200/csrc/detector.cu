// Proposal labelling (a10), fused losses with gradients (a12), row softmax, and the batched
// per-class inference post-processing (a13) -- small, latency-bound kernels whose value is fusion:
// each replaces 10-40 library launches and several host syncs of the reference path.
#include <math.h>
#include "common.cuh"

namespace frcnn {

// ---- a10: models/faster_rcnn.py:418-524 -----------------------------------------------------------
__global__ void label_proposals_kernel(const float *__restrict__ proposals, int n, const float *__restrict__ gt_boxes, const int32_t *__restrict__ gt_classes,
                                       int m, int num_classes, float min_object_iou, float *__restrict__ best_iou_out, int32_t *__restrict__ class_out,
                                       float *__restrict__ onehot, float *__restrict__ packed)
{
  pdl_enter();
  const int D = 4 * (num_classes - 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(reinterpret_cast<const float4 *>(proposals) + i);
    float area_p = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
    float best = -INFINITY;
    int which = 0;
    for (int j = 0; j < m; j++) {
      float4 g = __ldg(reinterpret_cast<const float4 *>(gt_boxes) + j);
      // models/math_utils.py:39-63: strict well-ordered mask, eps in the denominator
      float t0 = fmaxf(p.x, g.x), t1 = fmaxf(p.y, g.y), t2 = fminf(p.z, g.z), t3 = fminf(p.w, g.w);
      float inter = (t0 < t2 && t1 < t3) ? __fmul_rn(__fsub_rn(t2, t0), __fsub_rn(t3, t1)) : 0.f;
      float area_g = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
      float uni = __fsub_rn(__fadd_rn(area_p, area_g), inter);
      float iou = __fdiv_rn(inter, __fadd_rn(uni, 1e-7f));
      if (iou > best) { best = iou; which = j; }            // first maximum wins (t.max / t.argmax)
    }
    float4 g = __ldg(reinterpret_cast<const float4 *>(gt_boxes) + which);
    int cls = best < min_object_iou ? 0 : gt_classes[which];
    best_iou_out[i] = best;
    class_out[i] = cls;
    for (int c = 0; c < num_classes; c++) onehot[(size_t)i * num_classes + c] = (c == cls) ? 1.f : 0.f;
    // targets: ((gt_c - p_c)/p_s, log(gt_s/p_s)) / (0.1, 0.1, 0.2, 0.2)
    float pcy = __fmul_rn(0.5f, __fadd_rn(p.x, p.z)), pcx = __fmul_rn(0.5f, __fadd_rn(p.y, p.w));
    float psh = __fsub_rn(p.z, p.x), psw = __fsub_rn(p.w, p.y);
    float gcy = __fmul_rn(0.5f, __fadd_rn(g.x, g.z)), gcx = __fmul_rn(0.5f, __fadd_rn(g.y, g.w));
    float gsh = __fsub_rn(g.z, g.x), gsw = __fsub_rn(g.w, g.y);
    float tg[4];
    tg[0] = __fdiv_rn(__fdiv_rn(__fsub_rn(gcy, pcy), psh), 0.1f);
    tg[1] = __fdiv_rn(__fdiv_rn(__fsub_rn(gcx, pcx), psw), 0.1f);
    tg[2] = __fdiv_rn(__double2float_rn(log((double)__fdiv_rn(gsh, psh))), 0.2f);
    tg[3] = __fdiv_rn(__double2float_rn(log((double)__fdiv_rn(gsw, psw))), 0.2f);
    float *row_mask = packed + (size_t)i * 2 * D;
    float *row_tg = row_mask + D;
    for (int q = 0; q < D; q++) {
      row_mask[q] = (cls > 0 && (q >> 2) == cls - 1) ? 1.f : 0.f;
      row_tg[q] = tg[q & 3];
    }
  }
}

// ---- models/math_utils.py:39-63 (t_intersection_over_union) and :99-128 (t_convert_deltas_to_boxes) as standalone operators --------
// The hot path has both fused into larger kernels (label_proposals above, rpn_decode); these are the helpers under the reference's names.
__global__ void iou_matrix_kernel(const float *__restrict__ boxes1, int n, const float *__restrict__ boxes2, int m, float *__restrict__ out)
{
  pdl_enter();
  const size_t total = (size_t)n * m;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / m), j = (int)(idx - (size_t)i * m);
    const float4 p = __ldg(reinterpret_cast<const float4 *>(boxes1) + i);
    const float4 g = __ldg(reinterpret_cast<const float4 *>(boxes2) + j);
    const float t0 = fmaxf(p.x, g.x), t1 = fmaxf(p.y, g.y), t2 = fminf(p.z, g.z), t3 = fminf(p.w, g.w);
    const float inter = (t0 < t2 && t1 < t3) ? __fmul_rn(__fsub_rn(t2, t0), __fsub_rn(t3, t1)) : 0.f;       // strict well-ordered mask
    const float area_p = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
    const float area_g = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
    out[idx] = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_p, area_g), inter), 1e-7f));             // eps in the denominator
  }
}

struct Vec4f { float v[4]; };

__global__ void decode_boxes_kernel(const float *__restrict__ deltas, const float *__restrict__ anchors, int n, Vec4f means, Vec4f stds, float *__restrict__ boxes)
{
  pdl_enter();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 d0 = __ldg(reinterpret_cast<const float4 *>(deltas) + i);
    const float4 a = __ldg(reinterpret_cast<const float4 *>(anchors) + i);                                   // (cy, cx, h, w)
    // d = d * stds + means; c = a_hw * d_yx + a_yx; s = a_hw * exp(d_hw); box = c -/+ 0.5 s -- separate roundings, exp correctly rounded
    const float dy = __fadd_rn(__fmul_rn(d0.x, stds.v[0]), means.v[0]), dx = __fadd_rn(__fmul_rn(d0.y, stds.v[1]), means.v[1]);
    const float dh = __fadd_rn(__fmul_rn(d0.z, stds.v[2]), means.v[2]), dw = __fadd_rn(__fmul_rn(d0.w, stds.v[3]), means.v[3]);
    const float cy = __fadd_rn(__fmul_rn(a.z, dy), a.x), cx = __fadd_rn(__fmul_rn(a.w, dx), a.y);
    const float sh = __fmul_rn(a.z, __double2float_rn(exp((double)dh))), sw = __fmul_rn(a.w, __double2float_rn(exp((double)dw)));
    reinterpret_cast<float4 *>(boxes)[i] = make_float4(__fsub_rn(cy, __fmul_rn(0.5f, sh)), __fsub_rn(cx, __fmul_rn(0.5f, sw)),
                                                       __fadd_rn(cy, __fmul_rn(0.5f, sh)), __fadd_rn(cx, __fmul_rn(0.5f, sw)));
  }
}

// ---- block reduction helper (double, fixed tree order -> deterministic) -------------------------
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *scratch)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double total = 0.0;
  if (warp == 0) {
    total = lane < THREADS / 32 ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_down_sync(0xffffffffu, total, o);
    if (lane == 0) scratch[0] = total;
  }
  __syncthreads();
  total = scratch[0];
  return total;
}

__device__ __forceinline__ float smooth_l1(float x, float sigma_sq, float *grad)
{
  float ax = fabsf(x);
  if (ax < 1.0f / sigma_sq) { *grad = sigma_sq * x; return 0.5f * x * x * sigma_sq; }
  *grad = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
  return ax - 0.5f / sigma_sq;
}

// ---- a12 RPN: models/rpn.py:176-272 --------------------------------------------------------------
// pass 1 (count + sums) and pass 2 (gradients, which need the normaliser) in one single-CTA kernel.
__global__ void __launch_bounds__(1024)
rpn_losses_kernel(const float *__restrict__ scores, const float *__restrict__ deltas, const float *__restrict__ y_true, int A,
                  float *__restrict__ losses_out, float *__restrict__ d_scores, float *__restrict__ d_deltas)
{
  pdl_enter();
  __shared__ double scratch[32];
  double cls_sum = 0.0, reg_sum = 0.0, cnt = 0.0;
  for (int a = threadIdx.x; a < A; a += blockDim.x) {
    const float *y = y_true + (size_t)a * 6;
    float mask = y[0], tgt = y[1];
    if (mask != 0.f) cnt += 1.0;
    float s = scores[a];
    // F.binary_cross_entropy: logs clamped at -100
    float l = (tgt - 1.f) * fmaxf(logf(1.f - s), -100.f) - tgt * fmaxf(logf(s), -100.f);
    cls_sum += (double)(mask * l);
    float m4 = mask * tgt;
    float4 d = *reinterpret_cast<const float4 *>(deltas + (size_t)a * 4);
    float g;
    reg_sum += (double)(m4 * smooth_l1(y[2] - d.x, 9.f, &g));
    reg_sum += (double)(m4 * smooth_l1(y[3] - d.y, 9.f, &g));
    reg_sum += (double)(m4 * smooth_l1(y[4] - d.z, 9.f, &g));
    reg_sum += (double)(m4 * smooth_l1(y[5] - d.w, 9.f, &g));
  }
  double count = block_sum<1024>(cnt, scratch);
  double cls_total = block_sum<1024>(cls_sum, scratch);
  double reg_total = block_sum<1024>(reg_sum, scratch);
  const float n_cls = (float)count + 1e-7f;                  // count_nonzero(mask) + epsilon in fp32
  if (threadIdx.x == 0) {
    losses_out[0] = (float)cls_total / n_cls;
    losses_out[1] = (float)reg_total / n_cls;
  }
  if (d_scores == nullptr) return;
  const float inv = 1.0f / n_cls;
  for (int a = threadIdx.x; a < A; a += blockDim.x) {
    const float *y = y_true + (size_t)a * 6;
    float mask = y[0], tgt = y[1];
    float s = scores[a];
    // binary_cross_entropy backward: dL/ds = mask/N * (s - y) / max((1-s) s, 1e-12)
    d_scores[a] = mask * inv * (s - tgt) / fmaxf((1.f - s) * s, 1e-12f);
    float m4 = mask * tgt * inv;
    float4 d = *reinterpret_cast<const float4 *>(deltas + (size_t)a * 4);
    float4 gd;
    float g;
    smooth_l1(y[2] - d.x, 9.f, &g); gd.x = -m4 * g;
    smooth_l1(y[3] - d.y, 9.f, &g); gd.y = -m4 * g;
    smooth_l1(y[4] - d.z, 9.f, &g); gd.z = -m4 * g;
    smooth_l1(y[5] - d.w, 9.f, &g); gd.w = -m4 * g;
    *reinterpret_cast<float4 *>(d_deltas + (size_t)a * 4) = gd;
  }
}

// ---- softmax over rows (F.softmax(dim=1), models/detector.py:77): one warp per row ------------------
__global__ void softmax_rows_kernel(const float *__restrict__ logits, float *__restrict__ probs, int n, int C)
{
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float *x = logits + (size_t)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(x[c] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int c = lane; c < C; c += 32) probs[(size_t)row * C + c] = expf(x[c] - mx) / sum;
}

// softmax backward: dlogit_c = p_c (g_c - sum_k g_k p_k); one warp per row
__global__ void softmax_rows_bwd_kernel(const float *__restrict__ probs, const float *__restrict__ d_probs, float *__restrict__ d_logits, int n, int C)
{
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float *p = probs + (size_t)row * C;
  const float *g = d_probs + (size_t)row * C;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot += g[c] * p[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int c = lane; c < C; c += 32) d_logits[(size_t)row * C + c] = p[c] * (g[c] - dot);
}

// sigmoid backward: dz = dy * (1 - y) * y
__global__ void sigmoid_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ y, float *__restrict__ dz, size_t count)
{
  pdl_enter();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    float s = y[i];
    dz[i] = dy[i] * (1.f - s) * s;
  }
}

// ---- a12 detector: models/detector.py:83-155 -------------------------------------------------------
__global__ void __launch_bounds__(1024)
detector_losses_kernel(const float *__restrict__ probs, const float *__restrict__ deltas, const float *__restrict__ y_classes,
                       const float *__restrict__ y_deltas, int n, int C, float *__restrict__ losses_out,
                       float *__restrict__ d_probs, float *__restrict__ d_deltas)
{
  pdl_enter();
  __shared__ double scratch[32];
  const int D = 4 * (C - 1);
  const float N = (float)((double)n + 1e-7);
  const float inv = 1.0f / N;
  double cls_sum = 0.0, reg_sum = 0.0;
  // class loss + softmax backward: one thread per row (C is 21)
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const float *p = probs + (size_t)r * C;
    const float *y = y_classes + (size_t)r * C;
    float row = 0.f;
    for (int c = 0; c < C; c++) {
      row += y[c] * logf(p[c] + 1e-7f);
      if (d_probs) d_probs[(size_t)r * C + c] = -y[c] / (p[c] + 1e-7f) * inv;   // dL/dp
    }
    cls_sum += (double)(-row);
  }
  for (int e = threadIdx.x; e < n * D; e += blockDim.x) {
    int r = e / D, q = e - r * D;
    float mask = y_deltas[(size_t)r * 2 * D + q];
    float tgt = y_deltas[(size_t)r * 2 * D + D + q];
    float g;
    float l = smooth_l1(tgt - deltas[e], 1.f, &g);
    reg_sum += (double)(mask * l);
    if (d_deltas) d_deltas[e] = -mask * g * inv;
  }
  double cls_total = block_sum<1024>(cls_sum, scratch);
  double reg_total = block_sum<1024>(reg_sum, scratch);
  if (threadIdx.x == 0) {
    losses_out[0] = (float)cls_total / N;
    losses_out[1] = (float)reg_total / N;
  }
}

// ---- a13: models/faster_rcnn.py:179-226, all classes in one launch -------------------------------------
// One CTA per foreground class.  Decode in fp64 exactly as the NumPy code does (centres/sizes are
// formed in fp32 and widened; the delta arithmetic is fp64 with separate mul/add roundings), clip,
// threshold, rank by score (stable, descending), then the greedy scan: the CTA walks the ranked
// boxes; for each survivor every thread tests its own boxes against it in parallel (fp64 IoU).
constexpr int kMaxDet = 512;

__global__ void __launch_bounds__(kMaxDet)
detect_postprocess_kernel(const float *__restrict__ proposals, const float *__restrict__ classes, const float *__restrict__ deltas, int n, int C,
                          double max_y, double max_x, float score_threshold, double iou_threshold, double *__restrict__ out, int32_t *__restrict__ out_counts)
{
  pdl_enter();
  __shared__ double bx[kMaxDet][4];
  __shared__ double area[kMaxDet];
  __shared__ float sc[kMaxDet];
  __shared__ int rank_of[kMaxDet];
  __shared__ int sorted[kMaxDet];
  __shared__ unsigned char dead[kMaxDet];
  __shared__ int n_sel_s, kept_s;
  const int cls = blockIdx.x + 1;
  const int i = threadIdx.x;
  bool sel = false;
  if (i == 0) { n_sel_s = 0; kept_s = 0; }
  if (i < n) {
    float4 p = __ldg(reinterpret_cast<const float4 *>(proposals) + i);
    double cy = (double)__fmul_rn(0.5f, __fadd_rn(p.x, p.z));
    double cx = (double)__fmul_rn(0.5f, __fadd_rn(p.y, p.w));
    double h = (double)__fsub_rn(p.z, p.x), w = (double)__fsub_rn(p.w, p.y);
    const float *d = deltas + (size_t)i * 4 * (C - 1) + 4 * (cls - 1);
    double dy = __dadd_rn(__dmul_rn((double)d[0], 0.1), 0.0), dx = __dadd_rn(__dmul_rn((double)d[1], 0.1), 0.0);
    double dh = __dadd_rn(__dmul_rn((double)d[2], 0.2), 0.0), dw = __dadd_rn(__dmul_rn((double)d[3], 0.2), 0.0);
    double ccy = __dadd_rn(__dmul_rn(h, dy), cy), ccx = __dadd_rn(__dmul_rn(w, dx), cx);
    double sh = __dmul_rn(h, exp(dh)), sw = __dmul_rn(w, exp(dw));
    double y1 = __dsub_rn(ccy, __dmul_rn(0.5, sh)), x1 = __dsub_rn(ccx, __dmul_rn(0.5, sw));
    double y2 = __dadd_rn(ccy, __dmul_rn(0.5, sh)), x2 = __dadd_rn(ccx, __dmul_rn(0.5, sw));
    y1 = fmin(fmax(y1, 0.0), max_y); y2 = fmin(fmax(y2, 0.0), max_y);
    x1 = fmin(fmax(x1, 0.0), max_x); x2 = fmin(fmax(x2, 0.0), max_x);
    bx[i][0] = y1; bx[i][1] = x1; bx[i][2] = y2; bx[i][3] = x2;
    area[i] = __dmul_rn(__dsub_rn(y2, y1), __dsub_rn(x2, x1));
    float s = classes[(size_t)i * C + cls];
    sc[i] = s;
    sel = s > score_threshold;
  }
  dead[i] = sel ? 0 : 1;
  __syncthreads();
  // rank among the selected (descending score, ties -> lower proposal index first: stable sort)
  int r = 0;
  if (sel) {
    float s = sc[i];
    for (int j = 0; j < n; j++)
      if (!dead[j] && (sc[j] > s || (sc[j] == s && j < i))) r++;
    rank_of[i] = r;
    sorted[r] = i;
    atomicAdd(&n_sel_s, 1);
  }
  __syncthreads();
  const int n_sel = n_sel_s;
  // greedy scan over the ranked list
  for (int q = 0; q < n_sel; q++) {
    const int cur = sorted[q];
    const bool cur_dead = dead[cur] != 0;                    // uniform: everyone reads the same flag
    __syncthreads();
    if (!cur_dead) {
      if (sel && rank_of[i] > q && !dead[i]) {
        double t0 = fmax(bx[cur][0], bx[i][0]), t1 = fmax(bx[cur][1], bx[i][1]);
        double t2 = fmin(bx[cur][2], bx[i][2]), t3 = fmin(bx[cur][3], bx[i][3]);
        double ww = fmax(0.0, __dsub_rn(t2, t0)), hh = fmax(0.0, __dsub_rn(t3, t1));
        double inter = __dmul_rn(ww, hh);
        double ovr = __ddiv_rn(inter, __dsub_rn(__dadd_rn(area[cur], area[i]), inter));
        if (ovr > iou_threshold) dead[i] = 1;
      }
      if (i == 0) {
        double *o = out + ((size_t)blockIdx.x * n + kept_s) * 5;
        o[0] = bx[cur][0]; o[1] = bx[cur][1]; o[2] = bx[cur][2]; o[3] = bx[cur][3];
        o[4] = (double)sc[cur];
        kept_s = kept_s + 1;
      }
    }
    __syncthreads();
  }
  if (i == 0) out_counts[blockIdx.x] = kept_s;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_label_proposals(const float *proposals, int n, const float *gt_boxes, const int32_t *gt_classes, int m, int num_classes,
                          float min_object_iou, float *best_iou, int32_t *class_idx, float *onehot, float *packed_targets, void *stream)
{
  FRCNN_REQUIRE(proposals && gt_boxes && gt_classes && best_iou && class_idx && onehot && packed_targets && n > 0 && m > 0 && num_classes > 1, "label_proposals: bad argument");
  launch(label_proposals_kernel, elementwise_grid(n, 128, 2), 128, 0, as_stream(stream), proposals, n, gt_boxes, gt_classes, m, num_classes, min_object_iou, best_iou, class_idx, onehot, packed_targets);
  FRCNN_CHECK_LAUNCH("label_proposals_kernel");
  return FRCNN_OK;
}

int frcnn_rpn_losses(const float *scores, const float *deltas, const float *y_true, int A, float *losses_out, float *d_scores, float *d_deltas, void *stream)
{
  FRCNN_REQUIRE(scores && deltas && y_true && losses_out && A > 0, "rpn_losses: bad argument");
  FRCNN_REQUIRE((d_scores == nullptr) == (d_deltas == nullptr), "rpn_losses: gradients must both be given or both be NULL");
  launch(rpn_losses_kernel, 1, 1024, 0, as_stream(stream), scores, deltas, y_true, A, losses_out, d_scores, d_deltas);
  FRCNN_CHECK_LAUNCH("rpn_losses_kernel");
  return FRCNN_OK;
}

int frcnn_softmax_rows(const float *logits, float *probs, int n, int C, void *stream)
{
  FRCNN_REQUIRE(logits && probs && n >= 0 && C > 0, "softmax_rows: bad argument");
  if (n == 0) return FRCNN_OK;
  launch(softmax_rows_kernel, ceil_div(n, 4), 128, 0, as_stream(stream), logits, probs, n, C);
  FRCNN_CHECK_LAUNCH("softmax_rows_kernel");
  return FRCNN_OK;
}

int frcnn_softmax_rows_bwd(const float *probs, const float *d_probs, float *d_logits, int n, int C, void *stream)
{
  FRCNN_REQUIRE(probs && d_probs && d_logits && n >= 0 && C > 0, "softmax_rows_bwd: bad argument");
  if (n == 0) return FRCNN_OK;
  launch(softmax_rows_bwd_kernel, ceil_div(n, 4), 128, 0, as_stream(stream), probs, d_probs, d_logits, n, C);
  FRCNN_CHECK_LAUNCH("softmax_rows_bwd_kernel");
  return FRCNN_OK;
}

int frcnn_sigmoid_bwd(const float *dy, const float *y, float *dz, size_t count, void *stream)
{
  FRCNN_REQUIRE(dy && y && dz, "sigmoid_bwd: null pointer");
  if (count == 0) return FRCNN_OK;
  launch(sigmoid_bwd_kernel, elementwise_grid(count, 256), 256, 0, as_stream(stream), dy, y, dz, count);
  FRCNN_CHECK_LAUNCH("sigmoid_bwd_kernel");
  return FRCNN_OK;
}

int frcnn_detector_losses(const float *probs, const float *deltas, const float *y_classes, const float *y_deltas, int n, int C,
                          float *losses_out, float *d_probs, float *d_deltas, void *stream)
{
  FRCNN_REQUIRE(probs && deltas && y_classes && y_deltas && losses_out && n > 0 && C > 1, "detector_losses: bad argument");
  launch(detector_losses_kernel, 1, 1024, 0, as_stream(stream), probs, deltas, y_classes, y_deltas, n, C, losses_out, d_probs, d_deltas);
  FRCNN_CHECK_LAUNCH("detector_losses_kernel");
  return FRCNN_OK;
}

int frcnn_iou_matrix_f32(const float *boxes1, int n, const float *boxes2, int m, float *out, void *stream)
{
  FRCNN_REQUIRE(boxes1 && boxes2 && out && n > 0 && m > 0, "iou_matrix_f32: bad argument");
  launch(iou_matrix_kernel, elementwise_grid((size_t)n * m, 256, 4), 256, 0, as_stream(stream), boxes1, n, boxes2, m, out);
  FRCNN_CHECK_LAUNCH("iou_matrix_kernel");
  return FRCNN_OK;
}

int frcnn_decode_boxes_f32(const float *deltas, const float *anchors, int n, const float *means4, const float *stds4, float *boxes, void *stream)
{
  FRCNN_REQUIRE(deltas && anchors && boxes && means4 && stds4 && n > 0, "decode_boxes_f32: bad argument");
  Vec4f mu, sd;
  for (int k = 0; k < 4; k++) { mu.v[k] = means4[k]; sd.v[k] = stds4[k]; }                                   // host arrays (4 floats each)
  launch(decode_boxes_kernel, elementwise_grid((size_t)n, 256, 4), 256, 0, as_stream(stream), deltas, anchors, n, mu, sd, boxes);
  FRCNN_CHECK_LAUNCH("decode_boxes_kernel");
  return FRCNN_OK;
}

int frcnn_detect_postprocess(const float *proposals, const float *classes, const float *deltas, int n, int C, int img_h, int img_w,
                             float score_threshold, double iou_threshold, double *out, int32_t *out_counts, void *stream)
{
  FRCNN_REQUIRE(proposals && classes && deltas && out && out_counts && n > 0 && C > 1 && img_h > 0 && img_w > 0, "detect_postprocess: bad argument");
  FRCNN_REQUIRE(n <= kMaxDet, "detect_postprocess: more than 512 proposals");
  launch(detect_postprocess_kernel, C - 1, kMaxDet, 0, as_stream(stream), proposals, classes, deltas, n, C, (double)(img_h - 1), (double)(img_w - 1), score_threshold, iou_threshold, out, out_counts);
  FRCNN_CHECK_LAUNCH("detect_postprocess_kernel");
  return FRCNN_OK;
}

}  // extern "C"
