/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the two third-party compiled ops on
 * the reference's hot path.  Never linked into, imported by or called from the product
 * (fasterrcnn_b200/); used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker / baseline.
 *
 * The arithmetic lives in torchvision (pinned torchvision==0.15.0+cu117,
 * /root/reference/pytorch/requirements.txt:8), which is not under /root/reference.  The
 * functions below restate its published CPU algorithm and are anchored on the reference's
 * call sites:
 *   nms       <- pytorch/FasterRCNN/models/rpn.py:147-151 (f32 boxes, thr 0.7)
 *                pytorch/FasterRCNN/models/faster_rcnn.py:216-220 (f64 boxes, thr 0.3)
 *   roi_pool  <- pytorch/FasterRCNN/models/detector.py:27,72 (7x7, scale 1/16) + autograd
 * Pinning: tests/test_oracle.py checks both bit-for-bit against torchvision's own CPU ops
 * (torchvision.ops.nms / roi_pool from the image's torchvision 0.26 _C.so) and against the
 * golden vectors produced by running the reference (oracle/make_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/libfrcnn_oracle.so oracle/frcnn_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ---- stable descending argsort (merge sort on indices) ------------------------------------ */
static void merge_sort_desc_f64(const double *key, int64_t *idx, int64_t *tmp, int64_t n)
{
  for (int64_t width = 1; width < n; width *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * width) {
      int64_t mid = lo + width < n ? lo + width : n;
      int64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
      int64_t a = lo, b = mid, o = lo;
      while (a < mid && b < hi) {
        /* stable: take from the left run unless the right key is strictly greater */
        if (key[idx[b]] > key[idx[a]]) tmp[o++] = idx[b++]; else tmp[o++] = idx[a++];
      }
      while (a < mid) tmp[o++] = idx[a++];
      while (b < hi) tmp[o++] = idx[b++];
    }
    memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
  }
}

/*
 * Greedy NMS.  order = stable argsort(scores, descending); area = (b2-b0)*(b3-b1) (no +1);
 * walk the order, keep i unless suppressed, then suppress every later j with
 * inter/(area_i+area_j-inter) > thr (strict; the quotient is formed in the boxes' dtype and
 * compared against the DOUBLE threshold; 0/0 = NaN never suppresses).  Returns the number kept;
 * keep[] holds original indices in score order.
 */
#define DEFINE_NMS(NAME, T)                                                                    \
  int64_t NAME(const T *boxes, const T *scores, int64_t n, double thr, int64_t *keep)          \
  {                                                                                            \
    if (n <= 0) return 0;                                                                      \
    int64_t *order = (int64_t *)malloc((size_t)n * sizeof(int64_t));                           \
    int64_t *tmp = (int64_t *)malloc((size_t)n * sizeof(int64_t));                             \
    double *key = (double *)malloc((size_t)n * sizeof(double));                                \
    T *area = (T *)malloc((size_t)n * sizeof(T));                                              \
    unsigned char *dead = (unsigned char *)calloc((size_t)n, 1);                               \
    for (int64_t i = 0; i < n; i++) {                                                          \
      order[i] = i;                                                                            \
      key[i] = (double)scores[i];                                                              \
      area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]); \
    }                                                                                          \
    merge_sort_desc_f64(key, order, tmp, n);                                                   \
    int64_t kept = 0;                                                                          \
    for (int64_t oi = 0; oi < n; oi++) {                                                       \
      int64_t i = order[oi];                                                                   \
      if (dead[i]) continue;                                                                   \
      keep[kept++] = i;                                                                        \
      T i0 = boxes[4 * i + 0], i1 = boxes[4 * i + 1], i2 = boxes[4 * i + 2],                   \
        i3 = boxes[4 * i + 3], ia = area[i];                                                   \
      for (int64_t oj = oi + 1; oj < n; oj++) {                                                \
        int64_t j = order[oj];                                                                 \
        if (dead[j]) continue;                                                                 \
        T a0 = i0 > boxes[4 * j + 0] ? i0 : boxes[4 * j + 0];                                  \
        T a1 = i1 > boxes[4 * j + 1] ? i1 : boxes[4 * j + 1];                                  \
        T a2 = i2 < boxes[4 * j + 2] ? i2 : boxes[4 * j + 2];                                  \
        T a3 = i3 < boxes[4 * j + 3] ? i3 : boxes[4 * j + 3];                                  \
        T w = a2 - a0; if (!(w > (T)0)) w = (T)0;                                              \
        T h = a3 - a1; if (!(h > (T)0)) h = (T)0;                                              \
        T inter = w * h;                                                                       \
        T ovr = inter / (ia + area[j] - inter);                                                \
        if ((double)ovr > thr) dead[j] = 1;                                                    \
      }                                                                                        \
    }                                                                                          \
    free(order); free(tmp); free(key); free(area); free(dead);                                 \
    return kept;                                                                               \
  }

DEFINE_NMS(oracle_nms_f32, float)
DEFINE_NMS(oracle_nms_f64, double)

/*
 * RoIPool forward.  input NCHW (n_img, C, H, W) f32; rois (K,5) = [batch, x1, y1, x2, y2] f32;
 * output (K, C, PH, PW) f32; argmax (K, C, PH, PW) int32 (index h*W+w inside the channel plane,
 * -1 for an empty bin).  Integer bin semantics: round() half away from zero, max(.+1,1),
 * f32 bin size, floor/ceil, clip to [0,H]/[0,W], empty -> 0, strict '>' in a row-major scan.
 */
void oracle_roi_pool_fwd(const float *input, int n_img, int C, int H, int W,
                         const float *rois, int K, int PH, int PW, float spatial_scale,
                         float *output, int32_t *argmax)
{
  (void)n_img;
  for (int n = 0; n < K; n++) {
    const float *r = rois + 5 * n;
    int b = (int)r[0];
    int xs = (int)roundf(r[1] * spatial_scale);
    int ys = (int)roundf(r[2] * spatial_scale);
    int xe = (int)roundf(r[3] * spatial_scale);
    int ye = (int)roundf(r[4] * spatial_scale);
    int rw = xe - xs + 1; if (rw < 1) rw = 1;
    int rh = ye - ys + 1; if (rh < 1) rh = 1;
    float bh = (float)rh / (float)PH;
    float bw = (float)rw / (float)PW;
    for (int ph = 0; ph < PH; ph++) {
      for (int pw = 0; pw < PW; pw++) {
        int hs = (int)floorf((float)ph * bh);
        int ws = (int)floorf((float)pw * bw);
        int he = (int)ceilf((float)(ph + 1) * bh);
        int we = (int)ceilf((float)(pw + 1) * bw);
        hs += ys; he += ys; ws += xs; we += xs;
        hs = hs < 0 ? 0 : (hs > H ? H : hs);
        he = he < 0 ? 0 : (he > H ? H : he);
        ws = ws < 0 ? 0 : (ws > W ? W : ws);
        we = we < 0 ? 0 : (we > W ? W : we);
        int empty = (he <= hs) || (we <= ws);
        for (int c = 0; c < C; c++) {
          const float *plane = input + ((size_t)b * C + c) * H * W;
          float best = empty ? 0.0f : -FLT_MAX;
          int32_t besti = -1;
          for (int h = hs; h < he; h++)
            for (int w = ws; w < we; w++) {
              float v = plane[h * W + w];
              if (v > best) { best = v; besti = h * W + w; }
            }
          size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
          output[o] = best;
          argmax[o] = besti;
        }
      }
    }
  }
}

/* RoIPool backward: grad_input[b, c, argmax] += grad_output[k, c, ph, pw] (k ascending). */
void oracle_roi_pool_bwd(const float *grad_output, const int32_t *argmax, const float *rois,
                         int K, int C, int H, int W, int PH, int PW, int n_img, float *grad_input)
{
  memset(grad_input, 0, (size_t)n_img * C * H * W * sizeof(float));
  for (int n = 0; n < K; n++) {
    int b = (int)rois[5 * n];
    for (int c = 0; c < C; c++) {
      float *plane = grad_input + ((size_t)b * C + c) * H * W;
      for (int p = 0; p < PH * PW; p++) {
        size_t o = ((size_t)n * C + c) * PH * PW + p;
        int32_t a = argmax[o];
        if (a >= 0) plane[a] += grad_output[o];
      }
    }
  }
}

/*
 * RoIAlign (EXTENSION -- the reference has no RoIAlign; SURVEY.md 8c/8f: the only available oracle is
 * torchvision.ops.roi_align, restated here for fixed sampling_ratio > 0).  input NCHW, rois (K,5) =
 * [batch, x1, y1, x2, y2], output (K,C,PH,PW).  Pinned bit-for-bit against torchvision's CPU op in tests/test_oracle.py.
 */
static void oracle_bilinear(int H, int W, float y, float x, int *yl, int *xl, int *yh, int *xh, float *w1, float *w2, float *w3, float *w4)
{
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) { *yl = -1; *w1 = *w2 = *w3 = *w4 = 0.f; *xl = *yh = *xh = 0; return; }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
  float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
  *yl = y_low; *xl = x_low; *yh = y_high; *xh = x_high;
  *w1 = hy * hx; *w2 = hy * lx; *w3 = ly * hx; *w4 = ly * lx;
}

void oracle_roi_align_fwd(const float *input, int C, int H, int W, const float *rois, int K, int PH, int PW, float scale,
                          int S, int aligned, float *output)
{
  float offset = aligned ? 0.5f : 0.0f;
  for (int n = 0; n < K; n++) {
    const float *r = rois + 5 * n;
    int b = (int)r[0];
    float sw = r[1] * scale - offset, sh = r[2] * scale - offset, ew = r[3] * scale - offset, eh = r[4] * scale - offset;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) { rw = rw > 1.f ? rw : 1.f; rh = rh > 1.f ? rh : 1.f; }
    float bh = rh / (float)PH, bw = rw / (float)PW;
    float count = (float)(S * S > 0 ? S * S : 1);
    for (int c = 0; c < C; c++) {
      const float *plane = input + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          float acc = 0.f;
          for (int iy = 0; iy < S; iy++) {
            float yy = sh + ph * bh + (iy + .5f) * bh / (float)S;
            for (int ix = 0; ix < S; ix++) {
              float xx = sw + pw * bw + (ix + .5f) * bw / (float)S;
              int yl, xl, yh, xh; float w1, w2, w3, w4;
              oracle_bilinear(H, W, yy, xx, &yl, &xl, &yh, &xh, &w1, &w2, &w3, &w4);
              if (yl < 0) continue;
              acc += w1 * plane[yl * W + xl] + w2 * plane[yl * W + xh] + w3 * plane[yh * W + xl] + w4 * plane[yh * W + xh];
            }
          }
          output[(((size_t)n * C + c) * PH + ph) * PW + pw] = acc / count;
        }
    }
  }
}

void oracle_roi_align_bwd(const float *grad_output, const float *rois, int K, int C, int H, int W, int PH, int PW, float scale,
                          int S, int aligned, int n_img, float *grad_input)
{
  memset(grad_input, 0, (size_t)n_img * C * H * W * sizeof(float));
  float offset = aligned ? 0.5f : 0.0f;
  for (int n = 0; n < K; n++) {
    const float *r = rois + 5 * n;
    int b = (int)r[0];
    float sw = r[1] * scale - offset, sh = r[2] * scale - offset, ew = r[3] * scale - offset, eh = r[4] * scale - offset;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) { rw = rw > 1.f ? rw : 1.f; rh = rh > 1.f ? rh : 1.f; }
    float bh = rh / (float)PH, bw = rw / (float)PW;
    float count = (float)(S * S);
    for (int c = 0; c < C; c++) {
      float *plane = grad_input + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          float g = grad_output[(((size_t)n * C + c) * PH + ph) * PW + pw] / count;
          for (int iy = 0; iy < S; iy++) {
            float yy = sh + ph * bh + (iy + .5f) * bh / (float)S;
            for (int ix = 0; ix < S; ix++) {
              float xx = sw + pw * bw + (ix + .5f) * bw / (float)S;
              int yl, xl, yh, xh; float w1, w2, w3, w4;
              oracle_bilinear(H, W, yy, xx, &yl, &xl, &yh, &xh, &w1, &w2, &w3, &w4);
              if (yl < 0) continue;
              plane[yl * W + xl] += g * w1; plane[yl * W + xh] += g * w2; plane[yh * W + xl] += g * w3; plane[yh * W + xh] += g * w4;
            }
          }
        }
    }
  }
}
