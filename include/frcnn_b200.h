/*
 * frcnn_b200.h -- C ABI of libfrcnn_sm100.so: hand-written sm_100a CUDA kernels for the
 * Faster R-CNN per-image forward/backward hot path.
 *
 * The reference (trzy/FasterRCNN) has no FFI of its own: its hot path calls third-party compiled
 * ops from Python.  Each entry point below cites the reference call site(s) it replaces
 * (paths relative to pytorch/FasterRCNN/).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - raw DEVICE pointers + explicit sizes + a cudaStream_t passed as void*; the caller
 *     allocates every output and workspace (query the *_workspace_bytes functions);
 *   - no hidden allocation, no global mutable state, never synchronises the device;
 *     stream-ordered on the given stream; re-entrant and thread-safe;
 *   - returns 0 on success, a negative FRCNN_E_* for a bad argument, a positive cudaError_t
 *     for a CUDA failure; frcnn_last_error_string() (thread-local) says what went wrong;
 *   - activations are NHWC fp32; filters are (Cout, KH, KW, Cin) fp32 ("OHWI", i.e. the
 *     reference's OIHW tensors in torch.channels_last memory format); nn.Linear weights
 *     (out, in) are the KH=KW=1 case; boxes are (y1, x1, y2, x2) as in the reference.
 */
#ifndef FRCNN_B200_H
#define FRCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define FRCNN_OK 0
#define FRCNN_E_BADARG (-1)
#define FRCNN_E_WORKSPACE (-2)
#define FRCNN_E_UNSUPPORTED (-3)

#define FRCNN_ACT_NONE 0
#define FRCNN_ACT_RELU 1
#define FRCNN_ACT_SIGMOID 2

/* math engines for the GEMM-shaped ops */
#define FRCNN_ENGINE_AUTO 0
#define FRCNN_ENGINE_SIMT_FP32 1   /* CUDA-core fp32 FMA implicit GEMM (exact fp32 products) */
#define FRCNN_ENGINE_TC_3XTF32 2   /* tcgen05 kind::tf32, error-compensated 3-product split, fp32 TMEM accumulators */
#define FRCNN_ENGINE_TC_3XF16 3    /* tcgen05 kind::f16 on per-tensor power-of-two scaled fp16 hi/lo splits (same three products, 2x the rate) */

int frcnn_version(void);
const char *frcnn_last_error_string(void);
/* Programmatic dependent launch for every kernel of the library (default: the FRCNN_PDL environment variable, off when unset):
 * each kernel may become resident while its predecessor on the stream drains and waits (griddepcontrol.wait) before its first
 * global access, which removes the launch-to-launch gap of the ~145 dependent launches of a train step.  Results do not change.
 * Returns the previous setting.  (No reference counterpart: launch plumbing.) */
int frcnn_set_pdl(int enabled);
/* SMs set aside for a concurrent collective (default 0): the persistent tcgen05 GEMM launches -- one CTA per SM, equal stream-K shares --
 * use 148 - sms CTAs while it is set, so that they stay one wave next to NCCL's resident CTAs instead of queueing a second, nearly empty
 * one behind them.  Only the CTA count changes: decomposition, workspace sizes and results are the same (the summation order inside a
 * stream-K tile follows the CTA ranges, so the fp32 rounding of a straddling tile may differ in the last bit).  Returns the previous
 * value.  Used by optim.DataParallel around the window in which gradient all-reduces overlap the backward. */
int frcnn_set_sm_reserve(int sms);

/* ---- layout ------------------------------------------------------------------------------
 * API tensors are NCHW (models/faster_rcnn.py:86-89); kernels run NHWC. */
int frcnn_nchw_to_nhwc(const float *src, float *dst, int N, int C, int H, int W, void *stream);
int frcnn_nhwc_to_nchw(const float *src, float *dst, int N, int C, int H, int W, void *stream);

/* ---- K1/K4/K8: convolution / linear as implicit GEMM ------------------------------------
 * Replaces F.relu(nn.Conv2d(.., padding="same")) models/vgg16.py:76-96, models/rpn.py:88-90
 * (act = RELU / SIGMOID / NONE epilogues), nn.Linear models/vgg16.py:129-133,
 * models/detector.py:76-78 (KH=KW=1, H=W=1, N=rows), torchvision ResNet convs
 * models/resnet.py:38-46,96 (stride 2, 1x1, 7x7; frozen BN folded into scale/bias, residual add).
 *   y[n,oh,ow,co] = act( scale[co] * sum_{kh,kw,ci} x[n,oh*s-p+kh,ow*s-p+kw,ci] * w[co,kh,kw,ci]
 *                        + bias[co] + residual[n,oh,ow,co] )
 * scale, bias, residual may be NULL.  Ho = (H + 2p - KH)/s + 1, Wo likewise. */
size_t frcnn_conv2d_fwd_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine);
int frcnn_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual,
                     float *y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                     int engine, void *workspace, size_t workspace_bytes, void *stream);

/* K2: data gradient.  dx[n,ih,iw,ci] = sum_{kh,kw,co} dy[n,oh,ow,co] * w[co,kh,kw,ci] with
 * oh*s-p+kh == ih; if addend != NULL it is added (gradient accumulation from a second
 * consumer); if scale != NULL dy is multiplied per output channel first (folded frozen BN).
 * Replaces autograd's convolution_backward (input) reached from models/faster_rcnn.py:356. */
size_t frcnn_conv2d_dgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine);
int frcnn_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx,
                       int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                       int engine, void *workspace, size_t workspace_bytes, void *stream);

/* K2: filter gradient.  dw[co,kh,kw,ci] = sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh*s-p+kh,ow*s-p+kw,ci]
 * (deterministic two-stage split-K).  Replaces convolution_backward (weight) / mm of nn.Linear. */
size_t frcnn_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine);
int frcnn_conv2d_wgrad(const float *dy, const float *x, float *dw,
                       int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                       int engine, void *workspace, size_t workspace_bytes, void *stream);

/* tcgen05 engine, operand splits shared between passes: the 3xTF32 scheme needs x = hi + lo for both operands of a
 * GEMM.  A tensor used by several passes of one step (x: forward + filter gradient; dy: data + filter gradient; w:
 * forward + data gradient) can be split ONCE with frcnn_tf32_split into a caller-owned buffer of
 * frcnn_tf32_split_bytes(count) bytes and handed to the *_presplit variants (either split pointer may be NULL = split
 * internally).  frcnn_conv2d_uses_tensor_cores(pass, geometry, engine): pass 0 = fwd, 1 = dgrad, 2 = wgrad. */
/* Debug/profiling hook (not used by the product path): when buf != NULL every tcgen05 conv CTA writes 16 x u64 %globaltimer
 * stamps (entry, setup done, first TMA, first MMA, first item issued, all MMAs issued, first item drained / stored, epilogue
 * done, exit, SM id) to buf[blockIdx * 16 ...]; buf must hold 148 * 16 * 8 bytes.  Pass NULL to switch it off again. */
void frcnn_debug_tc_trace(void *buf);
/* debug: 2-CTA clusters of the pair GEMM kernel the current device holds at once (cudaOccupancyMaxActiveClusters), -1 on error */
int frcnn_debug_pair_max_active_clusters(void);
int frcnn_conv2d_uses_tensor_cores(int pass, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int engine);
size_t frcnn_tf32_split_bytes(size_t count);
int frcnn_tf32_split(const float *x, size_t count, void *out, void *stream);
int frcnn_conv2d_fwd_presplit(const float *x, const float *w, const void *x_split, const void *w_split, const float *scale, const float *bias,
                              const float *residual, float *y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                              void *workspace, size_t workspace_bytes, void *stream);
int frcnn_conv2d_dgrad_presplit(const float *dy, const float *w, const void *dy_split, const void *w_split, const float *addend, float *dx,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                                void *workspace, size_t workspace_bytes, void *stream);
int frcnn_conv2d_wgrad_presplit(const float *dy, const float *x, const void *dy_split, const void *x_split, float *dw,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                                void *workspace, size_t workspace_bytes, void *stream);

/* fp16 engine (FRCNN_ENGINE_TC_3XF16): the same three-product scheme on kind::f16.  frcnn_f16_split writes, into a caller-owned
 * buffer of frcnn_f16_split_bytes(count) bytes, [header | hi | lo] with x * 2^e = hi + lo / 2048 (e from the tensor's absolute maximum,
 * found on the device in the same call; stored in the header and read by the GEMM kernel, no host round trip).  The *_f16 entry points
 * mirror the *_presplit ones (either split pointer may be NULL = split internally into the workspace); results are fp32 and agree with
 * the tf32 engine to fp32 rounding (tests/test_kernels_gpu.py).  Shapes: channel counts multiples of 64. */
size_t frcnn_f16_split_bytes(size_t count);
int frcnn_f16_split(const float *x, size_t count, void *out, void *stream);
/* The absolute maximum can come from the kernel that PRODUCED the tensor: the fwd / dgrad entry points below take an optional device
 * buffer y_amax / dx_amax of (16 + slots) 32-bit words, slots = frcnn_conv2d_amax_slots(pass, geometry) (0 = this shape cannot provide
 * it), into which every CTA stores the maximum |output| of its tiles (words [16, 16 + slots)); frcnn_f16_split_from_amax then splits
 * without a pass over the tensor for its maximum.  The maxima may belong to a superset of x (an un-pooled map, an unmasked gradient). */
int frcnn_conv2d_amax_slots(int pass, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int frcnn_f16_split_from_amax(const float *x, size_t count, const void *amax, int slots, void *out, void *stream);
/* One-pass split that KEEPS the exponent `out` already carries (header word 1, left there by an earlier frcnn_f16_split of the same tensor):
 * for weights after an update that moved them by a few lr * gradient (values are saturated at +-65504, so a stale exponent cannot emit
 * inf; callers refresh it with a full frcnn_f16_split every few dozen steps).  ctas_per_sm: 0 = 8; small = a side-stream launch. */
int frcnn_f16_split_carried(const float *x, size_t count, void *out, int ctas_per_sm, void *stream);
int frcnn_conv2d_fwd_f16(const float *x, const float *w, const void *x_split, const void *w_split, const float *scale, const float *bias,
                         const float *residual, float *y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int act,
                         void *y_amax, void *workspace, size_t workspace_bytes, void *stream);
int frcnn_conv2d_dgrad_f16(const float *dy, const float *w, const void *dy_split, const void *w_split, const float *addend, float *dx,
                           int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           void *dx_amax, void *workspace, size_t workspace_bytes, void *stream);
int frcnn_conv2d_wgrad_f16(const float *dy, const float *x, const void *dy_split, const void *x_split, float *dw,
                           int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           void *workspace, size_t workspace_bytes, void *stream);

/* Backward of one `y = [maxpool2x2] act(conv / linear(x, w) + b)` layer on the fp16 engine in ONE call (autograd's convolution_backward +
 * relu / max_pool2d backward reached from models/faster_rcnn.py:356): optional pooling backward into dz_full (pooled != 0: dy is the pooled
 * map's gradient, y the un-pooled activation), fused activation-backward -> dz_split (+ dbias, may be NULL), data gradient dx (may be NULL;
 * dx_amax as in frcnn_conv2d_dgrad_f16), filter gradient dw (may be NULL).  x_split / w_split / dy_amax may be NULL (split / scanned here).
 * bias_workspace >= frcnn_act_bwd_fused_workspace_bytes, workspace >= max of the dgrad / wgrad workspace queries (engine 3). */
int frcnn_conv2d_bwd_f16(const float *dy, const float *y, int act, int pooled, const float *x, const void *x_split, const float *w, const void *w_split,
                         const void *dy_amax, int dy_amax_slots, float *dz_full, void *dz_split, float *dbias, float *dx, void *dx_amax, float *dw,
                         int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                         void *bias_workspace, size_t bias_workspace_bytes, void *workspace, size_t workspace_bytes, void *stream);

/* ---- elementwise / pooling pieces of the backward pass -------------------------------------
 * dz = dy * (y > 0)   (ReLU backward, models/vgg16.py:76-96 under autograd); in place allowed. */
int frcnn_relu_bwd(const float *dy, const float *y, float *dz, size_t count, void *stream);
/* Backward of y = relu(conv * s[c] + shift[c] (+ residual)) -- a frozen BatchNorm2d evaluated as the convolution's per-channel epilogue
 * (models/resnet.py:56-77,100-107) -- over (rows, C) row-major data, C % 4 == 0:  dz = dy * (y > 0) (the residual branch's gradient;
 * NULL = not wanted),  dzs = dz * s[c] (what dgrad / wgrad consume).  y == NULL: no activation. */
int frcnn_act_bwd_scale(const float *dy, const float *y, const float *scale, float *dz, float *dzs, size_t rows, int C, void *stream);
/* dbias[c] = sum over rows of dz[row, c]; workspace >= frcnn_bias_grad_workspace_bytes. */
size_t frcnn_bias_grad_workspace_bytes(size_t rows, int C);
int frcnn_bias_grad(const float *dz, float *dbias, size_t rows, int C, void *workspace, size_t workspace_bytes, void *stream);
/* One pass for the backward of `y = act(conv/linear(x) + b)` over the (rows, C) row-major gradient: dz = dy (act NONE) or
 * y > 0 ? dy : 0 (act RELU), written as fp32 (dz, may be NULL), as the tcgen05 engine's [hi | lo] operand split (dz_split, layout of
 * frcnn_tf32_split, may be NULL) and reduced over rows into dbias (may be NULL; needs the workspace).  Replaces the autograd
 * chain relu-backward -> (operand split) -> bias.sum(0) of F.relu(nn.Conv2d) / F.relu(nn.Linear) (vgg16.py:76-96,129-133).
 * C must be a multiple of a power-of-two slab of 64..1024 channels (frcnn_act_bwd_fused_supported), else FRCNN_E_UNSUPPORTED. */
int frcnn_act_bwd_fused_supported(size_t rows, int C);
size_t frcnn_act_bwd_fused_workspace_bytes(size_t rows, int C);
int frcnn_act_bwd_fused(const float *dy, const float *y, int act, float *dz, void *dz_split, float *dbias, size_t rows, int C,
                        void *workspace, size_t workspace_bytes, void *stream);
/* fp16-engine twin: dz_split receives the frcnn_f16_split layout; its exponent comes from max |dy| (max |dz| <= max |dy| under the ReLU
 * mask): either the producer's maxima (dy_amax, dy_amax_slots: see frcnn_conv2d_amax_slots) or, when dy_amax == NULL, an amax pass over
 * dy on the same stream.  dz is never materialised in fp32 unless dz != NULL. */
int frcnn_act_bwd_fused_f16(const float *dy, const float *y, int act, float *dz, void *dz_split, float *dbias, size_t rows, int C,
                            const void *dy_amax, int dy_amax_slots, void *workspace, size_t workspace_bytes, void *stream);
/* 2x2 stride-2 max pool, floor mode (nn.MaxPool2d(2,2) models/vgg16.py:78,82,87,92), NHWC. */
int frcnn_maxpool2x2_fwd(const float *x, float *y, int N, int H, int W, int C, void *stream);
/* dz[n,h,w,c] = dy[n,h/2,w/2,c] if (h,w) is the first maximum of its window and x > 0, else 0
 * (max-pool backward fused with the ReLU backward of the conv that produced x). */
int frcnn_maxpool2x2_relu_bwd(const float *dy, const float *x, float *dz, int N, int H, int W, int C, void *stream);
/* 3x3 stride-2 pad-1 max pool (torchvision ResNet stem, models/resnet.py:42), NHWC. */
int frcnn_maxpool3x3s2_fwd(const float *x, float *y, int N, int H, int W, int C, void *stream);
/* Stride-2 helpers for the stride-1 tensor-core engine (models/resnet.py:79-81,109-118: torchvision Bottleneck, stride on conv2 and on
 * the 1x1 downsample), NHWC.  subsample2: y (N, ceil(H/2), ceil(W/2), C) = x[:, ::2, ::2, :] of x (N, H, W, C).  upsample2_zero: the
 * adjoint -- y (N, H, W, C) gets x (N, ceil(H/2), ceil(W/2), C) at the even pixels and zeros elsewhere. */
int frcnn_subsample2(const float *x, float *y, int N, int H, int W, int C, void *stream);
int frcnn_upsample2_zero(const float *x, float *y, int N, int H, int W, int C, void *stream);
/* y.mean(-1).mean(-1) of (N, H, W, C) -> (N, C) (models/resnet.py:117) and its gradient (HW = H*W). */
int frcnn_spatial_mean_fwd(const float *x, float *y, int N, int H, int W, int C, void *stream);
int frcnn_spatial_mean_bwd(const float *dy, float *dx, int N, int HW, int C, void *stream);
/* out[r][:] = x[r][:] * scale[r]: folds a frozen BatchNorm (gamma/sqrt(var+eps), models/resnet.py:56-77) into
 * the filter rows before the conv kernels, and un-folds the filter gradient afterwards. */
int frcnn_scale_rows(const float *x, const float *scale, float *out, size_t rows, size_t row_len, void *stream);
/* out = a + b (residual / gradient accumulation). */
int frcnn_add(const float *a, const float *b, float *out, size_t count, void *stream);

/* ---- K5: fused RPN anchor generation + box-delta decode + clip + min-size flag -------------
 * Replaces anchors.generate_anchor_maps (models/anchors.py:43-135, regenerated in-kernel from
 * the 9x4 fp64 template), t_convert_deltas_to_boxes (models/math_utils.py:99-128), the clamp
 * and >=16 filter (models/rpn.py:135-144).  deltas: (A,4) fp32 NHWC map flattened, A = fh*fw*9,
 * anchor index a = (y*fw + x)*9 + k.  Outputs: boxes (A,4) clipped to [0,img_h]x[0,img_w],
 * size_ok (A) uint8 = both sides >= min_size; anchors_out (A,4) (cy,cx,h,w) fp32 and valid_out
 * (A) fp32 are optional (NULL) exports of the anchor map / valid map.  anchors_in (A,4), when not
 * NULL, overrides the regenerated anchors (a caller-supplied anchor_map that differs from the
 * standard one, models/rpn.py:51,120). */
int frcnn_rpn_decode(const float *deltas, const float *anchors_in, int fh, int fw, int feature_pixels, int img_h, int img_w, float min_size,
                     float *boxes, uint8_t *size_ok, float *anchors_out, float *valid_out, void *stream);

/* ---- (f)1: RPN ground truth (anchors.generate_rpn_map, models/anchors.py:137-262) -------------------
 * anchors (A,4) fp32 (cy,cx,h,w), valid (A) fp32, gt_boxes (M,4) fp32 -> rpn_map (A,6) fp32 =
 * [trainable, object, ty, tx, th, tw]: IoU in fp64, anchor positive if IoU >= object threshold or
 * it is a best anchor of some GT box, negative if < background threshold.  workspace >= 8*M bytes. */
int frcnn_rpn_targets(const float *anchors, const float *valid, int A, const float *gt_boxes, int M, double object_iou_threshold,
                      double background_iou_threshold, float *rpn_map, void *workspace, size_t workspace_bytes, void *stream);

/* ---- top-N ordering (t.argsort + flip + [0:N], models/rpn.py:129-132) -----------------------
 * order[r] = index of the r-th best score for r < min(n, top_n); descending by score, ties ->
 * higher index first (= stable ascending argsort then flip).  If keep_mask != NULL only
 * entries with keep_mask[i] != 0 take part (allow_edge_proposals=False, models/rpn.py:170-173).
 * count_out is a device int32 array of 1 + n entries: count_out[0] receives the number of
 * entries written, count_out[1..n] is scratch for the per-element ranks. */
int frcnn_topk_order(const float *scores, const uint8_t *keep_mask, int n, int top_n, int32_t *order, int32_t *count_out, void *stream);

/* gather + compaction of the ordered, size-filtered boxes: for r in [0,*count) in order, rows
 * with size_ok[order[r]] are appended to boxes_out/scores_out; count_out = number appended. */
int frcnn_gather_filtered(const float *boxes, const float *scores, const uint8_t *size_ok, const int32_t *order,
                          const int32_t *count, int capacity, float *boxes_out, float *scores_out, int32_t *count_out, void *stream);

/* ---- K6: greedy NMS (torchvision.ops.nms; models/rpn.py:147-151) ----------------------------
 * boxes (n,4) fp32 ALREADY in descending score order (n read from the device int32 *count,
 * at most capacity); keeps box i unless an earlier kept box j has IoU(i,j) > thr (IoU in fp32,
 * compared against the double threshold exactly as the CPU op does).  keep_out receives the
 * kept POSITIONS in order, at most max_keep of them; kept_count_out their number. */
size_t frcnn_nms_workspace_bytes(int capacity);
int frcnn_nms_sorted_f32(const float *boxes, const int32_t *count, int capacity, double iou_threshold, int max_keep,
                         int32_t *keep_out, int32_t *kept_count_out, void *workspace, size_t workspace_bytes, void *stream);
/* fp64 twin for the per-class call site (models/faster_rcnn.py:216-220: float64 boxes, float32 scores already used for the
 * ordering): boxes (n,4) fp64 in descending score order; IoU and the compare in IEEE double, as torchvision's CPU op
 * instantiated for double computes them.  Same workspace, same outputs. */
int frcnn_nms_sorted_f64(const double *boxes, const int32_t *count, int capacity, double iou_threshold, int max_keep,
                         int32_t *keep_out, int32_t *kept_count_out, void *workspace, size_t workspace_bytes, void *stream);
/* Batched variant (BASELINE config 5: 6000 boxes x 20 classes; the per-class loop of models/faster_rcnn.py:196-220 at a size where
 * it is worth a grid): B independent problems of n boxes each, UNSORTED -- boxes (B,n,4) fp32, scores (B,n) fp32 -- solved in
 * one stream-ordered sequence with no host round trip: stable descending order per problem (ties -> lower index first, as
 * torchvision.ops.nms orders), bit tiles of all problems in one launch, one greedy-scan CTA per problem.  keep_out
 * (B, min(max_keep, n)) int32 = ORIGINAL indices of the kept boxes in score order, -1 padded; kept_count_out (B) int32. */
size_t frcnn_nms_batched_workspace_bytes(int B, int n, int max_keep);
int frcnn_nms_batched_f32(const float *boxes, const float *scores, int B, int n, double iou_threshold, int max_keep,
                          int32_t *keep_out, int32_t *kept_count_out, void *workspace, size_t workspace_bytes, void *stream);
/* out[r] = boxes[keep[r]] for r < *kept_count. */
int frcnn_gather_rows_f32(const float *src, int row_floats, const int32_t *index, const int32_t *count, int capacity, float *dst, void *stream);
/* dst[*dst_count + r][:] = src[r][:] for r < m (rows past dst_capacity_rows are dropped): appends the ground-truth boxes to the
 * proposal list (reference faster_rcnn.py:467 `t.vstack([proposals, gt_box_corners])`) while the proposal count is still a
 * device-side value, so labelling can be enqueued without a host round trip.  dst_count is NOT updated. */
int frcnn_append_rows_f32(float *dst, const int32_t *dst_count, int dst_capacity_rows, int row_floats, const float *src, int m, void *stream);

/* ---- K7: RoI max pooling (torchvision.ops.RoIPool((7,7), 1/16); models/detector.py:27,65-72)
 * fm NHWC (1,H,W,C); proposals (K,4) fp32 (y1,x1,y2,x2) as the reference holds them (the
 * (b,x1,y1,x2,y2) swap of detector.py:68-69 is folded in).  out (K,C,PH,PW) fp32 -- the layout
 * fc1 consumes (models/vgg16.py:129) -- argmax int32 = h*W+w or -1, laid out bin-major (K,PH*PW,C): it is private to this pair of
 * entry points (forward writes it, backward reads it), and this order is the coalesced one for both. */
int frcnn_roi_pool_fwd(const float *fm, int H, int W, int C, const float *proposals, int K, int PH, int PW, float spatial_scale,
                       float *out, int32_t *argmax, void *stream);
/* dfm (H,W,C) NHWC = scatter-add of dout through argmax; deterministic (no atomics): one
 * thread per (cell, channel) walks the RoIs in ascending order.  addend (may be NULL) is added. */
int frcnn_roi_pool_bwd(const float *dout, const int32_t *argmax, const float *proposals, int K, int H, int W, int C, int PH, int PW,
                       float spatial_scale, const float *addend, float *dfm, void *stream);

/* ---- EXTENSION (BASELINE configs 3, 5): RoIAlign.  The reference has RoIPool only; semantics are
 * torchvision.ops.roi_align's (fixed sampling_ratio in 1..4, `aligned` flag), fp32, same layouts as roi_pool.
 * Backward is deterministic (no atomics); addend (may be NULL) is added to the result. */
int frcnn_roi_align_fwd(const float *fm, int H, int W, int C, const float *proposals, int K, int PH, int PW, float spatial_scale,
                        int sampling_ratio, int aligned, float *out, void *stream);
int frcnn_roi_align_bwd(const float *dout, const float *proposals, int K, int H, int W, int C, int PH, int PW, float spatial_scale,
                        int sampling_ratio, int aligned, const float *addend, float *dfm, void *stream);

/* ---- a10: proposal labelling (FasterRCNNModel._label_proposals, models/faster_rcnn.py:418-524)
 * proposals (n,4), gt boxes (m,4), gt classes (m) int32.  For each of the n proposals: best IoU
 * (math_utils.py:39-63 semantics), class (0 if best IoU < min_object_iou), one-hot row
 * (num_classes) and packed (2, 4*(num_classes-1)) mask/target rows. */
int frcnn_label_proposals(const float *proposals, int n, const float *gt_boxes, const int32_t *gt_classes, int m, int num_classes,
                          float min_object_iou, float *best_iou, int32_t *class_idx, float *onehot, float *packed_targets, void *stream);

/* ---- a12: losses, forward value + gradients in one pass -------------------------------------
 * RPN (models/rpn.py:176-272): scores (A) post-sigmoid, deltas (A,4), y_true (A,6).
 * losses_out[0] = class loss, [1] = regression loss.  d_scores (A) = dL/d(score) (the
 * binary_cross_entropy backward), d_deltas (A,4); both NULL for value only.  Deterministic
 * single-CTA reduction. */
int frcnn_rpn_losses(const float *scores, const float *deltas, const float *y_true, int A,
                     float *losses_out, float *d_scores, float *d_deltas, void *stream);
/* dz = dy * (1 - y) * y  (t.sigmoid backward, models/rpn.py:89). */
int frcnn_sigmoid_bwd(const float *dy, const float *y, float *dz, size_t count, void *stream);
/* ---- two narrow heads on one shared input: y1 = act1(x w1^T + b1), y2 = act2(x w2^T + b2) ------------------------------
 * Replaces the pairs of third-party calls `t.sigmoid(conv1x1(y))` + `conv1x1(y)` of the RPN (models/rpn.py:89-90; x = the (H*W, 512)
 * NHWC rows of the 3x3 conv's output, outputs = the (1,H,W,9) / (1,H,W,36) maps) and `Linear` + `Linear` of the detector
 * (models/detector.py:76-78; x (N, 4096), outputs (N, 21) logits / (N, 80) box deltas) -- GEMMs too narrow (45 / 101 columns) for
 * the implicit-GEMM engines -- and their autograd backward: dx (may be NULL), dw*, db* (db may be NULL); y1 / y2 are the forward
 * outputs (needed for a sigmoid head).  x (M, K) row-major, K % 4 == 0; weights (N, K) row-major; N2 may be 0. */
size_t frcnn_heads_workspace_bytes(int M, int K, int N1, int N2);
int frcnn_heads_fwd(const float *x, int M, int K, const float *w1, const float *b1, int N1, int act1, const float *w2, const float *b2, int N2, int act2,
                    float *y1, float *y2, void *workspace, size_t workspace_bytes, void *stream);
int frcnn_heads_bwd(const float *x, int M, int K, const float *w1, int N1, int act1, const float *y1, const float *dy1,
                    const float *w2, int N2, int act2, const float *y2, const float *dy2,
                    float *dx, float *dw1, float *db1, float *dw2, float *db2, void *workspace, size_t workspace_bytes, void *stream);

/* Detector (models/detector.py:76-78,83-155): logits (n, C) -> classes = softmax (written to
 * classes_out), deltas (n, 4(C-1)), y_classes (n, C) one-hot, y_deltas (n,2,4(C-1)).
 * losses_out[0] = class loss, [1] = regression loss; d_probs (n,C) = dL/d(softmax output),
 * d_deltas (n,4(C-1)); either may be NULL. */
int frcnn_softmax_rows(const float *logits, float *probs, int n, int C, void *stream);
int frcnn_softmax_rows_bwd(const float *probs, const float *d_probs, float *d_logits, int n, int C, void *stream);
int frcnn_detector_losses(const float *probs, const float *deltas, const float *y_classes, const float *y_deltas, int n, int C,
                          float *losses_out, float *d_probs, float *d_deltas, void *stream);

/* ---- K10: fused SGD (torch.optim.SGD as configured by __main__.py:98-105) -------------------
 * g = grad*grad_scale + wd*p; buf = first_step ? g : momentum*buf + g; p -= lr*buf. */
int frcnn_sgd_step(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                   float grad_scale, int first_step, void *stream);
/* Same update; additionally writes the UPDATED weights' tf32 [hi | lo] operand split (layout of frcnn_tf32_split, param_split
 * may be NULL) so the next step's tcgen05 GEMMs do not need a separate split pass over the weights. */
int frcnn_sgd_step_split(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                         float grad_scale, int first_step, void *param_split, void *stream);

/* fp16-engine twin: param_split is a frcnn_f16_split buffer that ALREADY holds a split of these weights; the updated weights are
 * re-split with the exponent stored there (weights move by lr * update per step; the caller refreshes the exponent with a full
 * frcnn_f16_split every few dozen steps; values saturate at the fp16 maximum instead of overflowing). */
int frcnn_sgd_step_split_f16(float *param, const float *grad, float *momentum_buf, size_t count, float lr, float momentum, float weight_decay,
                             float grad_scale, int first_step, void *param_split, void *stream);

/* The whole optimizer step (optimizer.step() of models/faster_rcnn.py:359 over every parameter group) in ONE call: n tensors, host arrays
 * of device pointers / sizes / per-group hyper-parameters; param_splits[i] (may be NULL) receives the updated weights' operand split in
 * the format named by split_format (0 none, 1 tf32, 2 fp16 -- see the single-tensor entry points).  Same kernels, launched back to back. */
int frcnn_sgd_step_multi(int n, float *const *params, const float *const *grads, float *const *momentum_bufs, const size_t *counts,
                         const float *lrs, const float *momenta, const float *weight_decays, const int *first_steps, void *const *param_splits,
                         int split_format, float grad_scale, void *stream);
/* Same, with the launch shape exposed: ctas_per_sm caps the grid-stride launch (<= 0: the default 8 CTAs per SM, each looping over its
 * share).  A large value gives one short-lived 256-thread CTA per 1024 elements: the shape for updating a tensor on a side stream WHILE
 * the backward's persistent GEMM kernels run (256 threads x 32 registers fit beside a GEMM CTA; a short CTA never holds an SM back from
 * the next GEMM launch).  Used by optim.FusedSGD(eager = True). */
int frcnn_sgd_step_multi_ex(int n, float *const *params, const float *const *grads, float *const *momentum_bufs, const size_t *counts,
                            const float *lrs, const float *momenta, const float *weight_decays, const int *first_steps, void *const *param_splits,
                            int split_format, float grad_scale, int ctas_per_sm, void *stream);

/* ---- 8e: the data-parallel optimizer step as one kernel over NVLink / NVSwitch (EXPERIMENT; optim.NvlsShardedSGD) --------------------
 * Replaces, for W > 1 replicas, the pair "NCCL all-reduce of the weight gradients + optimizer.step()" (the reference has no multi-GPU
 * path; models/faster_rcnn.py:356-359 is the single-GPU backward + step it extends).  Weights and gradients of the optimizer's tensors live
 * in two flat symmetric-memory arenas with the same layout on every rank.  This rank updates the shard [shard_begin, shard_begin +
 * shard_count) (elements, multiples of 4): gradient = sum over ranks (multimem.ld_reduce on grad_multicast -- reduced inside the NVSwitch --
 * or, when the multicast pointers are NULL, loads through grad_peers[0..world)), times grad_scale; torch.optim.SGD update with the shard's
 * own momentum buffer (momentum_shard, shard_count floats); the new weights are written to every rank (multimem.st on weight_multicast, or
 * stores through weight_peers).  weight_local = this rank's own arena (read side).  The caller brackets the launch with two cross-rank
 * barriers: every rank's gradients complete before | every rank's stores delivered and gradients consumed after.  world <= 8.
 * ctas_per_sm: 256-thread CTAs per SM (0 = 8: the kernel alone on the GPU; 1 = the shape for running under the convolution backward on a side
 * stream: one such CTA fits beside a resident GEMM CTA). */
int frcnn_dp_sgd_fused(const float *grad_multicast, float *weight_multicast, const void *const *grad_peers, void *const *weight_peers, int world,
                       const float *weight_local, float *momentum_shard, size_t shard_begin, size_t shard_count,
                       float lr, float momentum, float weight_decay, float grad_scale, int first_step, int ctas_per_sm, void *stream);

/* ---- helpers under the reference's names (models/math_utils.py) -----------------------------------------------------------------
 * t_intersection_over_union (math_utils.py:39-63): boxes1 (n,4), boxes2 (m,4) fp32 (y1,x1,y2,x2) -> out (n,m): intersection (strict
 * top-left < bottom-right mask) / (area1 + area2 - intersection + 1e-7), every operation a separate fp32 rounding. */
int frcnn_iou_matrix_f32(const float *boxes1, int n, const float *boxes2, int m, float *out, void *stream);
/* t_convert_deltas_to_boxes (math_utils.py:99-128): deltas (n,4) (ty,tx,th,tw), anchors (n,4) (cy,cx,h,w), means4 / stds4 = HOST arrays of
 * four floats -> boxes (n,4) (y1,x1,y2,x2); d*std+mean, a_hw*d_yx+a_yx, a_hw*exp(d_hw), c -/+ 0.5 s with separate roundings. */
int frcnn_decode_boxes_f32(const float *deltas, const float *anchors, int n, const float *means4, const float *stds4, float *boxes, void *stream);

/* ---- a13: inference post-processing (FasterRCNNModel.predict, models/faster_rcnn.py:179-226)
 * proposals (n,4) fp32, classes (n,C) fp32, deltas (n,4(C-1)) fp32.  For every class c>=1 in one
 * launch: decode in fp64 with stds (0.1,0.1,0.2,0.2), clip to [0,img_h-1]x[0,img_w-1], keep
 * score > threshold, greedy NMS (fp64 IoU, thr iou_threshold).  out (C-1, n, 5) fp64 rows
 * (y1,x1,y2,x2,score) in kept order; out_counts (C-1) int32.  n <= 512. */
int frcnn_detect_postprocess(const float *proposals, const float *classes, const float *deltas, int n, int C, int img_h, int img_w,
                             float score_threshold, double iou_threshold, double *out, int32_t *out_counts, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FRCNN_B200_H */
