"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the Faster R-CNN hot path.

A functional restatement (torch-CPU + NumPy + the C helpers in frcnn_oracle.c) of the
reference's per-image forward / predict / train_step, written from the reference's behaviour,
each function citing the file:line it follows (paths relative to
/root/reference/pytorch/FasterRCNN/).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(fasterrcnn_b200/) never does.

Pinning: oracle/make_golden.py executes the UNMODIFIED reference in the build container and
commits its stage-boundary tensors under tests/golden/; tests/test_oracle.py checks this
restatement against those vectors, and the C helpers against torchvision's CPU ops.

Conv / linear / pooling / softmax / autograd arithmetic is delegated to torch's CPU kernels,
exactly as the reference delegates it (models/vgg16.py:76-96, models/detector.py:76-78).
"""
import ctypes
import math
import os
import random
import subprocess
from dataclasses import dataclass

import numpy as np
import torch as t
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfrcnn_oracle.so")
_lib = None


def build_c(force = False):
  """Compiles frcnn_oracle.c (gcc, no FMA contraction) into oracle/_build/."""
  src = os.path.join(_HERE, "frcnn_oracle.c")
  if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
    return _LIB_PATH
  os.makedirs(os.path.dirname(_LIB_PATH), exist_ok = True)
  subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB_PATH, src, "-lm"])
  return _LIB_PATH


def _clib():
  global _lib
  if _lib is None:
    build_c()
    lib = ctypes.CDLL(_LIB_PATH)
    vp = ctypes.c_void_p
    lib.oracle_nms_f32.restype = ctypes.c_int64
    lib.oracle_nms_f32.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_double, vp]
    lib.oracle_nms_f64.restype = ctypes.c_int64
    lib.oracle_nms_f64.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_double, vp]
    lib.oracle_roi_pool_fwd.restype = None
    lib.oracle_roi_pool_fwd.argtypes = [vp] + [ctypes.c_int] * 4 + [vp] + [ctypes.c_int] * 3 + [ctypes.c_float, vp, vp]
    lib.oracle_roi_pool_bwd.restype = None
    lib.oracle_roi_pool_bwd.argtypes = [vp, vp, vp] + [ctypes.c_int] * 7 + [vp]
    lib.oracle_roi_align_fwd.restype = None
    lib.oracle_roi_align_fwd.argtypes = [vp] + [ctypes.c_int] * 3 + [vp] + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, ctypes.c_int, vp]
    lib.oracle_roi_align_bwd.restype = None
    lib.oracle_roi_align_bwd.argtypes = [vp, vp] + [ctypes.c_int] * 6 + [ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    _lib = lib
  return _lib


# --------------------------------------------------------------------------------------------
# Third-party op restatements (torchvision.ops.nms / RoIPool)
# --------------------------------------------------------------------------------------------

def nms(boxes, scores, iou_threshold):
  """
  Greedy NMS as torchvision.ops.nms computes it on CPU (call sites models/rpn.py:147-151 and
  models/faster_rcnn.py:216-220).  boxes (N,4) f32 or f64 ndarray/tensor, scores (N,).  Scores
  are compared after conversion to the boxes' dtype (the shimmed reference call does the same
  cast).  Returns int64 ndarray of kept indices in descending-score (stable) order.
  """
  b = boxes.detach().cpu().numpy() if isinstance(boxes, t.Tensor) else np.asarray(boxes)
  s = scores.detach().cpu().numpy() if isinstance(scores, t.Tensor) else np.asarray(scores)
  assert b.dtype in (np.float32, np.float64)
  b = np.ascontiguousarray(b)
  s = np.ascontiguousarray(s.astype(b.dtype))
  n = b.shape[0]
  keep = np.empty((max(n, 1),), dtype = np.int64)
  fn = _clib().oracle_nms_f32 if b.dtype == np.float32 else _clib().oracle_nms_f64
  k = fn(b.ctypes.data, s.ctypes.data, n, float(iou_threshold), keep.ctypes.data)
  return keep[:k].copy()


def roi_pool_forward(feature_map, rois, output_size = (7, 7), spatial_scale = 1.0 / 16.0):
  """
  torchvision.ops.RoIPool forward (call site models/detector.py:27,72).  feature_map (B,C,H,W)
  f32 ndarray, rois (K,5) = [batch, x1, y1, x2, y2] f32.  Returns (output (K,C,7,7) f32,
  argmax (K,C,7,7) int32 with -1 for empty bins).
  """
  fm = np.ascontiguousarray(feature_map, dtype = np.float32)
  r = np.ascontiguousarray(rois, dtype = np.float32)
  B, C, H, W = fm.shape
  K = r.shape[0]
  PH, PW = output_size
  out = np.empty((K, C, PH, PW), dtype = np.float32)
  arg = np.empty((K, C, PH, PW), dtype = np.int32)
  _clib().oracle_roi_pool_fwd(fm.ctypes.data, B, C, H, W, r.ctypes.data, K, PH, PW, np.float32(spatial_scale), out.ctypes.data, arg.ctypes.data)
  return out, arg


def roi_pool_backward(grad_output, argmax, rois, input_shape):
  """Gradient of roi_pool_forward w.r.t. the feature map (scatter-add to argmax)."""
  go = np.ascontiguousarray(grad_output, dtype = np.float32)
  arg = np.ascontiguousarray(argmax, dtype = np.int32)
  r = np.ascontiguousarray(rois, dtype = np.float32)
  B, C, H, W = input_shape
  K, _, PH, PW = go.shape
  gi = np.empty((B, C, H, W), dtype = np.float32)
  _clib().oracle_roi_pool_bwd(go.ctypes.data, arg.ctypes.data, r.ctypes.data, K, C, H, W, PH, PW, B, gi.ctypes.data)
  return gi


def roi_align_forward(feature_map, rois, output_size = (7, 7), spatial_scale = 1.0 / 16.0, sampling_ratio = 2, aligned = False):
  """EXTENSION oracle: torchvision.ops.roi_align (fixed sampling_ratio) restated in C; feature_map (1,C,H,W), rois (K,5)."""
  fm = np.ascontiguousarray(feature_map, dtype = np.float32)
  r = np.ascontiguousarray(rois, dtype = np.float32)
  _, C, H, W = fm.shape
  PH, PW = output_size
  out = np.empty((r.shape[0], C, PH, PW), dtype = np.float32)
  _clib().oracle_roi_align_fwd(fm.ctypes.data, C, H, W, r.ctypes.data, r.shape[0], PH, PW, np.float32(spatial_scale), int(sampling_ratio), int(aligned), out.ctypes.data)
  return out


def roi_align_backward(grad_output, rois, input_shape, spatial_scale = 1.0 / 16.0, sampling_ratio = 2, aligned = False):
  go = np.ascontiguousarray(grad_output, dtype = np.float32)
  r = np.ascontiguousarray(rois, dtype = np.float32)
  B, C, H, W = input_shape
  K, _, PH, PW = go.shape
  gi = np.empty((B, C, H, W), dtype = np.float32)
  _clib().oracle_roi_align_bwd(go.ctypes.data, r.ctypes.data, K, C, H, W, PH, PW, np.float32(spatial_scale), int(sampling_ratio), int(aligned), B, gi.ctypes.data)
  return gi


class _RoIPoolFn(t.autograd.Function):
  @staticmethod
  def forward(ctx, feature_map, rois):
    out, arg = roi_pool_forward(feature_map.detach().numpy(), rois.detach().numpy())
    ctx.save_for_backward(rois)
    ctx.arg = arg
    ctx.shape = tuple(feature_map.shape)
    return t.from_numpy(out)

  @staticmethod
  def backward(ctx, grad_output):
    (rois,) = ctx.saved_tensors
    gi = roi_pool_backward(grad_output.contiguous().numpy(), ctx.arg, rois.numpy(), ctx.shape)
    return t.from_numpy(gi), None


class _RoIAlignFn(t.autograd.Function):
  """EXTENSION oracle (no reference counterpart): torchvision.ops.roi_align semantics, fixed sampling_ratio; one image."""

  @staticmethod
  def forward(ctx, feature_map, rois, sampling_ratio, aligned):
    out = roi_align_forward(feature_map.detach().numpy(), rois.detach().numpy(), (7, 7), 1.0 / 16.0, sampling_ratio, aligned)
    ctx.save_for_backward(rois)
    ctx.cfg = (tuple(feature_map.shape), sampling_ratio, aligned)
    return t.from_numpy(out)

  @staticmethod
  def backward(ctx, grad_output):
    (rois,) = ctx.saved_tensors
    shape, s, al = ctx.cfg
    gi = roi_align_backward(grad_output.contiguous().numpy(), rois.numpy(), shape, 1.0 / 16.0, s, al)
    return t.from_numpy(gi), None, None, None


# --------------------------------------------------------------------------------------------
# Geometry (models/anchors.py, models/math_utils.py)
# --------------------------------------------------------------------------------------------

def anchor_sizes():
  """(9,2) f64 [height, width]; k = area_idx*3 + aspect_idx (models/anchors.py:25-41)."""
  out = np.empty((9, 2), dtype = np.float64)
  k = 0
  for area in (128 * 128, 256 * 256, 512 * 512):
    for aspect in (0.5, 1.0, 2.0):
      w = math.sqrt(area / aspect)
      out[k, 0] = aspect * w
      out[k, 1] = w
      k += 1
  return out


def generate_anchor_maps(image_shape, feature_map_shape, feature_pixels):
  """
  models/anchors.py:43-135.  Cell centres (cell+0.5)*stride are rounded to f32 BEFORE the f64
  template is added (anchors.py:118); validity is decided on the f64 corners; the returned
  (cy,cx,h,w) are f64 values cast to f32.  Returns anchor_map (H,W,36) f32, valid (H,W,9) f32.
  """
  assert len(image_shape) == 3
  sizes = anchor_sizes()
  template = np.concatenate([-0.5 * sizes, 0.5 * sizes], axis = 1)            # (9,4) y1,x1,y2,x2
  fh, fw = int(feature_map_shape[-2]), int(feature_map_shape[-1])
  cy = (np.arange(fh) * feature_pixels + 0.5 * feature_pixels).astype(np.float32).astype(np.float64)
  cx = (np.arange(fw) * feature_pixels + 0.5 * feature_pixels).astype(np.float32).astype(np.float64)
  centre = np.empty((fh, fw, 1, 4), dtype = np.float64)
  centre[:, :, 0, 0] = cy[:, None]
  centre[:, :, 0, 1] = cx[None, :]
  centre[:, :, 0, 2] = cy[:, None]
  centre[:, :, 0, 3] = cx[None, :]
  corners = (centre + template[None, None, :, :]).reshape(-1, 4)
  ih, iw = image_shape[1], image_shape[2]
  valid = (corners[:, 0] >= 0) & (corners[:, 1] >= 0) & (corners[:, 2] <= ih) & (corners[:, 3] <= iw)
  amap = np.empty_like(corners)
  amap[:, 0:2] = 0.5 * (corners[:, 0:2] + corners[:, 2:4])
  amap[:, 2:4] = corners[:, 2:4] - corners[:, 0:2]
  return amap.reshape(fh, fw, 36).astype(np.float32), valid.reshape(fh, fw, 9).astype(np.float32)


def iou_np(boxes1, boxes2):
  """models/math_utils.py:13-37 (strict well-ordered mask, 1e-7 in the denominator)."""
  tl = np.maximum(boxes1[:, None, 0:2], boxes2[:, 0:2])
  br = np.minimum(boxes1[:, None, 2:4], boxes2[:, 2:4])
  ok = np.all(tl < br, axis = 2)
  inter = ok * np.prod(br - tl, axis = 2)
  a1 = np.prod(boxes1[:, 2:4] - boxes1[:, 0:2], axis = 1)
  a2 = np.prod(boxes2[:, 2:4] - boxes2[:, 0:2], axis = 1)
  return inter / (a1[:, None] + a2 - inter + 1e-7)


def iou_t(boxes1, boxes2):
  """models/math_utils.py:39-63 (torch twin, f32)."""
  tl = t.maximum(boxes1[:, None, 0:2], boxes2[:, 0:2])
  br = t.minimum(boxes1[:, None, 2:4], boxes2[:, 2:4])
  ok = t.all(tl < br, dim = 2)
  inter = ok * t.prod(br - tl, dim = 2)
  a1 = t.prod(boxes1[:, 2:4] - boxes1[:, 0:2], dim = 1)
  a2 = t.prod(boxes2[:, 2:4] - boxes2[:, 0:2], dim = 1)
  return inter / (a1[:, None] + a2 - inter + 1e-7)


def deltas_to_boxes_np(box_deltas, anchors, means, stds):
  """models/math_utils.py:65-97 (NumPy; f64 when called from predict)."""
  d = box_deltas * stds + means
  c = anchors[:, 2:4] * d[:, 0:2] + anchors[:, 0:2]
  s = anchors[:, 2:4] * np.exp(d[:, 2:4])
  out = np.empty(d.shape)
  out[:, 0:2] = c - 0.5 * s
  out[:, 2:4] = c + 0.5 * s
  return out


def deltas_to_boxes_t(box_deltas, anchors, means, stds):
  """models/math_utils.py:99-128 (torch f32; separate mul / add kernels, no FMA)."""
  d = box_deltas * stds + means
  c = anchors[:, 2:4] * d[:, 0:2] + anchors[:, 0:2]
  s = anchors[:, 2:4] * t.exp(d[:, 2:4])
  out = t.empty(d.shape, dtype = t.float32)
  out[:, 0:2] = c - 0.5 * s
  out[:, 2:4] = c + 0.5 * s
  return out


def generate_rpn_map(anchor_map, anchor_valid_map, gt_corners, object_iou_threshold = 0.7, background_iou_threshold = 0.3):
  """
  models/anchors.py:137-262.  gt_corners (M,4) f32 (y1,x1,y2,x2).  Returns rpn_map (H,W,9,6)
  f32, object indices (n,3), background indices (m,3) in row-major (y,x,k) order.
  """
  fh, fw, k = anchor_valid_map.shape
  gt = np.asarray(gt_corners)
  gt_c = 0.5 * (gt[:, 0:2] + gt[:, 2:4])
  gt_s = gt[:, 2:4] - gt[:, 0:2]
  am = anchor_map.reshape(-1, 4)
  corners = np.empty(am.shape)                                               # f64 (anchors.py:185)
  corners[:, 0:2] = am[:, 0:2] - 0.5 * am[:, 2:4]
  corners[:, 2:4] = am[:, 0:2] + 0.5 * am[:, 2:4]
  n = corners.shape[0]
  ious = iou_np(corners, gt)
  ious[anchor_valid_map.reshape(-1) == 0, :] = -1.0
  best = np.max(ious, axis = 1)
  best_box = np.argmax(ious, axis = 1)
  best_per_gt = np.max(ious, axis = 0)
  top_anchor = np.where(ious == best_per_gt)[0]
  label = np.full(n, -1)
  label[best < background_iou_threshold] = 0
  label[best >= object_iou_threshold] = 1
  label[top_anchor] = 1
  enable = (label >= 0).astype(np.float32)
  label[label < 0] = 0
  targets = np.empty((n, 4))
  targets[:, 0:2] = (gt_c[best_box] - am[:, 0:2]) / am[:, 2:4]
  targets[:, 2:4] = np.log(gt_s[best_box] / am[:, 2:4])
  rpn_map = np.zeros((fh, fw, k, 6))
  rpn_map[:, :, :, 0] = anchor_valid_map * enable.reshape(fh, fw, k)
  rpn_map[:, :, :, 1] = label.reshape(fh, fw, k)
  rpn_map[:, :, :, 2:6] = targets.reshape(fh, fw, k, 4)
  obj = np.argwhere((rpn_map[:, :, :, 1] > 0) & (rpn_map[:, :, :, 0] > 0))
  bg = np.argwhere((rpn_map[:, :, :, 1] == 0) & (rpn_map[:, :, :, 0] > 0))
  return rpn_map.astype(np.float32), obj, bg


# --------------------------------------------------------------------------------------------
# Parameters (state-dict keys of the reference model)
# --------------------------------------------------------------------------------------------

VGG16_CONVS = [
  ("_block1_conv1", 3, 64), ("_block1_conv2", 64, 64),
  ("_block2_conv1", 64, 128), ("_block2_conv2", 128, 128),
  ("_block3_conv1", 128, 256), ("_block3_conv2", 256, 256), ("_block3_conv3", 256, 256),
  ("_block4_conv1", 256, 512), ("_block4_conv2", 512, 512), ("_block4_conv3", 512, 512),
  ("_block5_conv1", 512, 512), ("_block5_conv2", 512, 512), ("_block5_conv3", 512, 512),
]
VGG16_POOL_AFTER = ("_block1_conv2", "_block2_conv2", "_block3_conv3", "_block4_conv3")
VGG16_FROZEN = ("_block1_conv1", "_block1_conv2", "_block2_conv1", "_block2_conv2")   # models/vgg16.py:50-58
S1, S2, S3 = "_stage1_feature_extractor.", "_stage2_region_proposal_network.", "_stage3_detector_network."


def vgg16_param_shapes(num_classes = 21):
  """Ordered {state_dict key: shape} of the reference VGG-16 model (40 keys)."""
  shapes = {}
  for name, cin, cout in VGG16_CONVS:
    shapes[S1 + name + ".weight"] = (cout, cin, 3, 3)
    shapes[S1 + name + ".bias"] = (cout,)
  shapes[S2 + "_rpn_conv1.weight"] = (512, 512, 3, 3)
  shapes[S2 + "_rpn_conv1.bias"] = (512,)
  shapes[S2 + "_rpn_class.weight"] = (9, 512, 1, 1)
  shapes[S2 + "_rpn_class.bias"] = (9,)
  shapes[S2 + "_rpn_boxes.weight"] = (36, 512, 1, 1)
  shapes[S2 + "_rpn_boxes.bias"] = (36,)
  shapes[S3 + "_pool_to_feature_vector._fc1.weight"] = (4096, 512 * 7 * 7)
  shapes[S3 + "_pool_to_feature_vector._fc1.bias"] = (4096,)
  shapes[S3 + "_pool_to_feature_vector._fc2.weight"] = (4096, 4096)
  shapes[S3 + "_pool_to_feature_vector._fc2.bias"] = (4096,)
  shapes[S3 + "_classifier.weight"] = (num_classes, 4096)
  shapes[S3 + "_classifier.bias"] = (num_classes,)
  shapes[S3 + "_regressor.weight"] = ((num_classes - 1) * 4, 4096)
  shapes[S3 + "_regressor.bias"] = ((num_classes - 1) * 4,)
  return shapes


def synth_params(shapes, seed = 0, heads = "spread", input_scale = 50.0):
  """
  Deterministic synthetic weights (no pretrained files offline).  One CPU generator, keys in
  dict order: weights ~ N(0, sqrt(2/fan_in)) (keeps activations O(1) through 13 ReLU convs),
  biases ~ N(0, 0.01).  heads = "reference" uses the reference's head init
  (N(0,0.01)/N(0,0.001), zero bias; models/rpn.py:44-49, models/detector.py:33-36), which makes
  every objectness score ~0.5 (tie stress case); heads = "spread" scales the head weights so
  scores and classes are well separated.  The first conv is divided by input_scale (synthetic
  images are randn*50) so that activations are O(1) and SGD at lr 1e-3 is stable.
  """
  g = t.Generator(device = "cpu")
  g.manual_seed(seed)
  out = {}
  for key, shape in shapes.items():
    if key.endswith(".weight") and len(shape) > 1:
      fan_in = int(np.prod(shape[1:]))
      std = math.sqrt(2.0 / fan_in)
      if heads == "reference":
        if "_rpn_" in key or key.endswith("_classifier.weight"):
          std = 0.01
        elif key.endswith("_regressor.weight"):
          std = 0.001
      else:
        if key.endswith("_rpn_class.weight"):
          std = 0.08
        elif key.endswith("_rpn_boxes.weight"):
          std = 0.01
        elif key.endswith("_classifier.weight"):
          std = 0.02
        elif key.endswith("_regressor.weight"):
          std = 0.004
      out[key] = t.randn(shape, generator = g, dtype = t.float32) * std
      if len(shape) == 4 and shape[1] == 3:
        out[key] /= input_scale
    elif key.endswith("running_var"):
      out[key] = t.rand(shape, generator = g, dtype = t.float32) * 0.5 + 0.75
    elif key.endswith("num_batches_tracked"):
      out[key] = t.zeros(shape, dtype = t.long)
    elif key.endswith(".weight"):                     # batch-norm gamma
      out[key] = t.rand(shape, generator = g, dtype = t.float32) * 0.5 + 0.5
    else:
      std = 0.0 if (heads == "reference" and ("_rpn_" in key or "_classifier" in key or "_regressor" in key)) else 0.01
      out[key] = t.randn(shape, generator = g, dtype = t.float32) * std
  return out


def trainable_keys_vgg16(shapes):
  """Keys with requires_grad=True (everything except blocks 1-2; models/vgg16.py:50-58)."""
  return [k for k in shapes if not any((S1 + f + ".") in k for f in VGG16_FROZEN)]


def optimizer_keys(keys_with_grad):
  """__main__.py:98-105: only tensors that require grad AND have "weight" in their name."""
  return [k for k in keys_with_grad if "weight" in k]


# --------------------------------------------------------------------------------------------
# Stages
# --------------------------------------------------------------------------------------------

def vgg16_features(params, image, taps = None):
  """models/vgg16.py:60-98: 13x relu(conv3x3 'same'), 4x maxpool 2x2."""
  y = image
  for name, _, _ in VGG16_CONVS:
    y = F.relu(F.conv2d(y, params[S1 + name + ".weight"], params[S1 + name + ".bias"], stride = 1, padding = 1))
    if taps is not None:
      taps[name] = y
    if name in VGG16_POOL_AFTER:
      y = F.max_pool2d(y, kernel_size = 2, stride = 2)
  return y


def vgg16_pool_to_feature_vector(params, rois, dropout_p = 0.0, training = False):
  """models/vgg16.py:113-135 (flatten in (C,7,7) order; fc1+ReLU+dropout; fc2+ReLU+dropout)."""
  x = rois.reshape((rois.shape[0], 512 * 7 * 7))
  p = S3 + "_pool_to_feature_vector."
  y = F.dropout(F.relu(F.linear(x, params[p + "_fc1.weight"], params[p + "_fc1.bias"])), dropout_p, training)
  y = F.dropout(F.relu(F.linear(y, params[p + "_fc2.weight"], params[p + "_fc2.bias"])), dropout_p, training)
  return y


def rpn_heads(params, feature_map):
  """models/rpn.py:88-96: conv3x3+ReLU, 1x1 -> sigmoid scores, 1x1 -> deltas, NHWC maps."""
  y = F.relu(F.conv2d(feature_map, params[S2 + "_rpn_conv1.weight"], params[S2 + "_rpn_conv1.bias"], padding = 1))
  score = t.sigmoid(F.conv2d(y, params[S2 + "_rpn_class.weight"], params[S2 + "_rpn_class.bias"]))
  delta = F.conv2d(y, params[S2 + "_rpn_boxes.weight"], params[S2 + "_rpn_boxes.bias"])
  return score.permute(0, 2, 3, 1).contiguous(), delta.permute(0, 2, 3, 1).contiguous()


def rpn_proposals(score_map, delta_map, anchor_map, anchor_valid_map, image_shape, pre_nms, post_nms, allow_edge_proposals = True, taps = None):
  """
  models/rpn.py:99-156.  Decode all anchors, order by score (stable ascending argsort, then
  flip: among equal scores the HIGHER anchor index comes first), keep the first pre_nms, clip
  to [0,H]/[0,W], drop boxes with a side < 16, NMS(0.7), keep the first post_nms.
  """
  assert score_map.shape[0] == 1
  anchors = t.from_numpy(np.ascontiguousarray(anchor_map.reshape(-1, 4)))
  scores = score_map.detach().reshape(-1)
  deltas = delta_map.detach().reshape(-1, 4)
  if not allow_edge_proposals:
    ok = t.from_numpy(anchor_valid_map.reshape(-1) > 0)
    anchors, scores, deltas = anchors[ok], scores[ok], deltas[ok]
  boxes = deltas_to_boxes_t(deltas, anchors, t.tensor([0, 0, 0, 0], dtype = t.float32), t.tensor([1, 1, 1, 1], dtype = t.float32))
  order = t.argsort(scores, stable = True).flip(dims = (0,))
  boxes = boxes[order][0:pre_nms]
  scores = scores[order][0:pre_nms]
  boxes[:, 0:2] = t.clamp(boxes[:, 0:2], min = 0)
  boxes[:, 2] = t.clamp(boxes[:, 2], max = image_shape[1])
  boxes[:, 3] = t.clamp(boxes[:, 3], max = image_shape[2])
  big = t.where(((boxes[:, 2] - boxes[:, 0]) >= 16) & ((boxes[:, 3] - boxes[:, 1]) >= 16))[0]
  boxes = boxes[big]
  scores = scores[big]
  keep = t.from_numpy(nms(boxes, scores, 0.7))[0:post_nms]
  if taps is not None:
    taps.update(order = order[0:pre_nms].numpy(), size_ok = big.numpy(), nms_keep = keep.numpy(), pre_nms_boxes = boxes.numpy(), pre_nms_scores = scores.numpy())
  return boxes[keep]


def detector_forward(params, feature_map, proposals, pool_to_feature_vector):
  """models/detector.py:38-80: RoIPool 7x7 @1/16 on (b,x1,y1,x2,y2), head, softmax classes, deltas."""
  assert feature_map.shape[0] == 1
  rois = t.cat([t.zeros((proposals.shape[0], 1)), proposals], dim = 1)[:, [0, 2, 1, 4, 3]].contiguous()
  pooled = _RoIPoolFn.apply(feature_map, rois)
  y = pool_to_feature_vector(pooled)
  logits = F.linear(y, params[S3 + "_classifier.weight"], params[S3 + "_classifier.bias"])
  classes = F.softmax(logits, dim = 1)
  deltas = F.linear(y, params[S3 + "_regressor.weight"], params[S3 + "_regressor.bias"])
  return classes, deltas


def detector_forward_batch(params, feature_map, proposals_list, pool_to_feature_vector, roi_op = "pool", sampling_ratio = 2, aligned = False):
  """EXTENSION (BASELINE config 3): models/detector.py:38-80 for a batch -- the RoIs of image b are pooled from
  feature_map[b] (the batch column of detector.py:65 carries b instead of 0), stacked in image order, and go through the head
  together.  roi_op "pool" = torchvision RoIPool (the reference's), "align" = torchvision roi_align(sampling_ratio, aligned)."""
  pooled = []
  for b, proposals in enumerate(proposals_list):
    rois = t.cat([t.zeros((proposals.shape[0], 1)), proposals], dim = 1)[:, [0, 2, 1, 4, 3]].contiguous()
    fm_b = feature_map[b:b + 1]
    pooled.append(_RoIPoolFn.apply(fm_b, rois) if roi_op == "pool" else _RoIAlignFn.apply(fm_b, rois, sampling_ratio, aligned))
  y = pool_to_feature_vector(t.cat(pooled, dim = 0))
  logits = F.linear(y, params[S3 + "_classifier.weight"], params[S3 + "_classifier.bias"])
  classes = F.softmax(logits, dim = 1)
  deltas = F.linear(y, params[S3 + "_regressor.weight"], params[S3 + "_regressor.bias"])
  return classes, deltas


# --------------------------------------------------------------------------------------------
# Losses (models/rpn.py:176-272, models/detector.py:83-155)
# --------------------------------------------------------------------------------------------

def rpn_class_loss(scores, y_true):
  mask = y_true[:, :, :, :, 0].reshape(scores.shape)
  target = y_true[:, :, :, :, 1].reshape(scores.shape)
  n_cls = t.count_nonzero(mask) + 1e-7
  return t.sum(mask * F.binary_cross_entropy(scores, target, reduction = "none")) / n_cls


def _smooth_l1(x, sigma_squared):
  ax = t.abs(x)
  small = (ax < (1.0 / sigma_squared)).float()
  return small * (0.5 * x * x * sigma_squared) + (1.0 - small) * (ax - 0.5 / sigma_squared)


def rpn_regression_loss(deltas, y_true):
  target = y_true[:, :, :, :, 2:6].reshape(deltas.shape)
  included = y_true[:, :, :, :, 0].reshape(y_true.shape[0:4])
  positive = y_true[:, :, :, :, 1].reshape(y_true.shape[0:4])
  mask = (included * positive).repeat_interleave(4, dim = 3)
  n_cls = t.count_nonzero(included) + 1e-7
  return 1.0 * t.sum(mask * _smooth_l1(target - deltas, 9.0)) / n_cls


def detector_class_loss(classes, y_true):
  per_row = -(y_true * t.log(classes + 1e-7)).sum(dim = 1)
  return 1.0 * (t.sum(per_row) / (per_row.shape[0] + 1e-7))


def detector_regression_loss(deltas, y_true):
  mask = y_true[:, 0, :]
  target = y_true[:, 1, :]
  return 1.0 * t.sum(mask * _smooth_l1(target - deltas, 1.0)) / (y_true.shape[0] + 1e-7)


# --------------------------------------------------------------------------------------------
# Training-time sampling / labelling (models/faster_rcnn.py:364-561)
# --------------------------------------------------------------------------------------------

def sample_rpn_minibatch(gt_rpn_map, object_indices, background_indices, minibatch_size = 256):
  """faster_rcnn.py:364-416: <=128 positives + the rest negatives through python's random.sample."""
  assert gt_rpn_map.shape[0] == 1
  n_pos_all, n_neg_all = len(object_indices), len(background_indices)
  assert n_pos_all + n_neg_all >= minibatch_size
  assert n_pos_all > 0
  assert minibatch_size % 2 == 0
  n_pos = min(minibatch_size // 2, n_pos_all)
  n_neg = minibatch_size - n_pos
  pos_pick = random.sample(range(n_pos_all), n_pos)
  neg_pick = random.sample(range(n_neg_all), n_neg)
  chosen = np.concatenate([object_indices[pos_pick], background_indices[neg_pick]])
  out = gt_rpn_map.clone()
  out[:, :, :, :, 0] = 0
  out[0, chosen[:, 0], chosen[:, 1], chosen[:, 2], 0] = 1
  return out


def label_proposals(proposals, gt_corners, gt_class_idxs, num_classes = 21, min_background_iou = 0.0, min_object_iou = 0.5):
  """
  faster_rcnn.py:418-524.  Appends the GT boxes as proposals; IoU vs every GT box; best GT per
  proposal; label 0 below min_object_iou; one-hot classes; targets
  ((gt_c - p_c)/p_s, log(gt_s/p_s)) / (0.1,0.1,0.2,0.2) packed as (N,2,80) = (mask, targets).
  """
  assert min_background_iou < min_object_iou
  gt = t.from_numpy(np.asarray(gt_corners, dtype = np.float32))
  cls = t.tensor(list(gt_class_idxs), dtype = t.long)
  props = t.vstack([proposals, gt])
  ious = iou_t(props, gt)
  best = t.max(ious, dim = 1).values
  which = t.argmax(ious, dim = 1)
  cls = cls[which]
  gtb = gt[which]
  keep = t.where(best >= min_background_iou)[0]
  props, best, cls, gtb = props[keep], best[keep], cls[keep], gtb[keep]
  cls[best < min_object_iou] = 0
  n = props.shape[0]
  onehot = t.zeros((n, num_classes), dtype = t.float32)
  onehot[t.arange(n), cls] = 1.0
  pc = 0.5 * (props[:, 0:2] + props[:, 2:4])
  ps = props[:, 2:4] - props[:, 0:2]
  gc = 0.5 * (gtb[:, 0:2] + gtb[:, 2:4])
  gs = gtb[:, 2:4] - gtb[:, 0:2]
  tg = t.empty((n, 4), dtype = t.float32)
  tg[:, 0:2] = (gc - pc) / ps
  tg[:, 2:4] = t.log(gs / ps)
  tg[:, :] -= t.tensor([0, 0, 0, 0], dtype = t.float32)
  tg[:, :] /= t.tensor([0.1, 0.1, 0.2, 0.2], dtype = t.float32)
  packed = t.zeros((n, 2, 4 * (num_classes - 1)), dtype = t.float32)
  packed[:, 0, :] = t.repeat_interleave(onehot, repeats = 4, dim = 1)[:, 4:]
  packed[:, 1, :] = t.tile(tg, dims = (1, num_classes - 1))
  return props, onehot, packed


def sample_proposals(proposals, gt_classes, gt_box_deltas, max_proposals = 128, positive_fraction = 0.25):
  """faster_rcnn.py:526-561: round(n*0.25) positives max, rest negatives, t.randperm (CPU generator)."""
  if max_proposals <= 0:
    return proposals, gt_classes, gt_box_deltas
  cls = t.argmax(gt_classes, dim = 1)
  pos = t.where(cls > 0)[0]
  neg = t.where(cls <= 0)[0]
  n = min(max_proposals, len(cls))
  n_pos = min(round(n * positive_fraction), len(pos))
  n_neg = min(n - n_pos, len(neg))
  if n_pos <= 0 or n_neg <= 0:
    return proposals[[]], gt_classes[[]], gt_box_deltas[[]]
  pos_pick = pos[t.randperm(len(pos))[0:n_pos]]
  neg_pick = neg[t.randperm(len(neg))[0:n_neg]]
  idx = t.cat([pos_pick, neg_pick])
  return proposals[idx], gt_classes[idx], gt_box_deltas[idx]


# --------------------------------------------------------------------------------------------
# Model-level entry points (models/faster_rcnn.py:80-362)
# --------------------------------------------------------------------------------------------

@dataclass
class Loss:
  rpn_class: float
  rpn_regression: float
  detector_class: float
  detector_regression: float
  total: float


class OracleModel:
  """Functional twin of FasterRCNNModel (faster_rcnn.py:27-78) around a parameter dict."""

  def __init__(self, params, backbone = "vgg16", num_classes = 21, rpn_minibatch_size = 256, proposal_batch_size = 128, allow_edge_proposals = True, dropout_probability = 0.0):
    assert backbone == "vgg16" or backbone.startswith("resnet")
    self.backbone = backbone
    self.num_classes = num_classes
    self.rpn_minibatch_size = rpn_minibatch_size
    self.proposal_batch_size = proposal_batch_size
    self.allow_edge_proposals = allow_edge_proposals
    self.dropout_probability = dropout_probability
    self.params = {k: v.clone() for k, v in params.items()}
    if backbone == "vgg16":
      trainable = set(trainable_keys_vgg16(self.params))
    else:
      from . import resnet_oracle
      trainable = set(resnet_oracle.trainable_keys(self.params))
    for k, v in self.params.items():
      if v.dtype.is_floating_point:
        v.requires_grad_(k in trainable)
    self.optimizer_keys = optimizer_keys([k for k in self.params if self.params[k].requires_grad])
    self.momentum_buffers = {}
    self.training = False

  # -- backbone protocol (models/backbone.py:30-65) --
  @property
  def feature_pixels(self):
    return 16

  def compute_feature_map_shape(self, image_shape):
    if self.backbone == "vgg16":                                            # vgg16.py:155-158
      return (512, image_shape[-2] // 16, image_shape[-1] // 16)
    return (1024, math.ceil(image_shape[-2] / 16), math.ceil(image_shape[-1] / 16))   # resnet.py:183-185

  def features(self, image, taps = None):
    if self.backbone == "vgg16":
      return vgg16_features(self.params, image, taps)
    from . import resnet_oracle
    return resnet_oracle.features(self.params, image, self.backbone)

  def pool_to_feature_vector(self, rois):
    if self.backbone == "vgg16":
      return vgg16_pool_to_feature_vector(self.params, rois, self.dropout_probability, self.training)
    from . import resnet_oracle
    return resnet_oracle.pool_to_feature_vector(self.params, rois, self.backbone)

  # -- faster_rcnn.py:80-132 --
  def forward(self, image, anchor_map = None, anchor_valid_map = None, taps = None):
    assert image.shape[0] == 1, "Batch size must be 1"
    image_shape = tuple(image.shape[1:])
    if anchor_map is None or anchor_valid_map is None:
      anchor_map, anchor_valid_map = generate_anchor_maps(image_shape, self.compute_feature_map_shape(image_shape), self.feature_pixels)
    fm = self.features(image, taps)
    score_map, delta_map = rpn_heads(self.params, fm)
    proposals = rpn_proposals(score_map, delta_map, anchor_map, anchor_valid_map, image_shape, 6000, 300, self.allow_edge_proposals, taps)
    classes, deltas = detector_forward(self.params, fm, proposals, self.pool_to_feature_vector)
    if taps is not None:
      taps.update(feature_map = fm, score_map = score_map, delta_map = delta_map)
    return proposals, classes, deltas

  # -- faster_rcnn.py:134-226 --
  def predict(self, image, score_threshold, anchor_map = None, anchor_valid_map = None):
    self.training = False
    with t.no_grad():
      proposals, classes, deltas = self.forward(image, anchor_map, anchor_valid_map)
    proposals, classes, deltas = proposals.numpy(), classes.numpy(), deltas.numpy()
    pa = np.empty(proposals.shape)                                            # f64 (faster_rcnn.py:180)
    pa[:, 0] = 0.5 * (proposals[:, 0] + proposals[:, 2])
    pa[:, 1] = 0.5 * (proposals[:, 1] + proposals[:, 3])
    pa[:, 2:4] = proposals[:, 2:4] - proposals[:, 0:2]
    result = {}
    for c in range(1, classes.shape[1]):
      boxes = deltas_to_boxes_np(deltas[:, (c - 1) * 4:(c - 1) * 4 + 4], pa, [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2])
      boxes[:, 0::2] = np.clip(boxes[:, 0::2], 0, image.shape[2] - 1)
      boxes[:, 1::2] = np.clip(boxes[:, 1::2], 0, image.shape[3] - 1)
      sc = classes[:, c]
      sel = np.where(sc > score_threshold)[0]
      boxes, sc = boxes[sel], sc[sel]
      keep = nms(boxes, sc, 0.3)
      result[c] = np.hstack([boxes[keep], sc[keep][:, None]])
    return result

  # -- faster_rcnn.py:228-362 + __main__.py:98-105 (SGD momentum, weight decay on weights only) --
  def train_step(self, image, anchor_map, anchor_valid_map, gt_rpn_map, gt_rpn_object_indices, gt_rpn_background_indices, gt_corners, gt_class_idxs, lr = 1e-3, momentum = 0.9, weight_decay = 5e-4, apply_update = True, taps = None):
    self.training = True
    for v in self.params.values():
      v.grad = None
    assert image.shape[0] == 1, "Batch size must be 1"
    image_shape = tuple(image.shape[1:])
    fm = self.features(image, taps)
    score_map, delta_map = rpn_heads(self.params, fm)
    proposals = rpn_proposals(score_map, delta_map, anchor_map, anchor_valid_map, image_shape, 12000, 2000, self.allow_edge_proposals, taps)
    minibatch_map = sample_rpn_minibatch(gt_rpn_map, gt_rpn_object_indices, gt_rpn_background_indices, self.rpn_minibatch_size)
    props, gt_classes, gt_deltas = label_proposals(proposals, gt_corners, gt_class_idxs, self.num_classes)
    if taps is not None:
      taps.update(all_proposals = proposals.numpy().copy(), labelled_proposals = props.numpy().copy(), labelled_classes = gt_classes.numpy().copy(), labelled_deltas = gt_deltas.numpy().copy(), minibatch_mask = minibatch_map[0, :, :, :, 0].numpy().copy())
    props, gt_classes, gt_deltas = sample_proposals(props, gt_classes, gt_deltas, self.proposal_batch_size, 0.25)
    props, gt_classes, gt_deltas = props.detach(), gt_classes.detach(), gt_deltas.detach()
    classes, deltas = detector_forward(self.params, fm, props, self.pool_to_feature_vector)
    l1 = rpn_class_loss(score_map, minibatch_map)
    l2 = rpn_regression_loss(delta_map, minibatch_map)
    l3 = detector_class_loss(classes, gt_classes)
    l4 = detector_regression_loss(deltas, gt_deltas)
    total = l1 + l2 + l3 + l4
    loss = Loss(l1.item(), l2.item(), l3.item(), l4.item(), total.item())
    total.backward()
    if taps is not None:
      taps.update(feature_map = fm.detach(), score_map = score_map.detach(), delta_map = delta_map.detach(), sampled_proposals = props, sampled_classes = gt_classes, sampled_deltas = gt_deltas, classes = classes.detach(), deltas = deltas.detach())
    if apply_update:
      self.sgd_step(lr, momentum, weight_decay)
    return loss

  # -- EXTENSION: batch > 1 (no reference behaviour; the reference's batch-general loss formulas applied to stacked tensors) --
  def forward_batch(self, images, roi_op = "pool", sampling_ratio = 2, aligned = False, pre_nms = 6000, post_nms = 300):
    """images (B,3,H,W) -> [(proposals_b, classes_b, deltas_b)]: shared-size images, per-image proposals, one head pass."""
    image_shape = tuple(images.shape[1:])
    anchor_map, anchor_valid_map = generate_anchor_maps(image_shape, self.compute_feature_map_shape(image_shape), self.feature_pixels)
    fm = self.features(images)
    score_map, delta_map = rpn_heads(self.params, fm)
    props = [rpn_proposals(score_map[b:b + 1], delta_map[b:b + 1], anchor_map, anchor_valid_map, image_shape, pre_nms, post_nms, self.allow_edge_proposals) for b in range(images.shape[0])]
    classes, deltas = detector_forward_batch(self.params, fm, props, self.pool_to_feature_vector, roi_op, sampling_ratio, aligned)
    out, at = [], 0
    for p in props:
      out.append((p, classes[at:at + p.shape[0]], deltas[at:at + p.shape[0]]))
      at += p.shape[0]
    return out

  def train_step_batch(self, images, samples, roi_op = "pool", sampling_ratio = 2, aligned = False, lr = 1e-3, momentum = 0.9, weight_decay = 5e-4, apply_update = True):
    """One SGD step on a batch: samples[b] = dict(anchor_map, anchor_valid_map, gt_rpn_map (1,h,w,9,6), gt_rpn_object_indices,
    gt_rpn_background_indices, gt_corners, gt_class_idxs).  RNG order: every image's RPN minibatch (python random) in image order,
    then every image's proposal sample (torch CPU generator) in image order.  Losses = rpn.py:176-272 / detector.py:83-155 on the
    stacked maps / stacked RoIs (their normalisers count over the whole batch)."""
    self.training = True
    for v in self.params.values():
      v.grad = None
    image_shape = tuple(images.shape[1:])
    fm = self.features(images)
    score_map, delta_map = rpn_heads(self.params, fm)
    minibatch = t.cat([sample_rpn_minibatch(s["gt_rpn_map"], s["gt_rpn_object_indices"], s["gt_rpn_background_indices"], self.rpn_minibatch_size) for s in samples], dim = 0)
    props_l, cls_l, dlt_l = [], [], []
    for b, s in enumerate(samples):
      proposals = rpn_proposals(score_map[b:b + 1], delta_map[b:b + 1], s["anchor_map"], s["anchor_valid_map"], image_shape, 12000, 2000, self.allow_edge_proposals)
      props, gt_classes, gt_deltas = label_proposals(proposals, s["gt_corners"], s["gt_class_idxs"], self.num_classes)
      props, gt_classes, gt_deltas = sample_proposals(props, gt_classes, gt_deltas, self.proposal_batch_size, 0.25)
      props_l.append(props.detach()); cls_l.append(gt_classes.detach()); dlt_l.append(gt_deltas.detach())
    classes, deltas = detector_forward_batch(self.params, fm, props_l, self.pool_to_feature_vector, roi_op, sampling_ratio, aligned)
    l1 = rpn_class_loss(score_map, minibatch)
    l2 = rpn_regression_loss(delta_map, minibatch)
    l3 = detector_class_loss(classes, t.cat(cls_l))
    l4 = detector_regression_loss(deltas, t.cat(dlt_l))
    total = l1 + l2 + l3 + l4
    loss = Loss(l1.item(), l2.item(), l3.item(), l4.item(), total.item())
    total.backward()
    self.last_batch_rois = [p.shape[0] for p in props_l]
    if apply_update:
      self.sgd_step(lr, momentum, weight_decay)
    return loss

  def sgd_step(self, lr, momentum, weight_decay):
    """torch.optim.SGD as configured by __main__.py:98-105 (dampening 0, no nesterov)."""
    with t.no_grad():
      for k in self.optimizer_keys:
        p = self.params[k]
        if p.grad is None:
          continue
        g = p.grad + weight_decay * p
        if k not in self.momentum_buffers:
          self.momentum_buffers[k] = g.clone()
        else:
          self.momentum_buffers[k].mul_(momentum).add_(g)
        p.add_(self.momentum_buffers[k], alpha = -lr)


def synthetic_sample(image_hw = (600, 1000), seed = 0, backbone = "vgg16", gt = None):
  """
  The synthetic training sample of SURVEY.md section 8d: image randn*50, two VOC-shaped GT boxes,
  anchors / RPN ground truth generated by the restated reference functions.
  """
  g = t.Generator(device = "cpu")
  g.manual_seed(seed)
  h, w = image_hw
  image = t.randn((1, 3, h, w), generator = g, dtype = t.float32) * 50.0
  if gt is None:
    sy, sx = h / 600.0, w / 1000.0
    gt = [((100 * sy, 150 * sx, 400 * sy, 600 * sx), 7), ((50 * sy, 650 * sx, 500 * sy, 850 * sx), 15)]
  gt_corners = np.array([b for b, _ in gt], dtype = np.float32)
  gt_classes = [c for _, c in gt]
  if backbone == "vgg16":
    fm_shape = (512, h // 16, w // 16)
  else:
    fm_shape = (1024, math.ceil(h / 16), math.ceil(w / 16))
  amap, avalid = generate_anchor_maps((3, h, w), fm_shape, 16)
  rpn_map, obj, bg = generate_rpn_map(amap, avalid, gt_corners)
  return dict(image = image, anchor_map = amap, anchor_valid_map = avalid, gt_rpn_map = t.from_numpy(rpn_map).unsqueeze(0), gt_rpn_object_indices = obj, gt_rpn_background_indices = bg, gt_corners = gt_corners, gt_class_idxs = gt_classes)
