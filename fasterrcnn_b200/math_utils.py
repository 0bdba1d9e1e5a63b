"""
The reference's four geometry helpers under their own names (pytorch/FasterRCNN/models/math_utils.py:13,39,65,99), so that code written
against them keeps working when its import points at this package:

  intersection_over_union(boxes1, boxes2)                      NumPy, host   (math_utils.py:13-37; anchors.py / statistics.py call it)
  t_intersection_over_union(boxes1, boxes2)                    CUDA kernel   (math_utils.py:39-63)          frcnn_iou_matrix_f32
  convert_deltas_to_boxes(deltas, anchors, means, stds)        NumPy, host   (math_utils.py:65-97; predict's float64 decode)
  t_convert_deltas_to_boxes(deltas, anchors, means, stds)      CUDA kernel   (math_utils.py:99-128)         frcnn_decode_boxes_f32

The model's hot path does not go through them: the box decode is fused into the proposal kernels (ops.rpn_proposals), the IoU labelling
into ops.label_proposals, the per-class float64 decode into ops.detect_postprocess.  The two host functions stay NumPy because they ARE
host functions in the reference (dtype follows the inputs: float64 in predict); the two t_* functions launch this package's kernels.
"""
import numpy as np
import torch as t

from . import _lib
from ._lib import check, lib, ptr, stream


def intersection_over_union(boxes1, boxes2):
  """(N,4), (M,4) ndarrays (y1,x1,y2,x2) -> (N,M): intersection / (area1 + area2 - intersection + 1e-7); a pair counts as
  intersecting only if its top-left corner is strictly above-left of its bottom-right one."""
  b1, b2 = np.asarray(boxes1), np.asarray(boxes2)
  lo = np.maximum(b1[:, None, 0:2], b2[None, :, 0:2])
  hi = np.minimum(b1[:, None, 2:4], b2[None, :, 2:4])
  extent = hi - lo
  inter = np.all(lo < hi, axis = 2) * (extent[:, :, 0] * extent[:, :, 1])
  s1, s2 = b1[:, 2:4] - b1[:, 0:2], b2[:, 2:4] - b2[:, 0:2]
  union = (s1[:, 0] * s1[:, 1])[:, None] + (s2[:, 0] * s2[:, 1])[None, :] - inter
  return inter / (union + 1e-7)


def t_intersection_over_union(boxes1, boxes2):
  """(N,4), (M,4) CUDA fp32 tensors -> (N,M) IoU on the device, one launch (same arithmetic and roundings as the torch expression of
  math_utils.py:39-63)."""
  b1 = boxes1.detach().contiguous().float()
  b2 = boxes2.detach().contiguous().float()
  n, m = int(b1.shape[0]), int(b2.shape[0])
  out = t.empty((n, m), dtype = t.float32, device = b1.device)
  if n > 0 and m > 0:
    check(lib().frcnn_iou_matrix_f32(ptr(b1), n, ptr(b2), m, ptr(out), stream()), "frcnn_iou_matrix_f32")
    _lib.count()
  return out


def convert_deltas_to_boxes(box_deltas, anchors, box_delta_means, box_delta_stds):
  """(ty,tx,th,tw) deltas + (cy,cx,h,w) anchors -> (y1,x1,y2,x2) boxes, NumPy (result dtype as np.empty's default, float64, like the
  reference -- predict relies on it)."""
  d = np.asarray(box_deltas) * box_delta_stds + box_delta_means
  a = np.asarray(anchors)
  centre = a[:, 2:4] * d[:, 0:2] + a[:, 0:2]
  half = 0.5 * (a[:, 2:4] * np.exp(d[:, 2:4]))
  out = np.empty(d.shape)
  out[:, 0:2] = centre - half
  out[:, 2:4] = centre + half
  return out


def t_convert_deltas_to_boxes(box_deltas, anchors, box_delta_means, box_delta_stds):
  """CUDA fp32 tensors (N,4), (N,4); means / stds: 4 numbers (tensor, list or ndarray) -> (N,4) boxes on the device, one launch."""
  d = box_deltas.detach().contiguous().float()
  a = anchors.detach().contiguous().float()
  n = int(d.shape[0])
  to_host = lambda v: np.ascontiguousarray((v.detach().cpu().numpy() if isinstance(v, t.Tensor) else np.asarray(v)).astype(np.float32).reshape(4))
  mu, sd = to_host(box_delta_means), to_host(box_delta_stds)
  out = t.empty((n, 4), dtype = t.float32, device = d.device)
  if n > 0:
    check(lib().frcnn_decode_boxes_f32(ptr(d), ptr(a), n, mu.ctypes.data, sd.ctypes.data, ptr(out), stream()), "frcnn_decode_boxes_f32")
    _lib.count()
  return out
