"""
Anchor generation (reference: pytorch/FasterRCNN/models/anchors.py).  ``generate_anchor_maps``
returns the same NumPy maps as the reference but computes them with the RPN decode kernel's
anchor generator (fp64 template, fp32 rounding points of anchors.py:118-135); inside the model
the anchors never leave the GPU.
"""
import numpy as np

from . import ops


def generate_anchor_maps(image_shape, feature_map_shape, feature_pixels):
  """-> anchor_map (H,W,36) fp32 (cy,cx,h,w)x9, anchor_valid_map (H,W,9) fp32 (NumPy, host)."""
  assert len(image_shape) == 3
  fh, fw = int(feature_map_shape[-2]), int(feature_map_shape[-1])
  anchors, valid = ops.generate_anchors_device(image_shape, (fh, fw), feature_pixels)
  return anchors.cpu().numpy(), valid.cpu().numpy()


def generate_rpn_map(anchor_map, anchor_valid_map, gt_boxes, object_iou_threshold = 0.7, background_iou_threshold = 0.3):
  """
  RPN ground truth on the GPU (reference: anchors.py:137-262).  gt_boxes: objects with ``.corners``
  (y1,x1,y2,x2) fp32.  Returns rpn_map (H,W,9,6) fp32, object indices (n,3), background indices
  (m,3) as NumPy arrays like the reference (the index lists are host data there too).
  """
  import torch as t
  from ._lib import check, lib, ptr, stream
  fh, fw, k = anchor_valid_map.shape
  a = fh * fw * k
  anchors = t.from_numpy(np.ascontiguousarray(anchor_map.reshape(-1, 4), dtype = np.float32)).cuda()
  valid = t.from_numpy(np.ascontiguousarray(anchor_valid_map.reshape(-1), dtype = np.float32)).cuda()
  gt = t.from_numpy(np.array([box.corners for box in gt_boxes], dtype = np.float32)).cuda()
  rpn_map = t.empty((a, 6), dtype = t.float32, device = "cuda")
  ws = t.empty((max(gt.shape[0], 1),), dtype = t.int64, device = "cuda")
  check(lib().frcnn_rpn_targets(ptr(anchors), ptr(valid), a, ptr(gt), gt.shape[0], float(object_iou_threshold), float(background_iou_threshold),
                                ptr(rpn_map), ptr(ws), ws.numel() * 8, stream()), "frcnn_rpn_targets")
  host = rpn_map.cpu().numpy().reshape(fh, fw, k, 6)
  obj = np.argwhere((host[:, :, :, 1] > 0) & (host[:, :, :, 0] > 0))
  bg = np.argwhere((host[:, :, :, 1] == 0) & (host[:, :, :, 0] > 0))
  return host, obj, bg
