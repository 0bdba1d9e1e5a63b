"""
Geometry helper with the reference's name (pytorch/FasterRCNN/models/math_utils.py:39-63).  The
hot path does not call it (box decode is fused into ops.rpn_proposals, IoU labelling into
ops.label_proposals); it exists so code written against the reference's helper keeps working.
"""
import torch as t

from . import ops


def t_intersection_over_union(boxes1, boxes2):
  """(N,4),(M,4) CUDA fp32 -> (N,M) IoU, math_utils.py:39-63 semantics (M small: one labelling launch per column)."""
  n, m = boxes1.shape[0], boxes2.shape[0]
  out = t.empty((n, m), dtype = t.float32, device = boxes1.device)
  cls = t.zeros((1,), dtype = t.int32, device = boxes1.device)
  for j in range(m):
    best, _, _, _ = ops.label_proposals(boxes1, boxes2[j:j + 1].contiguous(), cls, 2)
    out[:, j] = best
  return out
