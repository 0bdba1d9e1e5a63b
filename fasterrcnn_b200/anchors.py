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
