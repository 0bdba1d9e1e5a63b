"""Record types handed from the dataset to ``FasterRCNNModel.train_step`` (reference: datasets/training_sample.py:15-39)."""
from dataclasses import dataclass
from typing import Any, List

import numpy as np


@dataclass
class Box:
  class_index: int
  class_name: str
  corners: np.ndarray                    # (y_min, x_min, y_max, x_max), fp32, pixels of the (scaled) image

  def __repr__(self):
    return "[class=%s (%f,%f,%f,%f)]" % (self.class_name, self.corners[0], self.corners[1], self.corners[2], self.corners[3])

  __str__ = __repr__


@dataclass
class TrainingSample:
  anchor_map: np.ndarray                 # (fh, fw, 36) fp32: (cy, cx, h, w) x 9 anchors per cell
  anchor_valid_map: np.ndarray           # (fh, fw, 9) fp32: 1 = anchor lies inside the image
  gt_rpn_map: np.ndarray                 # (fh, fw, 9, 6) fp32: trainable, object, ty, tx, th, tw
  gt_rpn_object_indices: Any             # (n, 3) (y, x, k) of object anchors
  gt_rpn_background_indices: Any         # (m, 3) (y, x, k) of background anchors
  gt_boxes: List[Box]                    # ground truth, scaled with the image
  image_data: np.ndarray                 # (3, H, W) fp32, preprocessed for the backbone
  image: Any                             # PIL image of the scaled picture (visualisation)
  filepath: str
