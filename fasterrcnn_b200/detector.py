"""
Detector stage (reference: pytorch/FasterRCNN/models/detector.py).  RoIPool 7x7 @ 1/16 ->
backbone.pool_to_feature_vector -> Linear + softmax (classes), Linear (box deltas); every op is
an sm_100a kernel (ops.roi_pool, ops.linear_act, ops.softmax_rows).
"""
import torch as t
from torch import nn

from . import ops
from .backbone import LinearParams


class DetectorNetwork(nn.Module):
  def __init__(self, num_classes, backbone, roi_op = "pool", roi_sampling_ratio = 2, roi_aligned = False):
    super().__init__()
    assert roi_op in ("pool", "align"), "roi_op must be 'pool' (the reference's RoIPool) or 'align' (extension: torchvision roi_align semantics)"
    self._roi_op, self._roi_sampling_ratio, self._roi_aligned = roi_op, roi_sampling_ratio, roi_aligned
    self._input_features = 7 * 7 * backbone.feature_map_channels
    self._spatial_scale = 1.0 / backbone.feature_pixels
    self._pool_to_feature_vector = backbone.pool_to_feature_vector
    self._classifier = LinearParams(backbone.feature_vector_size, num_classes)
    self._regressor = LinearParams(backbone.feature_vector_size, (num_classes - 1) * 4)
    self._classifier.weight.data.normal_(mean = 0.0, std = 0.01)      # detector.py:33-36
    self._classifier.bias.data.zero_()
    self._regressor.weight.data.normal_(mean = 0.0, std = 0.001)
    self._regressor.bias.data.zero_()

  def forward(self, feature_map, proposals):
    """feature_map (1,C,H,W), proposals (N,4) (y1,x1,y2,x2) -> classes (N,num_classes), box deltas (N,4(num_classes-1))."""
    assert feature_map.shape[0] == 1, "Batch size must be 1"
    rois = self._roi(feature_map, proposals)
    y = self._pool_to_feature_vector(rois = rois)
    # class logits (21) and box deltas (80) are one narrow GEMM over the feature vectors
    classes_raw, box_deltas = ops.two_heads(y, self._classifier.weight, self._classifier.bias, ops.ACT_NONE, self._regressor.weight, self._regressor.bias, ops.ACT_NONE)
    classes = ops.softmax_rows(classes_raw)
    return classes, box_deltas


  def _roi(self, feature_map, proposals):
    if self._roi_op == "pool":
      return ops.roi_pool(feature_map, proposals, (7, 7), self._spatial_scale)
    return ops.roi_align(feature_map, proposals, (7, 7), self._spatial_scale, self._roi_sampling_ratio, self._roi_aligned)

  def forward_batch(self, feature_map, proposals_list):
    """EXTENSION (batch > 1): the RoIs of image b are pooled from feature_map[b] (detector.py:65's batch column carrying b),
    stacked in image order and sent through the head in one pass -> classes (sum N_b, C), box deltas (sum N_b, 4(C-1))."""
    assert feature_map.shape[0] == len(proposals_list)
    pooled = [self._roi(feature_map[b:b + 1], p) for b, p in enumerate(proposals_list)]
    y = self._pool_to_feature_vector(rois = t.cat(pooled, dim = 0) if len(pooled) > 1 else pooled[0])
    classes_raw, box_deltas = ops.two_heads(y, self._classifier.weight, self._classifier.bias, ops.ACT_NONE, self._regressor.weight, self._regressor.bias, ops.ACT_NONE)
    return ops.softmax_rows(classes_raw), box_deltas


def class_loss(predicted_classes, y_true):
  """detector.py:83-104."""
  n, c = predicted_classes.shape
  d = 4 * (c - 1)
  zeros = t.zeros((n, d), dtype = t.float32, device = predicted_classes.device)
  return ops.detector_losses(predicted_classes, zeros, y_true, t.zeros((n, 2, d), dtype = t.float32, device = predicted_classes.device))[0]


def regression_loss(predicted_box_deltas, y_true):
  """detector.py:106-155."""
  n, d = predicted_box_deltas.shape
  c = d // 4 + 1
  probs = t.full((n, c), 1.0 / c, dtype = t.float32, device = predicted_box_deltas.device)
  return ops.detector_losses(probs, predicted_box_deltas, t.zeros((n, c), dtype = t.float32, device = predicted_box_deltas.device), y_true)[1]
