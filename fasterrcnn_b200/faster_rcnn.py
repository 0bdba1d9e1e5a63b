"""
Faster R-CNN model with the reference's API (pytorch/FasterRCNN/models/faster_rcnn.py):
``FasterRCNNModel(num_classes, backbone, rpn_minibatch_size, proposal_batch_size,
allow_edge_proposals)``, ``forward``, ``predict``, ``train_step`` -- same arguments, return values
and state-dict keys, so the reference's ``__main__`` drives it unchanged.  All device arithmetic
is the package's sm_100a kernels; the host keeps only what the reference keeps on the host (the
python / CPU-generator random sampling, so seeded runs pick the same samples).
"""
import random
from dataclasses import dataclass

import numpy as np
import torch as t
from torch import nn

from . import detector, ops, rpn


class FasterRCNNModel(nn.Module):
  @dataclass
  class Loss:
    rpn_class: float
    rpn_regression: float
    detector_class: float
    detector_regression: float
    total: float

  def __init__(self, num_classes, backbone, rpn_minibatch_size = 256, proposal_batch_size = 128, allow_edge_proposals = True):
    super().__init__()
    self._num_classes = num_classes
    self._rpn_minibatch_size = rpn_minibatch_size
    self._proposal_batch_size = proposal_batch_size
    self._detector_box_delta_means = [0, 0, 0, 0]
    self._detector_box_delta_stds = [0.1, 0.1, 0.2, 0.2]
    self.backbone = backbone
    self._stage1_feature_extractor = backbone.feature_extractor
    self._stage2_region_proposal_network = rpn.RegionProposalNetwork(feature_map_channels = backbone.feature_map_channels, allow_edge_proposals = allow_edge_proposals)
    self._stage3_detector_network = detector.DetectorNetwork(num_classes = num_classes, backbone = backbone)
    self.last_step_info = {}

  # ---- faster_rcnn.py:80-132 ---------------------------------------------------------------------
  def forward(self, image_data, anchor_map = None, anchor_valid_map = None):
    assert image_data.shape[0] == 1, "Batch size must be 1"
    image_shape = image_data.shape[1:]
    ops.begin_step()
    feature_map = self._stage1_feature_extractor(image_data = image_data)
    objectness_score_map, box_deltas_map, proposals = self._stage2_region_proposal_network(
      feature_map = feature_map, image_shape = image_shape, anchor_map = anchor_map, anchor_valid_map = anchor_valid_map,
      max_proposals_pre_nms = 6000, max_proposals_post_nms = 300)
    classes, box_deltas = self._stage3_detector_network(feature_map = feature_map, proposals = proposals)
    return proposals, classes, box_deltas

  # ---- faster_rcnn.py:134-226 --------------------------------------------------------------------
  def predict(self, image_data, score_threshold, anchor_map = None, anchor_valid_map = None):
    self.eval()
    assert image_data.shape[0] == 1, "Batch size must be 1"
    with t.no_grad():
      proposals, classes, box_deltas = self(image_data = image_data, anchor_map = anchor_map, anchor_valid_map = anchor_valid_map)
      # per-class fp64 decode + clip + threshold + NMS(0.3) for all classes in one launch
      return ops.detect_postprocess(proposals, classes, box_deltas, (image_data.shape[2], image_data.shape[3]), score_threshold, 0.3)

  # ---- faster_rcnn.py:228-362 --------------------------------------------------------------------
  def train_step(self, optimizer, image_data, anchor_map, anchor_valid_map, gt_rpn_map, gt_rpn_object_indices, gt_rpn_background_indices, gt_boxes):
    self.train()
    optimizer.zero_grad()
    assert image_data.shape[0] == 1, "Batch size must be 1"
    assert len(gt_rpn_map.shape) == 5 and gt_rpn_map.shape[0] == 1, "Batch size must be 1"
    assert len(gt_rpn_object_indices) == 1, "Batch size must be 1"
    assert len(gt_rpn_background_indices) == 1, "Batch size must be 1"
    assert len(gt_boxes) == 1, "Batch size must be 1"
    image_shape = image_data.shape[1:]
    ops.begin_step()

    feature_map = self._stage1_feature_extractor(image_data = image_data)
    rpn_score_map, rpn_box_deltas_map, proposals = self._stage2_region_proposal_network(
      feature_map = feature_map, image_shape = image_shape, anchor_map = anchor_map, anchor_valid_map = anchor_valid_map,
      max_proposals_pre_nms = 12000, max_proposals_post_nms = 2000)

    gt_rpn_minibatch_map = self._sample_rpn_minibatch(rpn_map = gt_rpn_map, object_indices = gt_rpn_object_indices, background_indices = gt_rpn_background_indices)
    proposals, gt_classes, gt_box_deltas = self._label_proposals(proposals = proposals, gt_boxes = gt_boxes[0], min_background_iou_threshold = 0.0, min_object_iou_threshold = 0.5)
    proposals, gt_classes, gt_box_deltas = self._sample_proposals(proposals = proposals, gt_classes = gt_classes, gt_box_deltas = gt_box_deltas, max_proposals = self._proposal_batch_size, positive_fraction = 0.25)
    proposals, gt_classes, gt_box_deltas = proposals.detach(), gt_classes.detach(), gt_box_deltas.detach()

    detector_classes, detector_box_deltas = self._stage3_detector_network(feature_map = feature_map, proposals = proposals)

    rpn_l = ops.rpn_losses(rpn_score_map, rpn_box_deltas_map, gt_rpn_minibatch_map)                # (class, regression)
    det_l = ops.detector_losses(detector_classes, detector_box_deltas, gt_classes, gt_box_deltas)    # (class, regression)
    all_l = t.cat([rpn_l, det_l])
    total_loss = all_l.sum()
    total_loss.backward()
    ops.begin_step()                       # operand splits of this step's weights are stale after the update
    optimizer.step()

    host = t.cat([all_l.detach(), total_loss.detach().reshape(1)]).cpu().numpy()                   # one D2H for the five numbers
    self.last_step_info = dict(num_rois = int(proposals.shape[0]))
    return FasterRCNNModel.Loss(rpn_class = float(host[0]), rpn_regression = float(host[1]), detector_class = float(host[2]), detector_regression = float(host[3]), total = float(host[4]))

  # ---- faster_rcnn.py:364-416 --------------------------------------------------------------------
  def _sample_rpn_minibatch(self, rpn_map, object_indices, background_indices):
    assert rpn_map.shape[0] == 1, "Batch size must be 1"
    assert len(object_indices) == 1, "Batch size must be 1"
    assert len(background_indices) == 1, "Batch size must be 1"
    positive_anchors = object_indices[0]
    negative_anchors = background_indices[0]
    assert len(positive_anchors) + len(negative_anchors) >= self._rpn_minibatch_size, "Image has insufficient anchors for RPN minibatch size of %d" % self._rpn_minibatch_size
    assert len(positive_anchors) > 0, "Image does not have any positive anchors"
    assert self._rpn_minibatch_size % 2 == 0, "RPN minibatch size must be evenly divisible"
    num_positive_samples = min(self._rpn_minibatch_size // 2, len(positive_anchors))
    num_negative_samples = self._rpn_minibatch_size - num_positive_samples
    positive_anchor_idxs = random.sample(range(len(positive_anchors)), num_positive_samples)      # same python RNG stream as the reference
    negative_anchor_idxs = random.sample(range(len(negative_anchors)), num_negative_samples)
    trainable = np.concatenate([np.asarray(positive_anchors)[positive_anchor_idxs], np.asarray(negative_anchors)[negative_anchor_idxs]])
    _, fh, fw, k, _ = rpn_map.shape
    flat = (trainable[:, 0] * fw + trainable[:, 1]) * k + trainable[:, 2]
    flat_dev = t.from_numpy(flat.astype(np.int64)).to(rpn_map.device, non_blocking = True)
    minibatch = rpn_map.clone()
    mask = minibatch.view(-1, 6)[:, 0]
    mask.zero_()
    mask[flat_dev] = 1.0
    return minibatch

  # ---- faster_rcnn.py:418-524 --------------------------------------------------------------------
  def _label_proposals(self, proposals, gt_boxes, min_background_iou_threshold, min_object_iou_threshold):
    assert min_background_iou_threshold < min_object_iou_threshold, "Object threshold must be greater than background threshold"
    dev = proposals.device
    gt_box_corners = t.from_numpy(np.array([box.corners for box in gt_boxes], dtype = np.float32)).to(dev)
    gt_box_class_idxs = t.tensor([box.class_index for box in gt_boxes], dtype = t.int32, device = dev)
    proposals = t.vstack([proposals, gt_box_corners])                                              # GT boxes join the proposals (:467)
    best_ious, class_idx, gt_classes, gt_box_deltas = ops.label_proposals(proposals, gt_box_corners, gt_box_class_idxs, self._num_classes, min_object_iou_threshold)
    if min_background_iou_threshold > 0.0:
      keep = t.where(best_ious >= min_background_iou_threshold)[0]
      proposals, class_idx, gt_classes, gt_box_deltas = proposals[keep], class_idx[keep], gt_classes[keep], gt_box_deltas[keep]
    self._last_class_idx = class_idx
    return proposals, gt_classes, gt_box_deltas

  # ---- faster_rcnn.py:526-561 --------------------------------------------------------------------
  def _sample_proposals(self, proposals, gt_classes, gt_box_deltas, max_proposals, positive_fraction):
    if max_proposals <= 0:
      return proposals, gt_classes, gt_box_deltas
    class_indices = getattr(self, "_last_class_idx", None)
    if class_indices is None or class_indices.shape[0] != gt_classes.shape[0]:
      class_indices = t.argmax(gt_classes, dim = 1)
    cls_host = class_indices.cpu().numpy()                     # the one D2H the sampling needs (counts decide the CPU-generator draws)
    positive_indices = np.where(cls_host > 0)[0]
    negative_indices = np.where(cls_host <= 0)[0]
    num_samples = min(max_proposals, len(cls_host))
    num_positive_samples = min(round(num_samples * positive_fraction), len(positive_indices))
    num_negative_samples = min(num_samples - num_positive_samples, len(negative_indices))
    if num_positive_samples <= 0 or num_negative_samples <= 0:
      return proposals[[]], gt_classes[[]], gt_box_deltas[[]]
    positive_sample_indices = positive_indices[t.randperm(len(positive_indices))[0:num_positive_samples].numpy()]   # CPU generator, as the reference
    negative_sample_indices = negative_indices[t.randperm(len(negative_indices))[0:num_negative_samples].numpy()]
    indices = t.from_numpy(np.concatenate([positive_sample_indices, negative_sample_indices]).astype(np.int64)).to(proposals.device)
    return proposals[indices], gt_classes[indices], gt_box_deltas[indices]
