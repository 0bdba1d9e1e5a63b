"""
Region proposal network (reference: pytorch/FasterRCNN/models/rpn.py).  3x3 conv 512->512 + ReLU,
1x1 -> 9 sigmoid objectness scores, 1x1 -> 36 box deltas, all through the implicit-GEMM kernels
whose NHWC output IS the (1,H,W,9)/(1,H,W,36) map layout the reference permutes into
(rpn.py:95-96); proposals through ops.rpn_proposals (decode+anchors, top-N, size filter, NMS).
"""
import hashlib

import numpy as np
import torch as t
from torch import nn

from . import ops
from .backbone import ConvParams


class RegionProposalNetwork(nn.Module):
  def __init__(self, feature_map_channels, allow_edge_proposals = False):
    super().__init__()
    self._allow_edge_proposals = allow_edge_proposals
    num_anchors = 9
    c = feature_map_channels
    self._rpn_conv1 = ConvParams(c, c, (3, 3))
    self._rpn_class = ConvParams(c, num_anchors, (1, 1))
    self._rpn_boxes = ConvParams(c, num_anchors * 4, (1, 1))
    for layer in (self._rpn_conv1, self._rpn_class, self._rpn_boxes):     # rpn.py:44-49
      layer.weight.data.normal_(mean = 0.0, std = 0.01)
      layer.bias.data.zero_()
    self._anchor_cache = {}

  def forward(self, feature_map, image_shape, anchor_map, anchor_valid_map, max_proposals_pre_nms, max_proposals_post_nms, deferred_extra_rows = None,
              proposal_stream = None):
    """-> objectness (1,H,W,9), box deltas (1,H,W,36), proposals (N,4) (y1,x1,y2,x2).
    deferred_extra_rows (not in the reference; used by FasterRCNNModel.train_step): when an int, the third result is the pair
    (capacity-padded proposals with that many spare rows, device-side count) and no host synchronisation happens here.
    proposal_stream (train_step): the proposal kernels -- a serial, latency-bound chain of small launches -- are queued on that stream
    behind the head outputs, so that the caller can run the RPN losses and the RPN branch's backward on the compute stream meanwhile."""
    assert feature_map.shape[0] == 1                                      # rpn.py:159
    y = ops.conv2d_act(feature_map, self._rpn_conv1.weight, self._rpn_conv1.bias, 1, 1, ops.ACT_RELU)
    # the two 1x1 heads (9 sigmoid scores, 36 deltas) are one narrow GEMM over the pixels' 512-channel rows; its row-major
    # outputs ARE the (1,H,W,9) / (1,H,W,36) maps the reference permutes into (rpn.py:95-96)
    scores, deltas = ops.two_heads(y, self._rpn_class.weight, self._rpn_class.bias, ops.ACT_SIGMOID, self._rpn_boxes.weight, self._rpn_boxes.bias, ops.ACT_NONE)
    fh, fw = int(y.shape[2]), int(y.shape[3])
    objectness_score_map = scores.view(1, fh, fw, scores.shape[1])
    box_deltas_map = deltas.view(1, fh, fw, deltas.shape[1])

    anchors_dev, keep_mask = self._resolve_anchors(anchor_map, anchor_valid_map, image_shape, objectness_score_map.shape[1:3], feature_map.device)
    if proposal_stream is not None:
      heads_done = t.cuda.Event()
      heads_done.record()
      with t.cuda.stream(proposal_stream):
        proposal_stream.wait_event(heads_done)
        proposals = ops.rpn_proposals(
          objectness_score_map, box_deltas_map, image_shape, 16,
          max_proposals_pre_nms, max_proposals_post_nms, anchors = anchors_dev, keep_mask = keep_mask,
          defer_count = deferred_extra_rows is not None, extra_rows = deferred_extra_rows or 0)
      return objectness_score_map, box_deltas_map, proposals
    proposals = ops.rpn_proposals(
      objectness_score_map, box_deltas_map, image_shape, 16,
      max_proposals_pre_nms, max_proposals_post_nms, anchors = anchors_dev, keep_mask = keep_mask,
      defer_count = deferred_extra_rows is not None, extra_rows = deferred_extra_rows or 0)
    return objectness_score_map, box_deltas_map, proposals

  def forward_batch(self, feature_map, image_shape, max_proposals_pre_nms, max_proposals_post_nms, anchor_map = None, anchor_valid_map = None):
    """EXTENSION (batch > 1): one conv + head pass over all B images (rows = B*H*W pixels), then the proposal
    path per image on its slice of the maps.  -> objectness (B,H,W,9), box deltas (B,H,W,36), [proposals_b (N_b,4)]."""
    bsz = int(feature_map.shape[0])
    y = ops.conv2d_act(feature_map, self._rpn_conv1.weight, self._rpn_conv1.bias, 1, 1, ops.ACT_RELU)
    scores, deltas = ops.two_heads(y, self._rpn_class.weight, self._rpn_class.bias, ops.ACT_SIGMOID, self._rpn_boxes.weight, self._rpn_boxes.bias, ops.ACT_NONE)
    fh, fw = int(y.shape[2]), int(y.shape[3])
    objectness_score_map = scores.view(bsz, fh, fw, scores.shape[1])
    box_deltas_map = deltas.view(bsz, fh, fw, deltas.shape[1])
    anchors_dev, keep_mask = self._resolve_anchors(anchor_map, anchor_valid_map, image_shape, (fh, fw), feature_map.device)
    proposals = [ops.rpn_proposals(objectness_score_map[b:b + 1], box_deltas_map[b:b + 1], image_shape, 16, max_proposals_pre_nms, max_proposals_post_nms,
                                   anchors = anchors_dev, keep_mask = keep_mask).clone() for b in range(bsz)]
    return objectness_score_map, box_deltas_map, proposals

  def _resolve_anchors(self, anchor_map, anchor_valid_map, image_shape, fm_hw, device):
    """The decode kernel regenerates the standard anchors itself (0 bytes of anchor traffic).  A
    caller-supplied anchor_map is honoured: if it differs from the standard map it is uploaded once
    and cached; the valid map is only needed when edge proposals are excluded (rpn.py:170-173)."""
    fh, fw = int(fm_hw[0]), int(fm_hw[1])
    key = (int(image_shape[1]), int(image_shape[2]), fh, fw, device.index)
    entry = self._anchor_cache.get(key)
    if entry is None:
      std_a, std_v = ops.generate_anchors_device(image_shape, (fh, fw), 16, device = device)
      entry = dict(std_anchors = std_a.cpu().numpy(), std_valid = std_v.cpu().numpy(), custom = {})
      self._anchor_cache[key] = entry
    anchors_dev = None
    if anchor_map is not None:
      am = anchor_map if isinstance(anchor_map, np.ndarray) else anchor_map.detach().cpu().numpy()
      if am is entry.get("verified_standard"):
        pass                                                              # same array object as last time: already compared equal
      elif am.shape == entry["std_anchors"].shape and np.array_equal(am, entry["std_anchors"]):
        entry["verified_standard"] = am                                   # holding the reference keeps the identity test sound
      else:
        ck = (am.shape, hashlib.blake2b(np.ascontiguousarray(am).tobytes(), digest_size = 16).digest())   # the WHOLE map: two maps sharing their first rows must not alias
        anchors_dev = entry["custom"].get(ck)
        if anchors_dev is None:
          anchors_dev = t.from_numpy(np.ascontiguousarray(am.reshape(-1, 4), dtype = np.float32)).to(device)
          if len(entry["custom"]) >= 8:
            entry["custom"].clear()                                       # bounded: a handful of alternating custom maps stay resident
          entry["custom"][ck] = anchors_dev
    keep_mask = None
    if not self._allow_edge_proposals:
      av = entry["std_valid"] if anchor_valid_map is None else (anchor_valid_map if isinstance(anchor_valid_map, np.ndarray) else anchor_valid_map.detach().cpu().numpy())
      mk = ("mask", hashlib.blake2b(np.ascontiguousarray(av).tobytes(), digest_size = 16).digest(), av.shape)
      keep_mask = entry.get(mk)
      if keep_mask is None:
        keep_mask = t.from_numpy((av.reshape(-1) > 0).astype(np.uint8)).to(device)
        entry[mk] = keep_mask
    return anchors_dev, keep_mask


def _fused(predicted_scores, predicted_box_deltas, y_true):
  return ops.rpn_losses(predicted_scores, predicted_box_deltas, y_true)


def class_loss(predicted_scores, y_true):
  """rpn.py:176-214.  (The model's train_step uses ops.rpn_losses, which yields both RPN losses in one launch.)"""
  a = predicted_scores.numel()
  dummy = t.zeros((a, 4), dtype = t.float32, device = predicted_scores.device)
  return ops.rpn_losses(predicted_scores, dummy, y_true)[0]


def regression_loss(predicted_box_deltas, y_true):
  """rpn.py:216-272."""
  a = predicted_box_deltas.numel() // 4
  dummy = t.full((a,), 0.5, dtype = t.float32, device = predicted_box_deltas.device)
  return ops.rpn_losses(dummy, predicted_box_deltas, y_true)[1]
