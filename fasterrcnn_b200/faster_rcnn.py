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

  def __init__(self, num_classes, backbone, rpn_minibatch_size = 256, proposal_batch_size = 128, allow_edge_proposals = True,
               roi_op = "pool", roi_sampling_ratio = 2, roi_aligned = False):
    """roi_op / roi_sampling_ratio / roi_aligned are EXTENSIONS (BASELINE config 3): "pool" is the reference's
    RoIPool; "align" swaps in RoIAlign with torchvision.ops.roi_align semantics."""
    super().__init__()
    self._num_classes = num_classes
    self._rpn_minibatch_size = rpn_minibatch_size
    self._proposal_batch_size = proposal_batch_size
    self._detector_box_delta_means = [0, 0, 0, 0]
    self._detector_box_delta_stds = [0.1, 0.1, 0.2, 0.2]
    self.backbone = backbone
    self._stage1_feature_extractor = backbone.feature_extractor
    self._stage2_region_proposal_network = rpn.RegionProposalNetwork(feature_map_channels = backbone.feature_map_channels, allow_edge_proposals = allow_edge_proposals)
    self._stage3_detector_network = detector.DetectorNetwork(num_classes = num_classes, backbone = backbone, roi_op = roi_op, roi_sampling_ratio = roi_sampling_ratio, roi_aligned = roi_aligned)
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
    if not self.training:
      self.train()
    optimizer.zero_grad()
    assert image_data.shape[0] == 1, "Batch size must be 1"
    assert len(gt_rpn_map.shape) == 5 and gt_rpn_map.shape[0] == 1, "Batch size must be 1"
    assert len(gt_rpn_object_indices) == 1, "Batch size must be 1"
    assert len(gt_rpn_background_indices) == 1, "Batch size must be 1"
    assert len(gt_boxes) == 1, "Batch size must be 1"
    image_shape = image_data.shape[1:]
    ops.begin_step()
    dev = image_data.device
    gt = gt_boxes[0]

    # Everything up to the proposal labels is enqueued without a host round trip: the proposal count stays on the device, the
    # GT boxes are appended behind it there (faster_rcnn.py:467), labels are computed for every row of the padded buffer.
    backbone_map = self._stage1_feature_extractor(image_data = image_data)
    # The two consumers of the shared feature map are back-propagated separately (the RPN branch early, see below), each leaving
    # its gradient on this detached leaf; the backbone then gets their sum -- the same two-term sum autograd would form.
    feature_map = backbone_map.detach().requires_grad_(True) if backbone_map.requires_grad else backbone_map
    # The proposal path (decode, top-N, filter, NMS, GT append, labelling, the read-back of count + labels) is a serial chain of small,
    # latency-bound launches that needs only the RPN head outputs; the RPN losses and the whole backward of the RPN branch need only those
    # outputs too.  The chain therefore runs on a side stream and the RPN branch on the compute stream AT THE SAME TIME (the GEMM CTAs
    # leave room for other kernels' CTAs on every SM, conv_tc.cu), instead of one after the other; the host waits for the chain's
    # read-back -- the step's one mid-step synchronisation -- draws the samples, and the detector starts when both streams are done.
    # FRCNN_PROPOSAL_STREAM=0 keeps everything on one stream.
    main = t.cuda.current_stream()
    side = self._proposal_stream(dev)
    rpn_score_map, rpn_box_deltas_map, (padded, count) = self._stage2_region_proposal_network(
      feature_map = feature_map, image_shape = image_shape, anchor_map = anchor_map, anchor_valid_map = anchor_valid_map,
      max_proposals_pre_nms = 12000, max_proposals_post_nms = 2000, deferred_extra_rows = len(gt), proposal_stream = side)

    # host work that needs nothing from the device overlaps the backbone / RPN kernels queued above
    gt_rpn_minibatch_map = self._sample_rpn_minibatch(rpn_map = gt_rpn_map, object_indices = gt_rpn_object_indices, background_indices = gt_rpn_background_indices)
    fetch = self._pinned("fetch", (1 + padded.shape[0],), t.int32)
    fetched = t.cuda.Event()
    with t.cuda.stream(side if side is not None else main):
      # small host arrays go up through page-locked staging buffers: a cudaMemcpyAsync from PAGEABLE memory first synchronises the
      # stream, which would stall the host here until the previous step's backward has drained
      gt_box_corners = self._upload("gt_corners", np.array([box.corners for box in gt], dtype = np.float32).reshape(-1, 4), dev)
      gt_box_class_idxs = self._upload("gt_classes", np.array([box.class_index for box in gt], dtype = np.int32), dev)
      ops.append_rows(padded, count, gt_box_corners)
      _, class_idx, gt_classes, gt_box_deltas = ops.label_proposals(padded, gt_box_corners, gt_box_class_idxs, self._num_classes, 0.5)
      # the step's one mid-step synchronisation: proposal count + class labels in a single pinned read-back
      fetch[0:1].copy_(count, non_blocking = True)
      fetch[1:].copy_(class_idx, non_blocking = True)
      fetched.record()
    if side is not None:
      for x in (padded, gt_classes, gt_box_deltas, class_idx, gt_box_corners, gt_box_class_idxs):
        x.record_stream(main)                                       # allocated on the side stream, consumed on the compute stream
    rpn_l = ops.rpn_losses(rpn_score_map, rpn_box_deltas_map, gt_rpn_minibatch_map)                # (class, regression)
    ones = self._ones2(dev)
    t.autograd.backward([rpn_l], [ones])                                                           # d(total)/d(loss) = 1 for every term
    fetched.synchronize()
    if side is not None:
      main.wait_event(fetched)                                      # the detector reads what the side stream produced
    n = int(fetch[0]) + len(gt)                                                                    # proposals + appended GT boxes
    indices = self._sample_proposal_indices(fetch[1:1 + n].numpy(), self._proposal_batch_size, 0.25)
    if indices is None or len(indices) == 0:
      keep = n if indices is None else 0                                                           # no sampling | no usable sample
      proposals, gt_classes, gt_box_deltas = padded[:keep], gt_classes[:keep], gt_box_deltas[:keep]
      detector_classes, detector_box_deltas = self._stage3_detector_network(feature_map = feature_map, proposals = proposals)
    else:
      # sampled row indices go up in one pinned copy ([count | indices]); the proposal rows are gathered first so that RoI pooling
      # -- the first sizeable kernel after the synchronisation -- is in flight before the label rows are gathered
      k = len(indices)
      upload = self._pinned("sample", (1 + self._proposal_batch_size if self._proposal_batch_size > 0 else 1 + padded.shape[0],), t.int32)
      upload[0] = k
      upload[1:1 + k] = t.from_numpy(indices.astype(np.int32))
      sample_dev = upload[:1 + k].to(dev, non_blocking = True)
      proposals = ops.gather_rows(padded, sample_dev[1:], sample_dev[0:1], k)
      detector_classes, detector_box_deltas = self._stage3_detector_network(feature_map = feature_map, proposals = proposals)
      gt_classes = ops.gather_rows(gt_classes, sample_dev[1:], sample_dev[0:1], k)
      gt_box_deltas = ops.gather_rows(gt_box_deltas, sample_dev[1:], sample_dev[0:1], k)

    det_l = ops.detector_losses(detector_classes, detector_box_deltas, gt_classes, gt_box_deltas)    # (class, regression)
    all_l = t.cat([rpn_l.detach(), det_l.detach()])
    # The five numbers are final once the forward kernels have run: their read-back is queued NOW, in front of the backward and
    # optimizer kernels, and only its event is awaited at the end -- train_step returns with the backward still in flight, so the
    # next step's launches queue up behind it and the device never drains between steps.
    loss_host = self._pinned("loss", (5,), t.float32)
    loss_host.copy_(t.cat([all_l, all_l.sum().reshape(1)]), non_blocking = True)
    loss_ready = t.cuda.Event()
    loss_ready.record()
    if det_l.requires_grad:
      t.autograd.backward([det_l], [ones])                                                         # detector branch -> its params, feature_map.grad
    if feature_map is not backbone_map and feature_map.grad is not None:
      backbone_map.backward(feature_map.grad)                                                      # both branches' gradient into the backbone
    ops.begin_step()                       # operand splits of this step's weights are stale after the update
    optimizer.step()

    loss_ready.synchronize()
    host = loss_host.numpy()
    self.last_step_info = dict(num_rois = int(proposals.shape[0]), sampled_proposals = proposals)      # (the RoIs the detector was trained on: device tensor, no copy)
    return FasterRCNNModel.Loss(rpn_class = float(host[0]), rpn_regression = float(host[1]), detector_class = float(host[2]), detector_regression = float(host[3]), total = float(host[4]))

  # ---- EXTENSION: batch > 1 (BASELINE config 3).  The reference asserts batch == 1 everywhere; its loss formulas
  # are written for batched tensors, so the batch semantics are those formulas on the stacked maps / stacked RoIs, with proposal
  # generation, labelling and sampling per image. ------------------------------------------------------------------------------
  def forward_batch(self, image_data, max_proposals_pre_nms = 6000, max_proposals_post_nms = 300):
    """image_data (B,3,H,W) -> [(proposals_b (N_b,4), classes_b (N_b,C), box_deltas_b (N_b,4(C-1)))]."""
    image_shape = image_data.shape[1:]
    ops.begin_step()
    feature_map = self._stage1_feature_extractor(image_data = image_data)
    _, _, proposals = self._stage2_region_proposal_network.forward_batch(feature_map, image_shape, max_proposals_pre_nms, max_proposals_post_nms)
    classes, box_deltas = self._stage3_detector_network.forward_batch(feature_map, proposals)
    out, at = [], 0
    for p in proposals:
      out.append((p, classes[at:at + p.shape[0]], box_deltas[at:at + p.shape[0]]))
      at += p.shape[0]
    return out

  def predict_batch(self, image_data, score_threshold):
    """-> [Dict[int, ndarray (n,5)]] per image (faster_rcnn.py:134-226 applied to each image's slice)."""
    self.eval()
    with t.no_grad():
      return [ops.detect_postprocess(p, c, d, (image_data.shape[2], image_data.shape[3]), score_threshold, 0.3) for p, c, d in self.forward_batch(image_data)]

  def train_step_batch(self, optimizer, image_data, samples):
    """One optimizer step on B images of one size.  samples[b] holds train_step's per-image arguments: anchor_map,
    anchor_valid_map, gt_rpn_map (1,h,w,9,6) CUDA, gt_rpn_object_indices, gt_rpn_background_indices (ndarrays), gt_boxes ([Box]).
    RNG order: every image's RPN minibatch first (python random), then every image's proposal sample (torch CPU generator)."""
    if not self.training:
      self.train()
    optimizer.zero_grad()
    bsz = int(image_data.shape[0])
    assert len(samples) == bsz, "one sample dict per image"
    image_shape = image_data.shape[1:]
    dev = image_data.device
    ops.begin_step()
    feature_map = self._stage1_feature_extractor(image_data = image_data)
    score_map, delta_map, all_proposals = self._stage2_region_proposal_network.forward_batch(
      feature_map, image_shape, 12000, 2000, anchor_map = samples[0]["anchor_map"], anchor_valid_map = samples[0]["anchor_valid_map"])
    minibatch = t.cat([self._sample_rpn_minibatch(rpn_map = s["gt_rpn_map"], object_indices = [s["gt_rpn_object_indices"]], background_indices = [s["gt_rpn_background_indices"]], slot = b)
                       for b, s in enumerate(samples)], dim = 0)
    props_l, cls_l, dlt_l = [], [], []
    for b, s in enumerate(samples):
      props, gt_classes, gt_box_deltas = self._label_proposals(all_proposals[b], s["gt_boxes"], 0.0, 0.5)
      props, gt_classes, gt_box_deltas = self._sample_proposals(props, gt_classes, gt_box_deltas, self._proposal_batch_size, 0.25)
      props_l.append(props.detach()); cls_l.append(gt_classes.detach()); dlt_l.append(gt_box_deltas.detach())
    classes, box_deltas = self._stage3_detector_network.forward_batch(feature_map, props_l)
    rpn_l = ops.rpn_losses(score_map, delta_map, minibatch)
    det_l = ops.detector_losses(classes, box_deltas, t.cat(cls_l), t.cat(dlt_l))
    ones = self._ones2(dev)
    t.autograd.backward([rpn_l, det_l], [ones, ones])
    ops.begin_step()
    optimizer.step()
    host = t.cat([rpn_l.detach(), det_l.detach()]).cpu().numpy()
    self.last_step_info = dict(num_rois = int(sum(p.shape[0] for p in props_l)), rois_per_image = [int(p.shape[0]) for p in props_l])
    return FasterRCNNModel.Loss(rpn_class = float(host[0]), rpn_regression = float(host[1]), detector_class = float(host[2]), detector_regression = float(host[3]), total = float(host.sum()))

  def _proposal_stream(self, device):
    import os
    if os.environ.get("FRCNN_PROPOSAL_STREAM", "1") in ("", "0"):
      return None
    cache = self.__dict__.setdefault("_proposal_streams", {})
    key = str(device)
    if key not in cache:
      cache[key] = t.cuda.Stream(device = device)
    return cache[key]

  def _ones2(self, device):
    cache = self.__dict__.setdefault("_ones2_cache", {})
    key = str(device)
    if key not in cache:
      cache[key] = t.ones((2,), dtype = t.float32, device = device)
    return cache[key]

  def _upload(self, name, array, device):
    """Asynchronous host-to-device copy of a small numpy array through a reusable page-locked buffer.  Two buffers per name,
    alternated per call: the previous call's copy is ordered before this step's mid-step synchronisation, so a buffer is never
    rewritten while its copy is still pending."""
    flip = self.__dict__.setdefault("_upload_flip", {})
    flip[name] = 1 - flip.get(name, 0)
    src = t.from_numpy(np.ascontiguousarray(array))
    capacity = 64
    while capacity < src.numel():
      capacity *= 2                                                   # size classes keep the set of pinned allocations small
    staging = self._pinned((name, flip[name]), (capacity,), src.dtype)
    view = staging[:src.numel()].view(src.shape)
    view.copy_(src)
    return view.to(device, non_blocking = True)

  def _pinned(self, name, shape, dtype):
    """Reusable page-locked host buffer for the small device-to-host reads of train_step."""
    buffers = self.__dict__.setdefault("_pinned_buffers", {})
    key = (name, tuple(shape), dtype)
    buf = buffers.get(key)
    if buf is None:
      buf = t.empty(shape, dtype = dtype).pin_memory()
      buffers[key] = buf
    return buf

  # ---- faster_rcnn.py:364-416 --------------------------------------------------------------------
  def _sample_rpn_minibatch(self, rpn_map, object_indices, background_indices, slot = 0):
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
    flat_dev = self._upload(("rpn_minibatch", slot), flat.astype(np.int64), rpn_map.device)    # slot: one staging pair per image of a batch
    minibatch = rpn_map.clone()
    mask = minibatch.view(-1, 6)[:, 0]
    mask.zero_()
    mask.index_fill_(0, flat_dev, 1.0)          # (mask[idx] = 1.0 would upload the scalar from pageable memory: a hidden stream sync)
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
  def _sample_proposal_indices(self, cls_host, max_proposals, positive_fraction):
    """Host half of faster_rcnn.py:526-561 on the class labels (numpy int): row indices to keep (int64), an empty array when the
    image yields no positive or no negative sample, None when sampling is off (max_proposals <= 0: keep every row)."""
    if max_proposals <= 0:
      return None
    positive_indices = np.where(cls_host > 0)[0]
    negative_indices = np.where(cls_host <= 0)[0]
    num_samples = min(max_proposals, len(cls_host))
    num_positive_samples = min(round(num_samples * positive_fraction), len(positive_indices))
    num_negative_samples = min(num_samples - num_positive_samples, len(negative_indices))
    if num_positive_samples <= 0 or num_negative_samples <= 0:
      return np.zeros((0,), dtype = np.int64)
    positive_sample_indices = positive_indices[t.randperm(len(positive_indices))[0:num_positive_samples].numpy()]   # CPU generator, as the reference
    negative_sample_indices = negative_indices[t.randperm(len(negative_indices))[0:num_negative_samples].numpy()]
    return np.concatenate([positive_sample_indices, negative_sample_indices]).astype(np.int64)

  def _sample_proposals(self, proposals, gt_classes, gt_box_deltas, max_proposals, positive_fraction):
    class_indices = getattr(self, "_last_class_idx", None)
    if class_indices is None or class_indices.shape[0] != gt_classes.shape[0]:
      class_indices = t.argmax(gt_classes, dim = 1)
    indices = self._sample_proposal_indices(class_indices.cpu().numpy(), max_proposals, positive_fraction)
    if indices is None:
      return proposals, gt_classes, gt_box_deltas
    indices = t.from_numpy(indices).to(proposals.device)
    return proposals[indices], gt_classes[indices], gt_box_deltas[indices]
