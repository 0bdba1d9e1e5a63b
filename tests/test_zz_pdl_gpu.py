"""
Programmatic dependent launch (include/frcnn_b200.h: frcnn_set_pdl, FRCNN_PDL=1) is launch plumbing: every kernel of the library waits
for its predecessor on the stream (griddepcontrol.wait) before its first global access, so switching it on may only change WHEN kernels
become resident, never a result.  The same seeded train steps and the same prediction must come out bit for bit with it off and on.
(Named zz so that it runs after the parity suites.)
"""
import random

import numpy as np
import pytest
import torch as t

from oracle import frcnn_oracle as orc

pytestmark = pytest.mark.gpu


class Box:
  def __init__(self, corners, class_index):
    self.corners, self.class_index, self.class_name = corners, class_index, str(class_index)


def _run(pdl, params, smp):
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import _lib, optim
  before = _lib.set_pdl(pdl)
  try:
    model = f.FasterRCNNModel(num_classes = 21, backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0), allow_edge_proposals = True)
    model.load_state_dict(params)
    model = model.cuda()
    optimizer = optim.create_optimizer(model, 1e-3, 0.9, 5e-4, fused = True)
    boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
    random.seed(0); np.random.seed(0); t.manual_seed(0)
    losses = []
    for _ in range(3):
      l = model.train_step(optimizer = optimizer, image_data = smp["image"].cuda(), anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                           gt_rpn_map = smp["gt_rpn_map"].cuda(), gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                           gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
      losses.append((l.rpn_class, l.rpn_regression, l.detector_class, l.detector_regression, l.total))
    weights = {k: p.detach().cpu().numpy().copy() for k, p in model.named_parameters()}
    detections = model.predict(image_data = smp["image"].cuda(), score_threshold = 0.0)
    t.cuda.synchronize()
    return losses, weights, detections
  finally:
    _lib.set_pdl(before)


def test_pdl_changes_no_result():
  params = orc.synth_params(orc.vgg16_param_shapes(), seed = 0, heads = "spread")
  smp = orc.synthetic_sample((384, 512), seed = 0)
  off = _run(False, params, smp)
  on = _run(True, params, smp)
  same = lambda a, b: a.shape == b.shape and a.tobytes() == b.tobytes()          # bit patterns (a NaN would still have to be the same NaN)
  assert same(np.array(off[0]), np.array(on[0])), (off[0], on[0])                # five losses of three steps
  for k in off[1]:
    assert same(off[1][k], on[1][k]), k                                          # weights after three SGD steps
  assert sorted(off[2]) == sorted(on[2])
  for c in off[2]:
    assert same(np.asarray(off[2][c]), np.asarray(on[2][c])), c                  # per-class boxes + scores
