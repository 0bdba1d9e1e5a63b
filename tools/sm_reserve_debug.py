"""Per-tensor gradient difference of ONE train step with the persistent GEMM grids on 148 vs 148 - n SMs (frcnn_set_sm_reserve): separates
"another summation order" (1e-6) from a decomposition bug (1e-3).  python tools/sm_reserve_debug.py [n]"""
import os
import random
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasterrcnn_b200 as f  # noqa: E402
from fasterrcnn_b200 import _lib  # noqa: E402
from oracle import frcnn_oracle as orc  # noqa: E402


class Box:
  def __init__(self, corners, class_index):
    self.corners, self.class_index = corners, class_index


def one_step(params, smp, reserve, hw):
  model = f.FasterRCNNModel(num_classes = 21, backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0))
  model.load_state_dict(params)
  model = model.cuda()
  opt = t.optim.SGD([{"params": [p], "weight_decay": 5e-4} for k, p in model.named_parameters() if p.requires_grad and "weight" in k], lr = 0.0, momentum = 0.9)
  boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  random.seed(0); np.random.seed(0); t.manual_seed(0)
  before = _lib.set_sm_reserve(reserve)
  try:
    loss = model.train_step(optimizer = opt, image_data = smp["image"].cuda(), anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                            gt_rpn_map = smp["gt_rpn_map"].cuda(), gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                            gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
    t.cuda.synchronize()
  finally:
    _lib.set_sm_reserve(before)
  return loss, {k: p.grad.detach().double().cpu() for k, p in model.named_parameters() if p.grad is not None}


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
  for hw in ((384, 512), (600, 1000)):
    params = orc.synth_params(orc.vgg16_param_shapes(), seed = 0, heads = "spread")
    smp = orc.synthetic_sample(hw, seed = 0)
    l0, g0 = one_step(params, smp, 0, hw)
    l1, g1 = one_step(params, smp, n, hw)
    l2, g2 = one_step(params, smp, 0, hw)
    print("image %dx%d  reserve 0 vs %d: total loss %.9g vs %.9g (repeat of 0: %.9g)" % (hw[0], hw[1], n, l0.total, l1.total, l2.total))
    for k in g0:
      rel = float((g0[k] - g1[k]).norm() / (g0[k].norm() + 1e-30))
      rep = float((g0[k] - g2[k]).norm() / (g0[k].norm() + 1e-30))
      print("  %-72s rel-L2 vs reserve %d: %.3g   (run-to-run at reserve 0: %.3g)" % (k, n, rel, rep))


if __name__ == "__main__":
  main()
