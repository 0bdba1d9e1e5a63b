"""
TEST INFRASTRUCTURE ONLY -- differential check of the restatement (oracle/frcnn_oracle.py) against the UNMODIFIED reference executed
live (oracle/ref_shim.py), on seeds / image sizes / ground-truth layouts that are NOT among the committed golden vectors.

  python oracle/check_vs_live_reference.py [--cases N]

Runs only where /root/reference exists (the build container); tests/test_oracle.py launches it in a subprocess, because the shim
replaces Tensor.cuda / Module.cuda process-wide.  For every case: forward (proposals, class scores, box deltas), predict (per-class boxes)
and two train steps (five losses each, every gradient after step 1, every weight after step 2) of reference vs restatement, VGG-16.
Both sides run the same torch CPU kernels, so the bars are tight: indices / counts exact, floats 1e-5 relative.
"""
import argparse
import os
import random
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import frcnn_oracle as orc
from oracle import ref_shim

# (image h, w), weight seed, sample seed, head init, ground truth (corners y1,x1,y2,x2 ; class)
CASES = [
  ((320, 480), 11, 5, "spread", [((40.0, 60.0, 200.0, 300.0), 3), ((100.0, 250.0, 300.0, 470.0), 12)]),
  ((352, 400), 12, 6, "spread", [((10.0, 20.0, 340.0, 380.0), 1)]),                                                   # one large object
  ((336, 448), 13, 7, "reference", [((50.0, 40.0, 150.0, 140.0), 20), ((60.0, 60.0, 170.0, 180.0), 20), ((200.0, 300.0, 330.0, 440.0), 9)]),   # overlapping, N(0,0.01) heads (scores ~0.5: ties)
]


# ResNet cases (--resnet): the same comparison through the reference's torchvision-bottleneck backbones (frozen BN; layer4 as the head)
RESNET_CASES = [
  ("resnet101", (320, 400), 21, 8, "spread", [((30.0, 40.0, 250.0, 300.0), 5), ((120.0, 200.0, 310.0, 390.0), 17)]),
  ("resnet50", (304, 368), 22, 9, "spread", [((20.0, 30.0, 280.0, 200.0), 2)]),
]


def run_case(ref, hw, wseed, sseed, heads, gt, backbone = "vgg16"):
  import math
  if backbone == "vgg16":
    params = orc.synth_params(orc.vgg16_param_shapes(), seed = wseed, heads = heads)
    ref_backbone = ref.vgg16.VGG16Backbone(dropout_probability = 0.0)
    fm_shape = (512, hw[0] // 16, hw[1] // 16)
  else:
    from oracle import resnet_oracle
    ref_shim.patch_resnet_offline()
    params = orc.synth_params(resnet_oracle.param_shapes(backbone), seed = wseed, heads = heads)
    ref_backbone = ref.resnet.ResNetBackbone(architecture = {"resnet50": ref.resnet.Architecture.ResNet50, "resnet101": ref.resnet.Architecture.ResNet101}[backbone])
    fm_shape = (1024, math.ceil(hw[0] / 16), math.ceil(hw[1] / 16))
  smp = orc.synthetic_sample(hw, seed = sseed, gt = gt, backbone = backbone)
  image = smp["image"]
  model = ref.faster_rcnn.FasterRCNNModel(num_classes = 21, backbone = ref_backbone, allow_edge_proposals = True)
  model.load_state_dict(params)
  oracle = orc.OracleModel(params, backbone = backbone)

  # the restated anchor / RPN-map generators against the reference's, bit for bit
  am, av = ref.anchors.generate_anchor_maps(image_shape = (3,) + hw, feature_map_shape = fm_shape, feature_pixels = 16)
  boxes = [ref.Box(class_index = c, class_name = str(c), corners = b) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  rm, obj, bg = ref.anchors.generate_rpn_map(anchor_map = am, anchor_valid_map = av, gt_boxes = boxes)
  assert np.array_equal(am, smp["anchor_map"]) and np.array_equal(av, smp["anchor_valid_map"])
  assert np.array_equal(rm, smp["gt_rpn_map"][0].numpy()) and np.array_equal(obj, smp["gt_rpn_object_indices"]) and np.array_equal(bg, smp["gt_rpn_background_indices"])

  model.eval()
  with t.no_grad():
    p_ref, c_ref, d_ref = model(image_data = image)
    p, c, d = oracle.forward(image)
  assert p.shape == p_ref.shape, (p.shape, p_ref.shape)
  np.testing.assert_allclose(p.numpy(), p_ref.numpy(), rtol = 0, atol = 1e-4)
  np.testing.assert_allclose(c.numpy(), c_ref.numpy(), rtol = 0, atol = 1e-5)
  np.testing.assert_allclose(d.numpy(), d_ref.numpy(), rtol = 0, atol = 1e-5)
  pr_ref = model.predict(image_data = image, score_threshold = 0.05)
  pr = oracle.predict(image, 0.05)
  for k in range(1, 21):
    assert pr[k].shape == pr_ref[k].shape, (k, pr[k].shape, pr_ref[k].shape)
    np.testing.assert_allclose(pr[k], pr_ref[k], rtol = 0, atol = 1e-4)

  opt_params = [{"params": [v], "weight_decay": 5e-4} for k, v in dict(model.named_parameters()).items() if v.requires_grad and "weight" in k]
  optimizer = t.optim.SGD(opt_params, lr = 1e-3, momentum = 0.9)                 # __main__.py:98-105
  out = {}
  for who in ("ref", "oracle"):
    random.seed(sseed); np.random.seed(sseed); t.manual_seed(sseed)
    losses, grads = [], None
    for step in range(2):
      if who == "ref":
        l = model.train_step(optimizer = optimizer, image_data = image, anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                             gt_rpn_map = smp["gt_rpn_map"], gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                             gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
        if step == 0:
          grads = {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}
      else:
        l = oracle.train_step(image, smp["anchor_map"], smp["anchor_valid_map"], smp["gt_rpn_map"], smp["gt_rpn_object_indices"],
                              smp["gt_rpn_background_indices"], smp["gt_corners"], smp["gt_class_idxs"])
        if step == 0:
          grads = {k: v.grad.clone() for k, v in oracle.params.items() if v.grad is not None}
      losses.append([l.rpn_class, l.rpn_regression, l.detector_class, l.detector_regression, l.total])
    out[who] = (np.array(losses), grads)
  np.testing.assert_allclose(out["oracle"][0], out["ref"][0], rtol = 1e-5, atol = 1e-6)
  assert set(out["oracle"][1]) == set(out["ref"][1])
  for k, g_ref in out["ref"][1].items():
    g = out["oracle"][1][k]
    scale = float(g_ref.abs().max()) + 1e-12
    np.testing.assert_allclose(g.numpy(), g_ref.numpy(), rtol = 1e-4, atol = 1e-5 * scale, err_msg = k)
  sd = model.state_dict()
  for k, v in oracle.params.items():
    np.testing.assert_allclose(v.detach().numpy(), sd[k].numpy(), rtol = 1e-5, atol = 1e-6, err_msg = k)    # lr * (gradient differences of 1e-4 relative)
  return dict(hw = hw, proposals = int(p.shape[0]), detections = int(sum(pr[k].shape[0] for k in range(1, 21))), losses = out["ref"][0][:, 4].tolist())


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--cases", type = int, default = len(CASES))
  ap.add_argument("--resnet", action = "store_true", help = "the ResNet-101 / ResNet-50 cases instead of the VGG-16 ones")
  args = ap.parse_args()
  if not ref_shim.available():
    print("SKIP: reference tree not present")
    return 0
  t.set_num_threads(min(8, os.cpu_count() or 8))
  ref = ref_shim.load()
  cases = [c[1:] + (c[0],) for c in RESNET_CASES] if args.resnet else CASES
  for case in cases[:args.cases]:
    print("OK", case[-1] if args.resnet else "vgg16", run_case(ref, *case), flush = True)
  print("PASS: restatement == live reference on %d fresh cases" % min(args.cases, len(cases)))
  return 0


if __name__ == "__main__":
  sys.exit(main())
