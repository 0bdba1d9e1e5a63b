"""Compares the device RPN ground-truth map (frcnn_rpn_targets via anchors.generate_rpn_map) with the CPU restatement, channel by channel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import frcnn_oracle as orc, golden_inputs as gi
from fasterrcnn_b200 import anchors


class B:
  def __init__(self, c):
    self.corners = np.asarray(c, dtype = np.float32)


for tag, (h, w) in list(gi.GEOMETRY_CASES.items()) + [("voc600x800", (600, 800)), ("voc_odd", (600, 901))]:
  fm = (512, h // 16, w // 16)
  am, av = orc.generate_anchor_maps((3, h, w), fm, 16)
  gt = np.array([b for b, _ in gi.gt_boxes_for(h, w)], dtype = np.float32) * np.float32(1.2345)
  gt = np.minimum(gt, np.float32(min(h, w) - 1))
  rm, obj, bg = orc.generate_rpn_map(am, av, gt)
  am2, av2 = anchors.generate_anchor_maps((3, h, w), fm, 16)
  rm2, obj2, bg2 = anchors.generate_rpn_map(am2, av2, [B(c) for c in gt])
  print(tag, "anchors eq", np.array_equal(am, am2), np.array_equal(av, av2), "dtype", rm.dtype, rm2.dtype, obj.dtype, obj2.dtype, obj.shape, obj2.shape, bg.shape, bg2.shape)
  for ch in range(6):
    d = np.argwhere(rm[..., ch] != rm2[..., ch])
    if len(d):
      i = tuple(d[0])
      print("   channel", ch, "mismatches", len(d), "first", i, rm[i + (slice(None),)], rm2[i + (slice(None),)], "valid", av[i])
  print("   obj eq", np.array_equal(obj, obj2), "bg eq", np.array_equal(bg, bg2))
