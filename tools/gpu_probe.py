"""Development probe (not part of the product): times the VGG-16 600x1000 train step and the
individual conv layers on the current engine.  Usage: python tools/gpu_probe.py [engine]"""
import os
import random
import sys
import time

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasterrcnn_b200 as f
from fasterrcnn_b200 import ops
from oracle import frcnn_oracle as orc


def ev_time(fn, iters = 5, warm = 2):
  for _ in range(warm):
    fn()
  t.cuda.synchronize()
  a, b = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  t.cuda.synchronize()
  return a.elapsed_time(b) / iters


def main():
  engine = sys.argv[1] if len(sys.argv) > 1 else "auto"
  ops.set_engine(engine)
  print("engine", engine, "device", t.cuda.get_device_name(0))
  layers = [(3, 64, 600, 1000), (64, 64, 600, 1000), (64, 128, 300, 500), (128, 128, 300, 500), (128, 256, 150, 250), (256, 256, 150, 250),
            (256, 512, 75, 125), (512, 512, 75, 125), (512, 512, 37, 62)]
  for cin, cout, h, w in layers:
    x = ops.as_nhwc(t.randn((1, cin, h, w), device = "cuda"))
    wt = t.randn((cout, cin, 3, 3), device = "cuda").contiguous(memory_format = t.channels_last)
    b = t.zeros((cout,), device = "cuda")
    dy = ops.as_nhwc(t.randn((1, cout, h, w), device = "cuda"))
    gf = 2 * 9 * cin * cout * h * w / 1e9
    tf = ev_time(lambda: ops.conv2d_fwd_raw(x, wt, b, 1, 1, ops.ACT_RELU))
    line = "conv %4d->%4d @%4dx%4d  %7.2f GFLOP  fwd %7.3f ms (%6.1f TF/s)" % (cin, cout, h, w, gf, tf, gf / tf)
    if cin % 4 == 0:
      td = ev_time(lambda: ops.conv2d_dgrad_raw(dy, wt, (1, cin, h, w), 1, 1))
      tw = ev_time(lambda: ops.conv2d_wgrad_raw(dy, x, (cout, cin, 3, 3), 1, 1))
      line += "  dgrad %7.3f ms (%6.1f)  wgrad %7.3f ms (%6.1f)" % (td, gf / td, tw, gf / tw)
    print(line, flush = True)
  for m, k, n in [(128, 25088, 4096), (128, 4096, 4096)]:
    x = t.randn((m, k), device = "cuda", requires_grad = True)
    wt = t.randn((n, k), device = "cuda", requires_grad = True)
    b = t.zeros((n,), device = "cuda")
    gf = 2 * m * k * n / 1e9
    tf = ev_time(lambda: ops.linear_act(x, wt, b, ops.ACT_RELU))
    y = ops.linear_act(x, wt, b, ops.ACT_RELU)
    g = t.randn_like(y)
    tb = ev_time(lambda: y.backward(g, retain_graph = True))
    print("linear %d x %d x %d  %.2f GFLOP  fwd %.3f ms (%.1f TF/s)  bwd(dx+dw) %.3f ms (%.1f TF/s)" % (m, k, n, gf, tf, gf / tf, tb, 2 * gf / tb), flush = True)

  params = orc.synth_params(orc.vgg16_param_shapes(), seed = 0, heads = "reference")
  model = f.FasterRCNNModel(num_classes = 21, backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0))
  model.load_state_dict(params)
  model = model.cuda()
  smp = orc.synthetic_sample((600, 1000), seed = 0)

  class Box:
    def __init__(self, c, k):
      self.corners, self.class_index = c, k
  boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  opt = t.optim.SGD([{"params": [p], "weight_decay": 5e-4} for k, p in model.named_parameters() if p.requires_grad and "weight" in k], lr = 1e-3, momentum = 0.9)
  img = smp["image"].cuda()
  gmap = smp["gt_rpn_map"].cuda()
  random.seed(0); t.manual_seed(0)

  def step():
    return model.train_step(optimizer = opt, image_data = img, anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"], gt_rpn_map = gmap,
                            gt_rpn_object_indices = [smp["gt_rpn_object_indices"]], gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
  for i in range(3):
    t.cuda.synchronize(); t0 = time.time()
    l = step()
    t.cuda.synchronize()
    print("train_step %d: %.1f ms  loss %s rois %s" % (i, (time.time() - t0) * 1e3, l, model.last_step_info), flush = True)
  ms = ev_time(step, iters = 5, warm = 1)
  print("train_step steady: %.2f ms -> %.2f img/s" % (ms, 1000.0 / ms))
  model.eval()
  with t.no_grad():
    ms = ev_time(lambda: model.predict(image_data = img, score_threshold = 0.05), iters = 5, warm = 1)
  print("predict steady: %.2f ms" % ms)
  from torch.profiler import profile, ProfilerActivity
  with profile(activities = [ProfilerActivity.CUDA]) as prof:
    step()
    t.cuda.synchronize()
  print(prof.key_averages().table(sort_by = "cuda_time_total", row_limit = 25))


if __name__ == "__main__":
  main()
