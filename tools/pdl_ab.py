#!/usr/bin/env python
"""A/B of programmatic dependent launch (frcnn_set_pdl) on the bench workload, one process, one GPU:
the same seeded 600x1000 VGG-16 train steps with PDL off / on / off again -- the loss sequences must be bit-identical (every kernel waits
for its predecessor before its first global access, so results cannot change), and the CUDA-event time per step is reported for each leg.
  python tools/pdl_ab.py [steps] > gpurun_out/pdl_ab.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t

import bench
from fasterrcnn_b200 import _lib

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
t0 = time.time()
dev = t.device("cuda", 0)
t.cuda.set_device(0)
legs = []
for name, on in (("off", False), ("on", True), ("off_again", False)):
  _lib.set_pdl(on)
  step = bench.make_train_step(dev, 0)
  losses = [step(False).total for _ in range(5)]
  t.cuda.synchronize()
  e0, e1 = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
  e0.record()
  for _ in range(steps):
    losses.append(step(False).total)
  e1.record()
  t.cuda.synchronize()
  legs.append(dict(leg = name, pdl = on, ms_per_step = e0.elapsed_time(e1) / steps, losses = losses))
  # inference path too (decode -> top-N -> NMS -> RoIPool -> heads -> per-class post-processing): boxes must not change either
  g = t.Generator(device = "cpu").manual_seed(7)
  img = (t.randn((1, 3, 600, 800), generator = g) * 50.0).cuda()
  det = step.model.predict(image_data = img, score_threshold = 0.0)
  legs[-1]["detections"] = {int(k): v.tobytes().hex() for k, v in det.items()}
  legs[-1]["n_detections"] = int(sum(v.shape[0] for v in det.values()))
  del step
  print("%s: %.3f ms/step (t+%.1f s)" % (name, legs[-1]["ms_per_step"], time.time() - t0), file = sys.stderr, flush = True)
_lib.set_pdl(False)
same_on = legs[0]["losses"] == legs[1]["losses"]
same_off = legs[0]["losses"] == legs[2]["losses"]
same_det = legs[0]["detections"] == legs[1]["detections"]
print(json.dumps(dict(steps = steps, bit_identical_on_vs_off = same_on, bit_identical_off_vs_off = same_off, predict_identical_on_vs_off = same_det,
                      n_detections = {l["leg"]: l["n_detections"] for l in legs},
                      ms_per_step = {l["leg"]: l["ms_per_step"] for l in legs},
                      images_per_s = {l["leg"]: 1e3 / l["ms_per_step"] for l in legs},
                      first_losses = {l["leg"]: l["losses"][:3] for l in legs}, last_losses = {l["leg"]: l["losses"][-2:] for l in legs})))
