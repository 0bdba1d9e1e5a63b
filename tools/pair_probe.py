"""A/B probe of the single-CTA vs CTA-pair GEMM kernels on the VGG-16 layer shapes (forward, dgrad, wgrad): CUDA-event time per launch.
FRCNN_TC_PAIR is read once per process:  FRCNN_TC_PAIR=0 python tools/pair_probe.py ; FRCNN_TC_PAIR=1 python tools/pair_probe.py
(under ncu: `ncu --set full -k regex:tc_conv_kernel -c N ...` with PROBE_ITERS=1 PROBE_ONLY=150x250)."""
import json
import os
import sys

import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasterrcnn_b200 import ops  # noqa: E402

SHAPES = [("600x1000 64->64", 600, 1000, 64, 64), ("300x500 128->128", 300, 500, 128, 128), ("150x250 256->256", 150, 250, 256, 256),
          ("75x125 512->512", 75, 125, 512, 512), ("37x62 512->512", 37, 62, 512, 512)]


def main():
  iters = int(os.environ.get("PROBE_ITERS", "10"))
  only = os.environ.get("PROBE_ONLY")
  out = []
  for name, h, w, cin, cout in SHAPES:
    if only and only not in name:
      continue
    x = ops.as_nhwc(t.randn((1, cin, h, w), device = "cuda"))
    wt = ops.as_nhwc(t.randn((cout, cin, 3, 3), device = "cuda") * 0.05)
    b = t.zeros((cout,), device = "cuda")
    dy = ops.as_nhwc(t.randn((1, cout, h, w), device = "cuda"))
    gflop = 2e-9 * h * w * cout * 9 * cin
    ops.begin_step()
    xs, ws, ds = ops.tf32_split(x), ops.tf32_split(wt), ops.tf32_split(dy)
    geom = (1, h, w, cin, cout, 3, 3, 1, 1)
    y = ops._empty_nhwc(1, cout, h, w, x.device)
    dx = ops._empty_nhwc(1, cin, h, w, x.device)
    dw = t.empty((cout, cin, 3, 3), device = "cuda", memory_format = t.channels_last)
    runs = {"fwd": lambda: ops._gemm(0, x, wt, y, geom, "probe", gflop, xs, ws, bias = b, act = ops.ACT_RELU),
            "dgrad": lambda: ops._gemm(1, dy, wt, dx, geom, "probe", gflop, ds, ws),
            "wgrad": lambda: ops._gemm(2, dy, x, dw, geom, "probe", gflop, ds, xs)}
    for kind, fn in runs.items():
      if kind != "fwd" and cin == 64:
        continue                                   # frozen block: no backward in the model
      for _ in range(2 if iters > 1 else 0):
        fn()
      a, e = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
      a.record()
      for _ in range(iters):
        fn()
      e.record()
      t.cuda.synchronize()
      ms = a.elapsed_time(e) / iters
      out.append(dict(shape = name, op = kind, ms = ms, tflops = gflop / ms))
      print("%-18s %-6s %8.3f ms  %6.1f TFLOP/s" % (name, kind, ms, gflop / ms), flush = True)
  # the detector's fully connected layers on 128 RoIs: fc1 (25088 -> 4096) and fc2 (4096 -> 4096); wgrad writes 411 MB / 67 MB
  for name, k, n in (("fc1 128x25088->4096", 25088, 4096), ("fc2 128x4096->4096", 4096, 4096)):
    if only and only not in name:
      continue
    m = 128
    x = t.randn((m, k), device = "cuda"); wt = t.randn((n, k), device = "cuda") * 0.01; dy = t.randn((m, n), device = "cuda")
    y = t.empty((m, n), device = "cuda"); dx = t.empty((m, k), device = "cuda"); dw = t.empty((n, k), device = "cuda")
    gflop = 2e-9 * m * k * n
    ops.begin_step()
    xs, ws, ds = ops.tf32_split(x), ops.tf32_split(wt), ops.tf32_split(dy)
    geom = (m, 1, 1, k, n, 1, 1, 1, 0)
    runs = {"fwd": lambda: ops._gemm(0, x, wt, y, geom, "probe", gflop, xs, ws),
            "dgrad": lambda: ops._gemm(1, dy, wt, dx, geom, "probe", gflop, ds, ws),
            "wgrad": lambda: ops._gemm(2, dy, x, dw, geom, "probe", gflop, ds, xs)}
    for kind, fn in runs.items():
      for _ in range(2 if iters > 1 else 0):
        fn()
      a, e = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
      a.record()
      for _ in range(iters):
        fn()
      e.record()
      t.cuda.synchronize()
      ms = a.elapsed_time(e) / iters
      hbm = (wt.numel() * 4 if kind != "wgrad" else dw.numel() * 4) / ms / 1e6
      out.append(dict(shape = name, op = kind, ms = ms, tflops = gflop / ms, weight_side_GBs = hbm))
      print("%-22s %-6s %8.3f ms  %6.1f TFLOP/s  %6.0f GB/s on the weight-sized operand" % (name, kind, ms, gflop / ms, hbm), flush = True)
  print(json.dumps(dict(pair = os.environ.get("FRCNN_TC_PAIR", "default"), rows = out)))


if __name__ == "__main__":
  main()
