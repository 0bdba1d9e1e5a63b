"""Development tool: per-CTA %globaltimer trace of the tcgen05 conv kernel (frcnn_debug_tc_trace) for a few VGG-16 layers.
Prints, per layer and pass, the kernel time (CUDA events) and the distribution over CTAs of: setup, pipeline fill, first item's
mainloop, drain, store, and exit, all relative to the earliest CTA entry.  Usage: python tools/tc_trace.py"""
import os
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasterrcnn_b200 import ops
from fasterrcnn_b200._lib import lib, ptr

NAMES = ["entry", "setup", "tma0", "mma0", "item0_issued", "all_issued", "item0_drained", "item0_stored", "epi_done", "exit"]


def ev(fn, iters = 10, warm = 3):
  for _ in range(warm):
    fn()
  t.cuda.synchronize()
  a, b = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  t.cuda.synchronize()
  return a.elapsed_time(b) / iters * 1e3


def trace(fn):
  buf = t.zeros((148 * 16,), dtype = t.int64, device = "cuda")
  t.cuda.synchronize()
  lib().frcnn_debug_tc_trace(ptr(buf))
  fn()
  t.cuda.synchronize()
  lib().frcnn_debug_tc_trace(None)
  tr = buf.cpu().numpy().reshape(148, 16)
  tr = tr[tr[:, 0] > 0]
  if tr.shape[0] == 0:
    return None, None                                    # this pass did not run on the tcgen05 engine
  t0 = tr[:, 0].min()
  rel = (tr[:, :10] - t0) / 1e3
  rel[tr[:, :10] == 0] = np.nan
  return rel, tr[:, 10]


def main():
  layers = [(64, 64, 600, 1000), (64, 128, 300, 500), (128, 128, 300, 500), (256, 256, 150, 250), (512, 512, 75, 125), (512, 512, 37, 62)]
  for cin, cout, h, w in layers:
    x = ops.as_nhwc(t.randn((1, cin, h, w), device = "cuda"))
    wt = t.randn((cout, cin, 3, 3), device = "cuda").contiguous(memory_format = t.channels_last)
    b = t.zeros((cout,), device = "cuda")
    dy = ops.as_nhwc(t.randn((1, cout, h, w), device = "cuda"))
    gf = 2 * 9 * cin * cout * h * w / 1e9
    passes = [("fwd", lambda: ops.conv2d_fwd_raw(x, wt, b, 1, 1, ops.ACT_RELU)),
              ("dgrad", lambda: ops.conv2d_dgrad_raw(dy, wt, (1, cin, h, w), 1, 1)),
              ("wgrad", lambda: ops.conv2d_wgrad_raw(dy, x, (cout, cin, 3, 3), 1, 1))]
    for name, fn in passes:
      us = ev(fn)
      rel, smid = trace(fn)
      if rel is None:
        print("conv %4d->%4d @%4dx%4d %-5s %8.1f us  %6.1f TF/s  (CUDA-core engine, no trace)" % (cin, cout, h, w, name, us, gf / us * 1e3), flush = True)
        continue
      print("conv %4d->%4d @%4dx%4d %-5s %8.1f us  %6.1f TF/s  ctas %d" % (cin, cout, h, w, name, us, gf / us * 1e3, rel.shape[0]), flush = True)
      med = np.nanmedian(rel, axis = 0)
      mx = np.nanmax(rel, axis = 0)
      mn = np.nanmin(rel, axis = 0)
      print("    " + "  ".join("%s %.1f/%.1f/%.1f" % (NAMES[i], mn[i], med[i], mx[i]) for i in range(10)) + "   (min/median/max us since first CTA entry)")


if __name__ == "__main__":
  main()
