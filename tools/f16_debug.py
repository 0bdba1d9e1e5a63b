"""fp16 engine bring-up: per-pass max error vs the CUDA-core engine on a few shapes, printed instead of asserted (so one GPU call shows
which of fwd / dgrad / wgrad (K-major / MN-major descriptors) is off), plus a timing comparison tf32 vs fp16 on VGG layer shapes."""
import os
import sys
import time

import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasterrcnn_b200 import ops  # noqa: E402

CASES = [("b5_512_37x62", 1, 37, 62, 512, 512, 3, 1), ("b2_64to128_60x100", 1, 60, 100, 64, 128, 3, 1), ("pw_512to128", 1, 37, 62, 512, 128, 1, 0),
         ("b3_256_150x250", 1, 150, 250, 256, 256, 3, 1), ("b1_64_600x1000", 1, 600, 1000, 64, 64, 3, 1)]


def run(engine, fn):
  ops.set_engine(engine)
  try:
    return fn()
  except Exception as e:                                             # noqa: BLE001
    return e


def main():
  g = t.Generator().manual_seed(1)
  for name, n, h, w, cin, cout, k, pad in CASES:
    x = ops.as_nhwc((t.randn((n, cin, h, w), generator = g) * 3.0).cuda())
    dy = ops.as_nhwc((t.randn((n, cout, h, w), generator = g) * 0.01).cuda())
    wt = (t.randn((cout, cin, k, k), generator = g) * (2.0 / (cin * k * k)) ** 0.5).cuda().contiguous(memory_format = t.channels_last)
    big = h * w > 100000
    passes = dict(fwd = lambda: ops.conv2d_fwd_raw(x, wt, None, 1, pad, ops.ACT_NONE), dgrad = lambda: ops.conv2d_dgrad_raw(dy, wt, (n, cin, h, w), 1, pad),
                  wgrad = lambda: ops.conv2d_wgrad_raw(dy, x, (cout, cin, k, k), 1, pad))
    for pname, fn in passes.items():
      ref = run("tc" if big else "simt", fn)
      got = run("f16", fn)
      t.cuda.synchronize()
      if isinstance(got, Exception) or isinstance(ref, Exception):
        print(name, pname, "ERROR", got if isinstance(got, Exception) else ref, flush = True)
        continue
      err = float((got - ref).abs().max()); sc = float(ref.abs().max())
      print("%-18s %-6s rel err %.3e  (max %.3e)" % (name, pname, err / sc, sc), flush = True)
      # timing, operands pre-split by the cache (reuse flags off -> includes internal split; so time the raw entry with cached splits)
      for eng in ("tc", "f16"):
        ops.set_engine(eng)
        for _ in range(3):
          fn()
        t.cuda.synchronize()
        e0, e1 = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
        e0.record()
        for _ in range(10):
          fn()
        e1.record(); t.cuda.synchronize()
        print("    %-4s %.1f us/launch (incl. its operand splits)" % (eng, e0.elapsed_time(e1) * 100), flush = True)
  ops.set_engine("auto")


if __name__ == "__main__":
  main()
