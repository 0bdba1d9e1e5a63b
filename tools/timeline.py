"""Development tool: GPU timeline of the VGG-16 600x1000 train step from torch.profiler (CUPTI sees the ctypes-launched kernels).
Prints GPU busy time vs step time, the largest idle gaps with the kernels on either side, and per-kernel totals (non-serialised,
warm caches -- the counterpart of the cold ncu launch list).  Usage: python tools/timeline.py [steps]"""
import os
import random
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
  dev = t.device("cuda:0")
  import argparse
  step_fn = bench.make_train_step(dev, argparse.Namespace(backbone = "vgg16", batch = 1, roi_op = "pool", rois = 128))
  for _ in range(4):
    step_fn()
  t.cuda.synchronize()
  from torch.profiler import profile, ProfilerActivity
  with profile(activities = [ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
      step_fn()
    t.cuda.synchronize()
  evs = [e for e in prof.events() if e.device_type == t.autograd.DeviceType.CUDA and e.time_range is not None]
  ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in evs if "Memcpy" not in e.name and "Memset" not in e.name or True])
  if not ks:
    print("no CUDA events captured")
    return
  span = ks[-1][1] - ks[0][0]
  busy = 0.0
  cur_s, cur_e = ks[0][0], ks[0][1]
  gaps = []
  prev_name = ks[0][2]
  for s, e, n in ks[1:]:
    if s > cur_e:
      busy += cur_e - cur_s
      gaps.append((s - cur_e, prev_name, n))
      cur_s, cur_e = s, e
    else:
      cur_e = max(cur_e, e)
    prev_name = n
  busy += cur_e - cur_s
  print("steps %d  span %.3f ms  (%.3f ms/step)  GPU busy %.3f ms (%.1f %%)  idle %.3f ms/step" %
        (steps, span / 1e3, span / 1e3 / steps, busy / 1e3, 100.0 * busy / span, (span - busy) / 1e3 / steps))
  gaps.sort(reverse = True)
  print("largest idle gaps (us)  after-kernel -> next-kernel:")
  for g, a, b in gaps[:25]:
    print("  %8.1f  %-60s -> %s" % (g, a[:60], b[:60]))
  hist = np.array([g for g, _, _ in gaps])
  for lo, hi in [(0, 2), (2, 5), (5, 10), (10, 20), (20, 50), (50, 1e9)]:
    sel = hist[(hist >= lo) & (hist < hi)]
    print("  gaps %4.0f-%-6.0f us: %5d  total %.3f ms/step" % (lo, min(hi, 9999), len(sel), sel.sum() / 1e3 / steps))
  tot = {}
  for s, e, n in ks:
    a = tot.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += e - s
  # concurrency: for every optimizer / exchange kernel, how much of its run time other kernels were running too
  for s0, e0, n0 in ks:
    if "sgd_kernel" in n0 or "dp_sgd" in n0 or "split_f16_carried" in n0:
      if e0 - s0 < 20:
        continue
      over = sum(max(0.0, min(e0, e) - max(s0, s)) for s, e, n in ks if n is not n0 and not (s == s0 and e == e0))
      names = sorted({n.split("(")[0][-40:] for s, e, n in ks if min(e0, e) - max(s0, s) > 1 and not (s == s0 and e == e0)})
      print("  overlap: %-28s %8.1f us long, %8.1f us of other kernels running meanwhile: %s" % (n0.split("(")[0][-28:], e0 - s0, over, ", ".join(names)[:160]))
  if os.environ.get("TIMELINE_DUMP"):
    t0 = ks[0][0]
    last = ks[len(ks) * (steps - 1) // steps:] if steps > 1 else ks
    print("kernels of the last step (start us, duration us, name):")
    for s, e, n in last:
      print("  %9.1f %8.1f  %s" % (s - t0, e - s, n.split("(")[0][-70:]))
  print("per-kernel totals per step (us):")
  for n, (c, us) in sorted(tot.items(), key = lambda kv: -kv[1][1])[:40]:
    print("  %9.1f  x%-4d %s" % (us / steps, c // steps, n[:110]))


if __name__ == "__main__":
  main()
