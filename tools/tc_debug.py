"""Development probe: decode how the tcgen05 MN-major descriptors address shared memory by running
one-hot crafted dgrad / wgrad problems.  Usage: python tools/tc_debug.py"""
import os
import subprocess
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
  from fasterrcnn_b200 import ops
  ops.set_engine("tc")
  dev = "cuda"
  # ---- DGRAD 1x1: dx[p][ci] = sum_co dy[p][co] w[co][ci];  dy one-hot co = p % 32  ->  dx[p][ci] = w[p%32][ci]
  H, W, Cout, Cin = 8, 16, 32, 64
  p = t.arange(H * W, device = dev)
  dy = t.zeros((H * W, Cout), device = dev)
  dy[p, p % Cout] = 1.0
  w = (t.arange(Cout, device = dev).float()[:, None] * 100 + t.arange(Cin, device = dev).float()[None, :])      # w[co][ci] = 100 co + ci
  dy4 = dy.reshape(1, H, W, Cout).permute(0, 3, 1, 2)
  w4 = w.reshape(Cout, Cin, 1, 1)
  dx = ops.conv2d_dgrad_raw(ops.as_nhwc(dy4), w4, (1, Cin, H, W), 1, 0)
  got = dx.permute(0, 2, 3, 1).reshape(H * W, Cin)
  exp = w[p % Cout]
  print("DGRAD match:", bool(t.equal(got, exp)), " nonzero:", int((got != 0).sum()), "/", got.numel())
  for r in (0, 1, 2, 33, 127):
    print("  row %3d got" % r, got[r, :10].int().tolist(), " exp", exp[r, :4].int().tolist())
  # which (co, ci) did each output come from?  value = 100 co + ci
  g = got[:40, :40].cpu().numpy()
  print("  decoded co of dx[0:6, 0:6]:\n", (g[:6, :6] // 100).astype(int), "\n  decoded ci:\n", (g[:6, :6] % 100).astype(int))

  # ---- WGRAD 1x1: dW[co][ci] = sum_p dy[p][co] x[p][ci]; dy one-hot p = co % 64 -> dW[co][ci] = x[co%64][ci]
  H, W, Cout, Cin = 8, 8, 128, 64
  npix = H * W
  co = t.arange(Cout, device = dev)
  dy = t.zeros((npix, Cout), device = dev)
  dy[co % npix, co] = 1.0
  x = (t.arange(npix, device = dev).float()[:, None] * 100 + t.arange(Cin, device = dev).float()[None, :])      # x[p][ci] = 100 p + ci
  dy4 = dy.reshape(1, H, W, Cout).permute(0, 3, 1, 2)
  x4 = x.reshape(1, H, W, Cin).permute(0, 3, 1, 2)
  dw = ops.conv2d_wgrad_raw(ops.as_nhwc(dy4), ops.as_nhwc(x4), (Cout, Cin, 1, 1), 1, 0).reshape(Cout, Cin)
  exp = x[co % npix]
  print("WGRAD match:", bool(t.equal(dw, exp)), " nonzero:", int((dw != 0).sum()), "/", dw.numel())
  for r in (0, 1, 33, 65, 127):
    print("  row %3d got" % r, dw[r, :10].int().tolist(), " exp", exp[r, :4].int().tolist())

  # ---- FWD accuracy vs fp64 (hi/lo split check)
  g = t.Generator().manual_seed(29)
  xx = t.randn((128, 25088), generator = g).cuda()
  wt = (t.randn((4096, 25088), generator = g) * (1.0 / 25088) ** 0.5).cuda()
  y64 = xx.double().cpu() @ wt.double().cpu().t()
  y = ops.linear_act(xx, wt, None, ops.ACT_NONE)
  ops.set_engine("simt")
  ys = ops.linear_act(xx, wt, None, ops.ACT_NONE)
  print("FWD K=25088 max err vs fp64: tc %.3e  simt %.3e  (scale %.2f)" % (float((y.double().cpu() - y64).abs().max()), float((ys.double().cpu() - y64).abs().max()), float(y64.abs().max())))


if __name__ == "__main__":
  if len(sys.argv) > 1 and sys.argv[1] == "child":
    child()
  else:
    combos = [dict(), dict(FRCNN_TC_MN_LBO = "4096", FRCNN_TC_MN_SBO = "1024"), dict(FRCNN_TC_MN_LBO = "512", FRCNN_TC_MN_SBO = "4096")]
    for c in combos:
      print("=" * 20, c or "default (LBO 4096, SBO 512)", flush = True)
      env = dict(os.environ); env.update(c)
      r = subprocess.run([sys.executable, __file__, "child"], env = env, capture_output = True, text = True, timeout = 300)
      print(r.stdout[-3500:])
      if r.returncode != 0:
        print("rc", r.returncode, r.stderr[-800:])
