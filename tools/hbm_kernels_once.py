"""Runs each HBM-bound kernel of the path twice (warm-up + one launch to profile) at the config-5 microbench shapes, for an
`ncu --set full` capture:  ncu --set full --clock-control none -k regex:'roi_|nms_|rpn_decode|rank_|sgd_' -o gpurun_out/prof_hbm python tools/hbm_kernels_once.py
Prints the algorithmic bytes per launch that tools/summarize_hbm_ncu.py divides by the measured durations."""
import json
import os
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasterrcnn_b200 import ops  # noqa: E402
from tools.microbench import boxes  # noqa: E402


def main():
  rng = np.random.default_rng(0)
  alg = {}
  N, C, H, W = 6000, 512, 37, 62
  fm = ops.as_nhwc(t.relu(t.randn((1, C, H, W), device = "cuda")))
  props = t.from_numpy(boxes(rng, N)).cuda()
  for _ in range(2):
    ops.roi_pool(fm, props)
    ops.roi_align(fm, props, (7, 7), 1.0 / 16.0, 2, False)
  alg["roi_pool_fwd_v4_kernel"] = fm.numel() * 4 + 20 * N + 196 * C * N * 2
  alg["roi_align_fwd_v4_kernel"] = fm.numel() * 4 + 20 * N + 196 * C * N
  fmg = fm.clone().requires_grad_(True)
  p128 = props[:128].contiguous()
  y = ops.roi_pool(fmg, p128)
  g = t.randn_like(y)
  for _ in range(2):
    y.backward(g, retain_graph = True)
  alg["roi_pool_bwd_kernel"] = 2 * 196 * C * 128 + fm.numel() * 4
  n = 6000
  bb = t.stack([t.from_numpy(boxes(rng, n)) for _ in range(20)]).cuda()
  sb = t.stack([t.from_numpy(rng.permutation(n).astype(np.float32) / n) for _ in range(20)]).cuda()
  for _ in range(2):
    ops.nms_batched(bb, sb, 0.3)
  tiles = 20 * n * ((n + 63) // 64) * 8 // 2
  alg["nms_mask_kernel"] = 20 * n * 16 + tiles
  alg["nms_scan_kernel"] = tiles
  alg["rank_count_kernel"] = 20 * n * 8
  fh, fw = 592, 992
  d = t.randn((fh * fw * 9, 4), device = "cuda") * 0.3
  for _ in range(2):
    ops.rpn_decode(d, fh, fw, 16, fh * 16, fw * 16)
  alg["rpn_decode_kernel"] = fh * fw * 9 * 33
  p = t.randn((4096, 25088), device = "cuda"); gr = t.randn_like(p); buf = t.zeros_like(p)
  for _ in range(2):
    ops.sgd_step(p, gr, buf, 1e-3, 0.9, 5e-4)
  alg["sgd_kernel"] = p.numel() * 20
  t.cuda.synchronize()
  print(json.dumps(alg))


if __name__ == "__main__":
  main()
