"""
Config-5 style microbenchmarks of the HBM-bound kernels at a scale where bandwidth (not launch
latency) is measured: RoIPool fwd/bwd with 6000 RoIs, NMS 6000 boxes x 20 classes and 12000 x 1,
RPN decode, fused SGD.  Prints one JSON object per kernel with the achieved GB/s (algorithmic bytes
of SURVEY.md 8d / CUDA-event time) against MEASURED_PEAKS.json's copy bandwidth.
"""
import json
import os
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasterrcnn_b200 import ops  # noqa: E402


def peak_gbs():
  p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
  return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def ev_time(fn, iters = 20, warm = 3):
  flush = t.empty((160 * 1024 * 1024,), dtype = t.uint8, device = "cuda")   # > 126 MB L2
  for _ in range(warm):
    fn()
  times = []
  for _ in range(iters):
    flush.zero_()                                                            # evict L2 between iterations
    a, b = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
    a.record(); fn(); b.record()
    t.cuda.synchronize()
    times.append(a.elapsed_time(b))
  return float(np.median(times))


def boxes(rng, n, h = 600.0, w = 1000.0):
  y1 = rng.uniform(0, h, n); x1 = rng.uniform(0, w, n)
  return np.stack([y1, x1, np.minimum(y1 + rng.uniform(16, 300, n), h), np.minimum(x1 + rng.uniform(16, 300, n), w)], axis = 1).astype(np.float32)


def run():
  """-> list of dicts, one per kernel (bench.py --micro prints them as ONE JSON line; python tools/microbench.py one line each)."""
  rng = np.random.default_rng(0)
  peak = peak_gbs()
  out = []
  for C, H, W in ((512, 37, 62), (1024, 38, 63)):
    N = 6000
    fm = ops.as_nhwc(t.relu(t.randn((1, C, H, W), device = "cuda")))
    props = t.from_numpy(boxes(rng, N)).cuda()
    ms = ev_time(lambda: ops.roi_pool(fm, props))
    alg = fm.numel() * 4 + 20 * N + 196 * C * N * 2          # output + argmax
    out.append(dict(kernel = "roi_pool_fwd", shape = "N=%d C=%d fm=%dx%d" % (N, C, H, W), ms = ms, algorithmic_MB = alg / 1e6, achieved_GBs = alg / ms / 1e6, peak_GBs = peak, frac = alg / ms / 1e6 / peak))
    ms = ev_time(lambda: ops.roi_align(fm, props, (7, 7), 1.0 / 16.0, 2, False))
    alg = fm.numel() * 4 + 20 * N + 196 * C * N                # output only (no argmax)
    out.append(dict(kernel = "roi_align_fwd (sampling_ratio 2)", shape = "N=%d C=%d fm=%dx%d" % (N, C, H, W), ms = ms, algorithmic_MB = alg / 1e6, achieved_GBs = alg / ms / 1e6, peak_GBs = peak, frac = alg / ms / 1e6 / peak))
  # training-size RoIPool forward (128 RoIs: the size inside train_step; latency, not bandwidth)
  fm = ops.as_nhwc(t.relu(t.randn((1, 512, 37, 62), device = "cuda")))
  props = t.from_numpy(boxes(rng, 128)).cuda()
  ms = ev_time(lambda: ops.roi_pool(fm, props))
  alg = fm.numel() * 4 + 20 * 128 + 196 * 512 * 128 * 2
  out.append(dict(kernel = "roi_pool_fwd", shape = "N=128 C=512 fm=37x62", ms = ms, algorithmic_MB = alg / 1e6, achieved_GBs = alg / ms / 1e6, peak_GBs = peak, frac = alg / ms / 1e6 / peak))
  # training-size RoIPool backward (128 RoIs)
  fm = ops.as_nhwc(t.relu(t.randn((1, 512, 37, 62), device = "cuda"))).requires_grad_(True)
  props = t.from_numpy(boxes(rng, 128)).cuda()
  y = ops.roi_pool(fm, props)
  g = t.randn_like(y)
  ms = ev_time(lambda: y.backward(g, retain_graph = True))
  alg = 2 * 196 * 512 * 128 + fm.numel() * 4
  out.append(dict(kernel = "roi_pool_bwd (deterministic)", shape = "N=128 C=512 fm=37x62", ms = ms, algorithmic_MB = alg / 1e6, achieved_GBs = alg / ms / 1e6, peak_GBs = peak, frac = alg / ms / 1e6 / peak))
  # NMS
  for n, thr in ((6000, 0.3), (12000, 0.7)):
    b = t.from_numpy(boxes(rng, n)).cuda(); s = t.from_numpy(rng.permutation(n).astype(np.float32) / n).cuda()
    ms = ev_time(lambda: ops.nms(b, s, thr), iters = 10)
    pairs = n * (n - 1) / 2
    out.append(dict(kernel = "nms (rank + mask + scan, incl. 1 host sync)", shape = "N=%d thr=%.1f" % (n, thr), ms = ms, iou_pairs_per_s = pairs / ms * 1e3, algorithmic_MB = n * 20 / 1e6, achieved_GBs = n * 20 / ms / 1e6))
  n = 6000
  bs = [t.from_numpy(boxes(rng, n)).cuda() for _ in range(20)]; ss = [t.from_numpy(rng.permutation(n).astype(np.float32) / n).cuda() for _ in range(20)]
  ms = ev_time(lambda: [ops.nms(bs[i], ss[i], 0.3) for i in range(20)], iters = 5)
  out.append(dict(kernel = "nms x 20 classes", shape = "20 x N=6000 thr=0.3", ms = ms, iou_pairs_per_s = 20 * n * (n - 1) / 2 / ms * 1e3, algorithmic_MB = 20 * n * 20 / 1e6))
  bb = t.stack(bs); sb = t.stack(ss)
  ms = ev_time(lambda: ops.nms_batched(bb, sb, 0.3), iters = 10)
  pairs = 20 * n * (n - 1) / 2
  mask_bytes = 20 * n * ((n + 63) // 64) * 8 / 2               # upper-triangle bit tiles written by the mask kernel, read by the scan
  out.append(dict(kernel = "nms_batched (rank + sort + mask + scan + finish, no host sync)", shape = "20 x N=6000 thr=0.3", ms = ms, iou_pairs_per_s = pairs / ms * 1e3,
                  algorithmic_MB = 20 * n * 20 / 1e6, bit_tile_MB = mask_bytes / 1e6, achieved_GBs = (20 * n * 20 + mask_bytes) / ms / 1e6, peak_GBs = peak))
  # fused anchor + decode + clip + size flag at a scale where bandwidth shows: a 592 x 992 cell map = 5.29 M anchors, 37 B / anchor
  # (16 B deltas in, 16 B box + 1 B flag out; the score's 4 B in / 4 B out of SURVEY 8d belong to the rank / gather kernels)
  fh, fw = 592, 992
  deltas_big = t.randn((fh * fw * 9, 4), device = "cuda") * 0.3
  ms = ev_time(lambda: ops.rpn_decode(deltas_big, fh, fw, 16, fh * 16, fw * 16))
  alg = fh * fw * 9 * 33
  out.append(dict(kernel = "rpn_decode (anchors regenerated in-kernel)", shape = "%d anchors" % (fh * fw * 9), ms = ms, algorithmic_MB = alg / 1e6, achieved_GBs = alg / ms / 1e6, peak_GBs = peak, frac = alg / ms / 1e6 / peak))
  # RPN decode at 600x1000 (20,646 anchors: latency bound) -- 40 B/anchor
  fh, fw = 37, 62
  deltas = t.randn((1, fh, fw, 36), device = "cuda") * 0.3; scores = t.rand((1, fh, fw, 9), device = "cuda")
  ms = ev_time(lambda: ops.rpn_proposals(scores, deltas, (3, 600, 1000), 16, 12000, 2000), iters = 10)
  out.append(dict(kernel = "rpn_proposals (decode+rank+filter+nms+gather, 1 host sync)", shape = "20646 anchors, 12000 -> 2000", ms = ms))
  # fused SGD on the fc1-sized tensor: 5 accesses x 4 B / element
  p = t.randn((4096, 25088), device = "cuda"); gr = t.randn_like(p); buf = t.zeros_like(p)
  ms = ev_time(lambda: ops.sgd_step(p, gr, buf, 1e-3, 0.9, 5e-4))
  alg = p.numel() * 20
  out.append(dict(kernel = "sgd_kernel", shape = "102.8 M elements (fc1)", ms = ms, algorithmic_MB = alg / 1e6, achieved_GBs = alg / ms / 1e6, peak_GBs = peak, frac = alg / ms / 1e6 / peak))
  return out


def main():
  for o in run():
    print(json.dumps(o))


if __name__ == "__main__":
  main()
