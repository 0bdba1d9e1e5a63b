"""
Parity margin recorder (test infrastructure).  The end-to-end GPU tests assert ceilings; what they MEASURE -- largest deviation,
unmatched rows, distance of the discontinuous decisions from their thresholds -- is appended here, one JSON object per line, to
gpurun_out/parity_margins_<test module>.jsonl (come back from the GPU box) and summarised by tools/summarize_margins.py into
profiles/r02_parity_margins.md.  The bars in the tests are set from those measurements (2x the measured value).
"""
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _path():
  """One file per test module (gpurun merges files by name: the two-GPU run's records must survive the next one-GPU run)."""
  module = os.path.splitext(os.path.basename((os.environ.get("PYTEST_CURRENT_TEST") or "unknown").split("::")[0]))[0]
  return os.path.join(ROOT, "gpurun_out", "parity_margins_%s.jsonl" % module)


def record(test, **metrics):
  row = dict(test = test, when = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()))
  for k, v in metrics.items():
    if isinstance(v, (np.floating, np.integer)):
      v = v.item()
    row[k] = v
  try:
    path = _path()
    os.makedirs(os.path.dirname(path), exist_ok = True)
    with open(path, "a") as f:
      f.write(json.dumps(row) + "\n")
  except OSError:
    pass
  print("[margins] " + json.dumps(row))
  return row


def match_rows(got, ref):
  """Nearest oracle row (max-abs over the 4 coordinates) for every produced row: (partner index, distance in px)."""
  if got.shape[0] == 0 or ref.shape[0] == 0:
    return np.zeros((got.shape[0],), dtype = np.int64), np.full((got.shape[0],), np.inf)
  dist = np.abs(got[:, None, :] - ref[None, :, :]).max(axis = 2)
  partner = dist.argmin(axis = 1)
  return partner, dist[np.arange(got.shape[0]), partner]


def proposal_margins(got, ref, tol):
  """How far the produced proposal set is from the oracle's: rows matched within `tol` px, the largest deviation among them, rows without
  a partner (a flipped discontinuous decision upstream), row-count difference."""
  partner, dist = match_rows(got, ref)
  ok = dist <= tol
  return partner, ok, dict(rows = int(got.shape[0]), ref_rows = int(ref.shape[0]), matched_frac = float(ok.mean()) if ok.size else 1.0,
                           unmatched_rows = int((~ok).sum()), max_px_matched = float(dist[ok].max()) if ok.any() else 0.0,
                           median_px_matched = float(np.median(dist[ok])) if ok.any() else 0.0)


def decision_margins(taps, pre_nms, image_hw, iou_threshold = 0.7, min_size = 16.0):
  """Distance of the oracle's own discontinuous decisions from their thresholds (SURVEY.md 7 'margin analysis'): a decision whose
  margin is below the upstream floating-point agreement can legitimately flip on another summation order.
    topn_cut_gap      score gap between the last anchor inside the top-N cut and the first one outside (inf when nothing is cut)
    min_adjacent_gap  smallest gap between consecutive scores inside the cut (an order swap moves NMS priority)
    min_size_margin   min | side - 16 | over the clipped boxes entering the size filter
    min_iou_margin    min | IoU - thr | over (kept box, later box) pairs of the NMS input -- the pairs the greedy chain evaluates
  """
  scores = np.sort(taps["score_map"].detach().numpy().reshape(-1).astype(np.float64))[::-1]
  out = {}
  out["topn_cut_gap"] = float(scores[pre_nms - 1] - scores[pre_nms]) if scores.shape[0] > pre_nms else float("inf")
  top = scores[:min(pre_nms, scores.shape[0])]
  out["min_adjacent_gap"] = float(np.min(top[:-1] - top[1:])) if top.shape[0] > 1 else float("inf")
  boxes = taps["pre_nms_boxes"].astype(np.float64)                          # clipped, size-filtered, score order
  if boxes.shape[0]:
    sides = np.stack([boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]], axis = 1)
    out["min_size_margin_survivors"] = float(np.abs(sides - min_size).min())
    keep = np.asarray(taps["nms_keep"], dtype = np.int64)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    best = np.inf
    for i0 in range(0, keep.shape[0], 256):                                  # (kept, all later boxes) in slabs
      k = keep[i0:i0 + 256]
      kb = boxes[k]
      y1 = np.maximum(kb[:, None, 0], boxes[None, :, 0]); x1 = np.maximum(kb[:, None, 1], boxes[None, :, 1])
      y2 = np.minimum(kb[:, None, 2], boxes[None, :, 2]); x2 = np.minimum(kb[:, None, 3], boxes[None, :, 3])
      inter = np.clip(y2 - y1, 0, None) * np.clip(x2 - x1, 0, None)
      iou = inter / (area[k][:, None] + area[None, :] - inter + 1e-300)
      later = np.arange(boxes.shape[0])[None, :] > k[:, None]
      m = np.abs(iou - iou_threshold)[later]
      if m.size:
        best = min(best, float(m.min()))
    out["min_iou_margin"] = best
    out["nms_in"] = int(boxes.shape[0]); out["nms_kept"] = int(keep.shape[0])
  return out


def grad_margins(grads, ref_grads):
  """Worst and median relative L2 error over the parameter gradients (+ which tensor is worst)."""
  rels = {}
  for k, b in ref_grads.items():
    if k not in grads:
      continue
    a, b = grads[k].double(), b.double()
    rels[k] = float((a - b).norm() / (b.norm() + 1e-12))
  worst = max(rels, key = rels.get)
  return rels, dict(worst_grad_rel_l2 = rels[worst], worst_grad = worst, median_grad_rel_l2 = float(np.median(list(rels.values()))), grads = len(rels))
