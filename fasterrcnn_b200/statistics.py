"""
Evaluation/training statistics with the reference's interface: pytorch/FasterRCNN/statistics.py
(TrainingStatistics :16-62, PrecisionRecallCurveCalculator :65-257).  Host-side bookkeeping over the few boxes predict()
returns per image (tens), so it stays on the host; the detections themselves come from the device kernel
frcnn_detect_postprocess (ops.detect_postprocess).

Matching semantics kept from the reference, including one that looks accidental: the (IoU, box, gt) triples are "sorted" by
`key = lambda iou: ious[0]` (statistics.py:99), a constant key, so Python's stable sort leaves them in construction order --
ground-truth-major, prediction-minor -- and the greedy assignment visits them in THAT order, not by descending IoU.  The mAP
numbers the reference publishes were produced this way, so this mirror reproduces it (tests/golden/statistics.npz holds the
reference's outputs on seeded detections).
"""
from collections import defaultdict

import numpy as np


def _pair_iou(boxes, gt):
  """IoU of every prediction (P,4) with every ground truth (G,4) -> (G,P); arithmetic of math_utils.py:29-37 in the operand dtypes."""
  top_left = np.maximum(boxes[None, :, 0:2], gt[:, None, 0:2])
  bottom_right = np.minimum(boxes[None, :, 2:4], gt[:, None, 2:4])
  ordered = np.all(top_left < bottom_right, axis = 2)
  inter = ordered * np.prod(bottom_right - top_left, axis = 2)
  areas_b = np.prod(boxes[:, 2:4] - boxes[:, 0:2], axis = 1)
  areas_g = np.prod(gt[:, 2:4] - gt[:, 0:2], axis = 1)
  union = areas_b[None, :] + areas_g[:, None] - inter
  return inter / (union + 1e-7)


class TrainingStatistics:
  """Running means of the four losses over an epoch (statistics.py:16-62)."""
  def __init__(self):
    self.rpn_class_loss = float("inf")
    self.rpn_regression_loss = float("inf")
    self.detector_class_loss = float("inf")
    self.detector_regression_loss = float("inf")
    self._sums = np.zeros(4, dtype = np.float64)
    self._losses = [[], [], [], []]

  def on_training_step(self, loss):
    for acc, v in zip(self._losses, (loss.rpn_class, loss.rpn_regression, loss.detector_class, loss.detector_regression)):
      acc.append(v)
    self.rpn_class_loss = np.mean(self._losses[0])
    self.rpn_regression_loss = np.mean(self._losses[1])
    self.detector_class_loss = np.mean(self._losses[2])
    self.detector_regression_loss = np.mean(self._losses[3])

  def get_progbar_postfix(self):
    return {
      "rpn_class_loss": "%1.4f" % self.rpn_class_loss,
      "rpn_regr_loss": "%1.4f" % self.rpn_regression_loss,
      "detector_class_loss": "%1.4f" % self.detector_class_loss,
      "detector_regr_loss": "%1.4f" % self.detector_regression_loss,
      "total_loss": "%1.2f" % (self.rpn_class_loss + self.rpn_regression_loss + self.detector_class_loss + self.detector_regression_loss)
    }


class PrecisionRecallCurveCalculator:
  """Accumulates per-image detections and computes AP / mAP (statistics.py:65-257)."""
  def __init__(self):
    self._unsorted_predictions_by_class_index = defaultdict(list)   # class -> [(score, is_true_positive)]
    self._object_count_by_class_index = defaultdict(int)

  def _compute_correctness_of_predictions(self, scored_boxes_by_class_index, gt_boxes):
    predictions = {}
    counts = defaultdict(int)
    for gt_box in gt_boxes:
      counts[gt_box.class_index] += 1
    for class_index, scored_boxes in scored_boxes_by_class_index.items():
      scored_boxes = np.asarray(scored_boxes)
      num = len(scored_boxes)
      is_tp = [False] * num
      gts = [g.corners for g in gt_boxes if g.class_index == class_index]
      if num > 0 and len(gts) > 0:
        iou = _pair_iou(scored_boxes[:, 0:4], np.stack(gts, axis = 0))    # (G,P): row-major walk == the reference's list order
        gt_done = [False] * len(gts)
        for g, p in zip(*np.nonzero(iou > 0.5)):                            # nonzero is row-major: gt-major, prediction-minor
          if is_tp[p] or gt_done[g]:
            continue
          is_tp[p] = True
          gt_done[g] = True
      predictions[class_index] = [(scored_boxes[i][4], is_tp[i]) for i in range(num)]
    return predictions, counts

  def add_image_results(self, scored_boxes_by_class_index, gt_boxes):
    predictions, counts = self._compute_correctness_of_predictions(scored_boxes_by_class_index = scored_boxes_by_class_index, gt_boxes = gt_boxes)
    for class_index, preds in predictions.items():
      self._unsorted_predictions_by_class_index[class_index] += preds
    for class_index, count in counts.items():
      self._object_count_by_class_index[class_index] += count

  def _compute_average_precision(self, class_index):
    preds = sorted(self._unsorted_predictions_by_class_index[class_index], key = lambda p: p[0], reverse = True)
    num_positives = self._object_count_by_class_index[class_index]
    correct = np.array([1 if p[1] == True else 0 for p in preds], dtype = np.int64)
    tp = np.cumsum(correct)
    fp = np.cumsum(1 - correct)
    recall = [0.0] + [int(a) / num_positives for a in tp] + [1.0]
    precision = [0.0] + [int(a) / (int(a) + int(b)) for a, b in zip(tp, fp)] + [0.0]
    # interpolation: highest precision from each point onward (statistics.py:193-194)
    precision = list(np.maximum.accumulate(np.array(precision, dtype = np.float64)[::-1])[::-1])
    average_precision = 0
    for i in range(len(recall) - 1):
      average_precision += precision[i + 1] * (recall[i + 1] - recall[i])
    return average_precision, recall, precision

  def compute_mean_average_precision(self):
    aps = [self._compute_average_precision(class_index = c)[0] for c in self._object_count_by_class_index]
    return np.mean(aps)

  def compute_class_average_precisions(self):
    """class index -> AP (the per-class table print_average_precisions shows, statistics.py:234-257)."""
    return {c: self._compute_average_precision(class_index = c)[0] for c in self._object_count_by_class_index}

  def print_average_precisions(self, class_index_to_name):
    labels = [class_index_to_name[c] for c in self._object_count_by_class_index]
    aps = self.compute_class_average_precisions()
    width = max([len(s) for s in labels] + [1])
    for (label, ap) in sorted(zip(labels, [aps[c] for c in self._object_count_by_class_index]), reverse = True, key = lambda pair: pair[1]):
      print("%s: %1.1f%%" % (label.ljust(width), ap * 100.0))
