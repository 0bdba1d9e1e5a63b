"""
TEST / BENCH INFRASTRUCTURE -- never imported by the product (fasterrcnn_b200/).

Secondary comparator: the reference's train_step the way the reference itself would run it on the B200 --
eager PyTorch: cuDNN / cuBLAS convolutions and linears, ATen elementwise kernels, torchvision's CUDA nms / roi_pool, torch.optim.SGD, the
python / CPU-generator sampling with its host round trips (models/faster_rcnn.py:228-362).  It reuses the device-agnostic pieces of the
CPU restatement (oracle/frcnn_oracle.py: backbone, RPN heads, the four losses) and restates the rest with the tensors on the device.
bench.py times it as `gpu_eager_baseline` (its own leg, after the product's timed regions).  VGG-16 only.
"""
import random

import numpy as np
import torch as t
import torch.nn.functional as F

from . import frcnn_oracle as orc


def _nms(boxes, scores, thr):
  from torchvision.ops import nms
  return nms(boxes, scores, thr)


def _roi_pool(fm, rois):
  from torchvision.ops import roi_pool
  return roi_pool(fm, rois, output_size = (7, 7), spatial_scale = 1.0 / 16.0)


class EagerGpuModel:
  def __init__(self, params, device = "cuda", tf32 = False, num_classes = 21, torch_default = False):
    self.device = t.device(device)
    self.num_classes = num_classes
    self.params = {k: v.clone().to(self.device) for k, v in params.items()}
    trainable = set(orc.trainable_keys_vgg16(self.params))
    for k, v in self.params.items():
      v.requires_grad_(k in trainable)
    keys = orc.optimizer_keys([k for k in self.params if self.params[k].requires_grad])
    self.optimizer = t.optim.SGD([{"params": [self.params[k]], "weight_decay": 5e-4} for k in keys], lr = 1e-3, momentum = 0.9)   # __main__.py:98-105
    self.tf32 = bool(tf32)
    self.torch_default = bool(torch_default)                      # leave torch's own TF32 switches alone (cuDNN convs TF32, matmuls fp32)

  def _proposals(self, score_map, delta_map, anchor_map, image_shape, pre_nms, post_nms):
    """models/rpn.py:99-156 with the tensors on the device (the anchor map is uploaded every step, rpn.py:120)."""
    anchors = t.from_numpy(np.ascontiguousarray(anchor_map.reshape(-1, 4))).to(self.device)
    scores = score_map.detach().reshape(-1)
    deltas = delta_map.detach().reshape(-1, 4)
    c = anchors[:, 2:4] * deltas[:, 0:2] + anchors[:, 0:2]
    s = anchors[:, 2:4] * t.exp(deltas[:, 2:4])
    boxes = t.cat([c - 0.5 * s, c + 0.5 * s], dim = 1)
    order = t.argsort(scores).flip(dims = (0,))
    boxes, scores = boxes[order][0:pre_nms], scores[order][0:pre_nms]
    boxes[:, 0:2] = t.clamp(boxes[:, 0:2], min = 0)
    boxes[:, 2] = t.clamp(boxes[:, 2], max = image_shape[1])
    boxes[:, 3] = t.clamp(boxes[:, 3], max = image_shape[2])
    big = t.where(((boxes[:, 2] - boxes[:, 0]) >= 16) & ((boxes[:, 3] - boxes[:, 1]) >= 16))[0]
    boxes, scores = boxes[big], scores[big]
    keep = _nms(boxes, scores, 0.7)[0:post_nms]                     # (y1,x1,y2,x2) passed as is, like the reference: IoU is symmetric in the axes
    return boxes[keep]

  def _label(self, proposals, gt_corners, gt_class_idxs):
    """models/faster_rcnn.py:418-524."""
    dev = self.device
    gt = t.from_numpy(np.asarray(gt_corners, dtype = np.float32)).to(dev)
    cls = t.tensor(list(gt_class_idxs), dtype = t.long, device = dev)
    props = t.vstack([proposals, gt])
    ious = orc.iou_t(props, gt)
    best = t.max(ious, dim = 1).values
    which = t.argmax(ious, dim = 1)
    cls, gtb = cls[which], gt[which]
    cls[best < 0.5] = 0
    n = props.shape[0]
    onehot = t.zeros((n, self.num_classes), dtype = t.float32, device = dev)
    onehot[t.arange(n, device = dev), cls] = 1.0
    pc, ps = 0.5 * (props[:, 0:2] + props[:, 2:4]), props[:, 2:4] - props[:, 0:2]
    gc, gs = 0.5 * (gtb[:, 0:2] + gtb[:, 2:4]), gtb[:, 2:4] - gtb[:, 0:2]
    tg = t.cat([(gc - pc) / ps, t.log(gs / ps)], dim = 1) / t.tensor([0.1, 0.1, 0.2, 0.2], dtype = t.float32, device = dev)
    packed = t.zeros((n, 2, 4 * (self.num_classes - 1)), dtype = t.float32, device = dev)
    packed[:, 0, :] = t.repeat_interleave(onehot, repeats = 4, dim = 1)[:, 4:]
    packed[:, 1, :] = t.tile(tg, dims = (1, self.num_classes - 1))
    return props, onehot, packed

  def train_step(self, image, anchor_map, anchor_valid_map, gt_rpn_map, obj_idx, bg_idx, gt_corners, gt_class_idxs):
    if self.torch_default:
      t.backends.cudnn.allow_tf32, t.backends.cuda.matmul.allow_tf32 = True, False
    else:
      t.backends.cudnn.allow_tf32 = self.tf32
      t.backends.cuda.matmul.allow_tf32 = self.tf32
    P = self.params
    self.optimizer.zero_grad()
    image_shape = tuple(image.shape[1:])
    fm = orc.vgg16_features(P, image)
    score_map, delta_map = orc.rpn_heads(P, fm)
    proposals = self._proposals(score_map, delta_map, anchor_map, image_shape, 12000, 2000)
    minibatch = orc.sample_rpn_minibatch(gt_rpn_map, obj_idx, bg_idx, 256)
    props, gt_classes, gt_deltas = self._label(proposals, gt_corners, gt_class_idxs)
    props, gt_classes, gt_deltas = orc.sample_proposals(props, gt_classes, gt_deltas, 128, 0.25)      # .where / CPU randperm: host round trips, as in the reference
    props = props.detach()
    rois = t.cat([t.zeros((props.shape[0], 1), device = self.device), props], dim = 1)[:, [0, 2, 1, 4, 3]].contiguous()
    pooled = _roi_pool(fm, rois)
    y = orc.vgg16_pool_to_feature_vector(P, pooled)
    classes = F.softmax(F.linear(y, P[orc.S3 + "_classifier.weight"], P[orc.S3 + "_classifier.bias"]), dim = 1)
    deltas = F.linear(y, P[orc.S3 + "_regressor.weight"], P[orc.S3 + "_regressor.bias"])
    l1 = orc.rpn_class_loss(score_map, minibatch)
    l2 = orc.rpn_regression_loss(delta_map, minibatch)
    l3 = orc.detector_class_loss(classes, gt_classes.detach())
    l4 = orc.detector_regression_loss(deltas, gt_deltas.detach())
    total = l1 + l2 + l3 + l4
    out = (l1.item(), l2.item(), l3.item(), l4.item(), total.item())                                   # faster_rcnn.py:347-353: five .item() syncs
    total.backward()
    self.optimizer.step()
    return out


def time_train_steps(image_hw, steps, warmup, tf32 = False, seed = 0, torch_default = False):
  """-> (list of per-step seconds measured with CUDA events around each step, losses of the last step)."""
  params = orc.synth_params(orc.vgg16_param_shapes(), seed = seed, heads = "reference")
  model = EagerGpuModel(params, "cuda", tf32 = tf32, torch_default = torch_default)
  smp = orc.synthetic_sample(image_hw, seed = seed)
  image, gmap = smp["image"].cuda(), smp["gt_rpn_map"].cuda()
  random.seed(0); np.random.seed(0); t.manual_seed(0)
  times, last = [], None
  for i in range(warmup + steps):
    a, b = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
    a.record()
    last = model.train_step(image, smp["anchor_map"], smp["anchor_valid_map"], gmap, smp["gt_rpn_object_indices"], smp["gt_rpn_background_indices"],
                            smp["gt_corners"], smp["gt_class_idxs"])
    b.record()
    t.cuda.synchronize()
    if i >= warmup:
      times.append(a.elapsed_time(b) / 1e3)
  return times, last
