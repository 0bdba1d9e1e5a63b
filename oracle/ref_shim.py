"""
TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (trzy/FasterRCNN).

Imports ``pytorch.FasterRCNN`` from ``/root/reference`` on CPU so that
``oracle/make_golden.py`` can execute the reference itself and dump golden vectors,
and so that ``oracle/check_vs_live_reference.py`` (run by tests/test_oracle.py in a subprocess) can pin
the restatement in ``oracle/frcnn_oracle.py`` against it on fresh cases.  ``/root/reference`` only exists in the build
container: nothing that runs on the GPU box may import this module.

The patches below are ENVIRONMENT patches, not algorithm changes:
  1. ``imageio`` is not installed -> stub module (pytorch/FasterRCNN/datasets/image.py:11).
  2. hard-coded ``.cuda()`` / ``device="cuda"`` placement -> CPU
     (rpn.py:120-122, detector.py:65, faster_rcnn.py:217-218,460-461,494,508,511-512,520,
      math_utils.py:125).
  3. torchvision's CPU ``nms`` insists on boxes.dtype == scores.dtype while the CUDA op the
     reference was written against accepts f64 boxes + f32 scores (faster_rcnn.py:216-220):
     the scores are up-cast, the f64 boxes are kept.
  4. ``t.argsort`` at rpn.py:129 is made stable so that ties have a defined order
     (stable ascending, then flip => higher anchor index first among equal scores).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FRCNN_REFERENCE_ROOT", "/root/reference")


def available():
  return os.path.isdir(os.path.join(REFERENCE_ROOT, "pytorch", "FasterRCNN"))


_loaded = None


def load():
  """Returns a namespace with the reference's modules (vgg16, resnet, anchors, faster_rcnn, ...)."""
  global _loaded
  if _loaded is not None:
    return _loaded
  if not available():
    raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
  import torch as t
  import torchvision

  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  sys.modules.setdefault("imageio", types.ModuleType("imageio"))

  # (2) CUDA placement -> identity / cpu
  t.Tensor.cuda = lambda self, *a, **k: self
  t.nn.Module.cuda = lambda self, *a, **k: self
  for name in ("tensor", "zeros", "empty"):
    orig = getattr(t, name)
    if getattr(orig, "_frcnn_shimmed", False):
      continue

    def make(o):
      def f(*a, **k):
        if k.get("device") == "cuda":
          k = dict(k)
          k["device"] = "cpu"
        return o(*a, **k)
      f._frcnn_shimmed = True
      return f
    setattr(t, name, make(orig))

  from pytorch.FasterRCNN.models import vgg16, resnet, anchors, faster_rcnn, rpn, detector, math_utils
  from pytorch.FasterRCNN.datasets.training_sample import Box

  # (3) dtype cast for the final per-class NMS
  _nms = torchvision.ops.nms
  faster_rcnn.nms = lambda boxes, scores, iou_threshold: _nms(boxes, scores.to(boxes.dtype), iou_threshold)

  # (4) stable argsort inside the RPN module only
  class _StableArgsortTorch:
    def __getattr__(self, k):
      return getattr(t, k)

    @staticmethod
    def argsort(x, *a, **k):
      k.setdefault("stable", True)
      return t.argsort(x, *a, **k)
  rpn.t = _StableArgsortTorch()

  ns = types.SimpleNamespace(
    vgg16 = vgg16, resnet = resnet, anchors = anchors, faster_rcnn = faster_rcnn, rpn = rpn,
    detector = detector, math_utils = math_utils, Box = Box, torchvision = torchvision
  )
  _loaded = ns
  return ns


def patch_resnet_offline(zero_init_residual = True):
  """ResNet backbones download IMAGENET weights (resnet.py:145-149): impossible offline."""
  import torchvision
  for name in ("resnet50", "resnet101", "resnet152"):
    orig = getattr(torchvision.models, name)
    if getattr(orig, "_frcnn_shimmed", False):
      continue

    def make(o):
      def f(weights = None, **k):
        return o(weights = None, zero_init_residual = zero_init_residual, **k)
      f._frcnn_shimmed = True
      return f
    setattr(torchvision.models, name, make(orig))
