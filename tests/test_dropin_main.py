"""
Static proof of the drop-in claim: every name the reference's entry point imports from the packages this repository
mirrors exists here, and every call it makes into them -- constructor, method, keyword by keyword -- binds to this package's signatures.
The reference's pytorch/FasterRCNN/__main__.py is parsed (never executed: it needs CUDA, a dataset and imageio), its call sites are
collected from the AST and bound with inspect.signature.  Runs where /root/reference exists (this container); the GPU box has no copy.
"""
import ast
import inspect
import os

import pytest

REF_MAIN = "/root/reference/pytorch/FasterRCNN/__main__.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_MAIN), reason = "the reference tree is not present on this machine")

# modules of the reference that are host-side UI / logging, out of the hot path's scope: imported by __main__, not mirrored
OUT_OF_SCOPE = {"utils", "visualize", "profile"}


def _tree():
  with open(REF_MAIN) as f:
    return ast.parse(f.read())


def _ours(module):
  import importlib
  return importlib.import_module("fasterrcnn_b200" + ("." + module if module else ""))


def test_every_import_of_the_reference_entry_point_resolves_here():
  found = []
  for node in ast.walk(_tree()):
    if isinstance(node, ast.ImportFrom) and node.level == 1:
      module = node.module or ""
      for alias in node.names:
        top = (module.split(".")[0] if module else alias.name)
        if top in OUT_OF_SCOPE:
          continue
        mod = _ours(module)
        assert hasattr(mod, alias.name) or _ours((module + "." if module else "") + alias.name), (module, alias.name)
        found.append((module, alias.name))
  # pytorch/FasterRCNN/__main__.py:26-33,238
  for want in (("datasets", "voc"), ("models.faster_rcnn", "FasterRCNNModel"), ("models", "vgg16"), ("models", "vgg16_torch"), ("models", "resnet"),
               ("statistics", "TrainingStatistics"), ("statistics", "PrecisionRecallCurveCalculator"), ("", "state"), ("datasets", "image")):
    assert want in found, want


def _call_sites():
  """dotted callee text -> list of (positional count, keyword names, line)."""
  sites = {}
  for node in ast.walk(_tree()):
    if isinstance(node, ast.Call):
      sites.setdefault(ast.unparse(node.func), []).append((len(node.args), [k.arg for k in node.keywords if k.arg is not None], node.lineno))
  return sites


def test_every_call_into_the_mirrored_packages_binds_to_our_signatures():
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import resnet, state, statistics, vgg16, vgg16_torch
  from fasterrcnn_b200.datasets import image, voc
  model_cls = f.FasterRCNNModel
  targets = {
    # callee text in the reference            our callable                                   bound method? (skip `self`)
    "voc.Dataset": (voc.Dataset, False),
    "FasterRCNNModel": (model_cls, False),
    "model.train_step": (model_cls.train_step, True),
    "model.predict": (model_cls.predict, True),
    "vgg16.VGG16Backbone": (vgg16.VGG16Backbone, False),
    "vgg16_torch.VGG16Backbone": (vgg16_torch.VGG16Backbone, False),
    "resnet.ResNetBackbone": (resnet.ResNetBackbone, False),
    "state.load": (state.load, False),
    "state.BestWeightsTracker": (state.BestWeightsTracker, False),
    "best_weights_tracker.on_epoch_end": (state.BestWeightsTracker.on_epoch_end, True),
    "best_weights_tracker.save_best_weights": (state.BestWeightsTracker.save_best_weights, True),
    "TrainingStatistics": (statistics.TrainingStatistics, False),
    "stats.on_training_step": (statistics.TrainingStatistics.on_training_step, True),
    "stats.get_progbar_postfix": (statistics.TrainingStatistics.get_progbar_postfix, True),
    "PrecisionRecallCurveCalculator": (statistics.PrecisionRecallCurveCalculator, False),
    "precision_recall_curve.add_image_results": (statistics.PrecisionRecallCurveCalculator.add_image_results, True),
    "precision_recall_curve.compute_mean_average_precision": (statistics.PrecisionRecallCurveCalculator.compute_mean_average_precision, True),
    "precision_recall_curve.print_average_precisions": (statistics.PrecisionRecallCurveCalculator.print_average_precisions, True),
    "image.load_image": (image.load_image, False),
  }
  sites = _call_sites()
  checked = 0
  for text, (fn, bound) in targets.items():
    assert text in sites, "the reference's __main__ no longer calls %s" % text
    sig = inspect.signature(fn)
    for npos, kwargs, line in sites[text]:
      args = ([object()] if bound else []) + [object()] * npos
      try:
        sig.bind(*args, **{k: object() for k in kwargs})
      except TypeError as e:
        raise AssertionError("__main__.py:%d  %s(%s): does not bind to %s%s: %s" % (line, text, ", ".join(kwargs), fn.__qualname__, sig, e))
      checked += 1
  assert checked >= 25
  # attributes the entry point reads off the objects it builds (__main__.py:40-42,70,173,100-104,201,222)
  backbone = vgg16.VGG16Backbone(dropout_probability = 0.0)
  for attr in ("image_preprocessing_params", "compute_feature_map_shape", "feature_pixels"):
    assert hasattr(backbone, attr), attr
  assert hasattr(voc.Dataset, "class_index_to_name") and hasattr(voc.Dataset, "num_classes")
  model = model_cls(num_classes = voc.Dataset.num_classes, backbone = backbone, allow_edge_proposals = True)
  assert model.backbone is backbone
  keys = [k for k, v in dict(model.named_parameters()).items() if v.requires_grad and "weight" in k]      # create_optimizer's filter (__main__.py:98-105)
  assert len(keys) == 16
  assert list(model.state_dict().keys())[0] == "_stage1_feature_extractor._block1_conv1.weight" and len(model.state_dict()) == 40
  assert set(inspect.signature(model_cls.Loss).parameters) == {"rpn_class", "rpn_regression", "detector_class", "detector_regression", "total"}
