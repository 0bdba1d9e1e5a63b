"""Import aliases with the reference's package layout (`from .models import vgg16, vgg16_torch, resnet`,
`from .models.faster_rcnn import FasterRCNNModel`; pytorch/FasterRCNN/__main__.py:25-29)."""
import sys

from .. import anchors, backbone, detector, faster_rcnn, math_utils, resnet, rpn, vgg16, vgg16_torch   # noqa: F401

for _name in ("anchors", "backbone", "detector", "faster_rcnn", "math_utils", "resnet", "rpn", "vgg16", "vgg16_torch"):
  sys.modules[__name__ + "." + _name] = globals()[_name]
