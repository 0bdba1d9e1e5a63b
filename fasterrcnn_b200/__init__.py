"""
fasterrcnn_b200 -- B200-native Faster R-CNN hot path (hand-written sm_100a CUDA behind a C ABI)
with the model / training-loop API of trzy/FasterRCNN's pytorch/FasterRCNN package.

    from fasterrcnn_b200 import FasterRCNNModel, vgg16, resnet
    model = FasterRCNNModel(num_classes = 21, backbone = vgg16.VGG16Backbone(dropout_probability = 0.0)).cuda()
"""
from . import _lib, ops                                             # noqa: F401
from . import anchors, math_utils, backbone, vgg16, vgg16_torch, resnet, rpn, detector, optim   # noqa: F401
from .faster_rcnn import FasterRCNNModel                            # noqa: F401

__all__ = ["FasterRCNNModel", "vgg16", "vgg16_torch", "resnet", "rpn", "detector", "anchors", "math_utils", "backbone", "ops", "optim"]
