"""
VGG-16 backbone (reference: pytorch/FasterRCNN/models/vgg16.py).  13x (3x3 conv + ReLU) with four
2x2 max pools -> stride-16, 512-channel map; blocks 1-2 frozen (vgg16.py:50-58); RoI head =
fc1(25088->4096)+ReLU+dropout, fc2(4096->4096)+ReLU+dropout (vgg16.py:101-135).  Every conv /
pool / linear runs in the sm_100a kernels (ops.conv2d_act with the pool fused into the layer's
autograd node, ops.linear_act).
"""
import torch as t
from torch import nn

from . import ops
from .backbone import Backbone, ChannelOrder, ConvParams, LinearParams, PreprocessingParams

_LAYERS = [
  ("_block1_conv1", 3, 64, False), ("_block1_conv2", 64, 64, True),
  ("_block2_conv1", 64, 128, False), ("_block2_conv2", 128, 128, True),
  ("_block3_conv1", 128, 256, False), ("_block3_conv2", 256, 256, False), ("_block3_conv3", 256, 256, True),
  ("_block4_conv1", 256, 512, False), ("_block4_conv2", 512, 512, False), ("_block4_conv3", 512, 512, True),
  ("_block5_conv1", 512, 512, False), ("_block5_conv2", 512, 512, False), ("_block5_conv3", 512, 512, False),
]
_FROZEN = ("_block1_conv1", "_block1_conv2", "_block2_conv1", "_block2_conv2")


class FeatureExtractor(nn.Module):
  def __init__(self):
    super().__init__()
    for name, cin, cout, _ in _LAYERS:
      setattr(self, name, ConvParams(cin, cout, (3, 3)))
    for name in _FROZEN:                                   # vgg16.py:50-58
      layer = getattr(self, name)
      layer.weight.requires_grad = False
      layer.bias.requires_grad = False

  def forward(self, image_data):
    """(1,3,H,W) fp32 -> (1,512,H//16,W//16) (logical NCHW, physically NHWC)."""
    y = image_data
    for name, _, _, pool in _LAYERS:
      layer = getattr(self, name)
      y = ops.conv2d_act(y, layer.weight, layer.bias, stride = 1, pad = 1, act = ops.ACT_RELU, pool = pool)
    return y


class PoolToFeatureVector(nn.Module):
  def __init__(self, dropout_probability):
    super().__init__()
    self._fc1 = LinearParams(512 * 7 * 7, 4096)
    self._fc2 = LinearParams(4096, 4096)
    self._dropout1 = nn.Dropout(p = dropout_probability)
    self._dropout2 = nn.Dropout(p = dropout_probability)

  def forward(self, rois):
    """(N,512,7,7) -> (N,4096); the flatten is in (C,7,7) order (vgg16.py:129)."""
    x = rois.reshape((rois.shape[0], 512 * 7 * 7))
    y = ops.linear_act(x, self._fc1.weight, self._fc1.bias, ops.ACT_RELU)
    if self.training and self._dropout1.p > 0:
      y = self._dropout1(y)
    y = ops.linear_act(y, self._fc2.weight, self._fc2.bias, ops.ACT_RELU)
    if self.training and self._dropout2.p > 0:
      y = self._dropout2(y)
    return y


class VGG16Backbone(Backbone):
  def __init__(self, dropout_probability):
    super().__init__()
    self.feature_map_channels = 512
    self.feature_pixels = 16
    self.feature_vector_size = 4096
    self.image_preprocessing_params = PreprocessingParams(channel_order = ChannelOrder.BGR, scaling = 1.0, means = [103.939, 116.779, 123.680], stds = [1, 1, 1])
    self.feature_extractor = FeatureExtractor()
    self.pool_to_feature_vector = PoolToFeatureVector(dropout_probability = dropout_probability)

  def compute_feature_map_shape(self, image_shape):
    return (self.feature_map_channels, image_shape[-2] // self.feature_pixels, image_shape[-1] // self.feature_pixels)
