"""
Backbone protocol (reference: pytorch/FasterRCNN/models/backbone.py:30-65) and the parameter
containers shared by the backbones.  A backbone is a plain object exposing
``feature_map_channels``, ``feature_pixels``, ``feature_vector_size``,
``image_preprocessing_params``, ``feature_extractor`` (nn.Module: image -> feature map),
``pool_to_feature_vector`` (nn.Module: RoIs -> vectors) and ``compute_feature_map_shape``.
"""
import math
from dataclasses import dataclass
from enum import Enum
from typing import List

import torch as t
from torch import nn


class ChannelOrder(Enum):          # datasets/image.py:19-21
  RGB = "RGB"
  BGR = "BGR"


@dataclass
class PreprocessingParams:         # datasets/image.py:24-32
  channel_order: ChannelOrder
  scaling: float
  means: List[float]
  stds: List[float]


class Backbone:
  def __init__(self):
    self.feature_map_channels = 0
    self.feature_pixels = 0
    self.feature_vector_size = 0
    self.image_preprocessing_params = PreprocessingParams(channel_order = ChannelOrder.BGR, scaling = 1.0, means = [103.939, 116.779, 123.680], stds = [1, 1, 1])
    self.feature_extractor = None
    self.pool_to_feature_vector = None

  def compute_feature_map_shape(self, image_shape):
    return image_shape[-3:]


class ConvParams(nn.Module):
  """weight (Cout,Cin,KH,KW) held in channels_last (= OHWI, the kernels' filter layout) + bias.
  Same names, shapes and default initialisation as nn.Conv2d, so state dicts interchange."""

  def __init__(self, in_channels, out_channels, kernel_size, bias = True):
    super().__init__()
    kh, kw = kernel_size
    w = t.empty((out_channels, in_channels, kh, kw), dtype = t.float32)
    nn.init.kaiming_uniform_(w, a = math.sqrt(5))
    self.weight = nn.Parameter(w.contiguous(memory_format = t.channels_last))
    if bias:
      bound = 1.0 / math.sqrt(in_channels * kh * kw)
      self.bias = nn.Parameter(t.empty((out_channels,), dtype = t.float32).uniform_(-bound, bound))
    else:
      self.register_parameter("bias", None)


class LinearParams(nn.Module):
  """weight (out,in) + bias; names, shapes and default initialisation of nn.Linear."""

  def __init__(self, in_features, out_features):
    super().__init__()
    w = t.empty((out_features, in_features), dtype = t.float32)
    nn.init.kaiming_uniform_(w, a = math.sqrt(5))
    self.weight = nn.Parameter(w)
    bound = 1.0 / math.sqrt(in_features)
    self.bias = nn.Parameter(t.empty((out_features,), dtype = t.float32).uniform_(-bound, bound))
