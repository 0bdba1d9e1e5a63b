"""
Torchvision-layout VGG-16 backbone (reference: pytorch/FasterRCNN/models/vgg16_torch.py): the same
network as vgg16.py, addressed by torchvision's sequential indices (`_layers.0` ... `_layers.28` for the
convs, `_layers.0` / `_layers.3` for the two classifier linears) and fed RGB/255 mean-std images
(vgg16_torch.py:63).  First four convs frozen (vgg16_torch.py:29-34).  Same kernels as vgg16.py.
"""
from torch import nn

from . import ops
from .backbone import Backbone, ChannelOrder, ConvParams, LinearParams, PreprocessingParams

# torchvision.models.vgg16().features[0:30]: (index, cin, cout) for convs, "M" = MaxPool2d(2,2); ReLUs in between
_CFG = [(0, 3, 64), (2, 64, 64), "M", (5, 64, 128), (7, 128, 128), "M", (10, 128, 256), (12, 256, 256), (14, 256, 256), "M",
        (17, 256, 512), (19, 512, 512), (21, 512, 512), "M", (24, 512, 512), (26, 512, 512), (28, 512, 512)]


class FeatureExtractor(nn.Module):
  def __init__(self):
    super().__init__()
    slots = [nn.Identity() for _ in range(30)]
    self._plan = []
    frozen = 0
    for i, item in enumerate(_CFG):
      if item == "M":
        continue
      idx, cin, cout = item
      layer = ConvParams(cin, cout, (3, 3))
      if frozen < 4:                                          # vgg16_torch.py:29-34
        layer.weight.requires_grad = False
        layer.bias.requires_grad = False
        frozen += 1
      slots[idx] = layer
      pool = i + 1 < len(_CFG) and _CFG[i + 1] == "M"
      self._plan.append((idx, pool))
    self._layers = nn.Sequential(*slots)

  def forward(self, image_data):
    y = image_data
    for idx, pool in self._plan:
      layer = self._layers[idx]
      y = ops.conv2d_act(y, layer.weight, layer.bias, 1, 1, ops.ACT_RELU, pool = pool)
    return y


class PoolToFeatureVector(nn.Module):
  def __init__(self, dropout_probability):
    super().__init__()
    # classifier[0:6] = Linear, ReLU, Dropout, Linear, ReLU, Dropout
    self._layers = nn.Sequential(LinearParams(512 * 7 * 7, 4096), nn.Identity(), nn.Dropout(p = dropout_probability),
                                 LinearParams(4096, 4096), nn.Identity(), nn.Dropout(p = dropout_probability))

  def forward(self, rois):
    x = rois.reshape((rois.shape[0], 512 * 7 * 7))
    y = ops.linear_act(x, self._layers[0].weight, self._layers[0].bias, ops.ACT_RELU)
    if self.training and self._layers[2].p > 0:
      y = self._layers[2](y)
    y = ops.linear_act(y, self._layers[3].weight, self._layers[3].bias, ops.ACT_RELU)
    if self.training and self._layers[5].p > 0:
      y = self._layers[5](y)
    return y


class VGG16Backbone(Backbone):
  def __init__(self, dropout_probability):
    super().__init__()
    self.feature_map_channels = 512
    self.feature_pixels = 16
    self.feature_vector_size = 4096
    self.image_preprocessing_params = PreprocessingParams(channel_order = ChannelOrder.RGB, scaling = 1.0 / 255.0, means = [0.485, 0.456, 0.406], stds = [0.229, 0.224, 0.225])
    # The reference pre-loads IMAGENET1K_V1 through torchvision (vgg16_torch.py:67); here weights come from a state dict.
    self.feature_extractor = FeatureExtractor()
    self.pool_to_feature_vector = PoolToFeatureVector(dropout_probability)

  def compute_feature_map_shape(self, image_shape):
    return (self.feature_map_channels, image_shape[-2] // self.feature_pixels, image_shape[-1] // self.feature_pixels)
