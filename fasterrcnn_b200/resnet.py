"""
ResNet-50/101/152 backbones (reference: pytorch/FasterRCNN/models/resnet.py, which wraps torchvision's
v1.5 ResNets): feature extractor = conv1, bn1, relu, maxpool, layer1..3 (1024 channels, ceil(H/16)),
RoI head = layer4 on the (N,1024,7,7) pooled RoIs followed by a spatial mean -> (N,2048).
All BatchNorm layers are frozen and evaluated with running statistics (resnet.py:56-77,100-107), so
each conv+BN pair runs as ONE implicit-GEMM launch: the per-channel factor gamma/sqrt(var+eps) is
folded into the filter rows, the shift beta - mean*factor is the epilogue bias, the residual add
and ReLU are the epilogue too.  conv1 / bn1 / layer1 and every BN affine parameter have
requires_grad=False (resnet.py:48-55,96).  Parameter and buffer names equal torchvision's, so the
reference's state dicts load unchanged.
"""
import os
from enum import Enum
from math import ceil

import torch as t
from torch import nn

from . import _lib, ops
from ._lib import check, lib, ptr, stream
from .backbone import Backbone, ChannelOrder, ConvParams, PreprocessingParams

_BLOCKS = {"ResNet50": (3, 4, 6, 3), "ResNet101": (3, 4, 23, 3), "ResNet152": (3, 8, 36, 3)}


class Architecture(Enum):
  ResNet50 = "ResNet50"
  ResNet101 = "ResNet101"
  ResNet152 = "ResNet152"


class BNParams(nn.Module):
  """Frozen BatchNorm2d parameter container (names of nn.BatchNorm2d); never updates its statistics."""

  def __init__(self, channels, eps = 1e-5):
    super().__init__()
    self.eps = eps
    self.weight = nn.Parameter(t.ones(channels), requires_grad = False)
    self.bias = nn.Parameter(t.zeros(channels), requires_grad = False)
    self.register_buffer("running_mean", t.zeros(channels))
    self.register_buffer("running_var", t.ones(channels))
    self.register_buffer("num_batches_tracked", t.tensor(0, dtype = t.long))
    self._folded = None

  def folded(self):
    """(scale, shift) with scale = gamma / sqrt(var + eps), shift = beta - mean * scale; cached until the tensors change."""
    key = (self.weight._version, self.bias._version, self.running_mean._version, self.running_var._version, self.weight.device)
    if self._folded is None or self._folded[0] != key:
      with t.no_grad():
        scale = self.weight / t.sqrt(self.running_var + self.eps)
        shift = self.bias - self.running_mean * scale
      self._folded = (key, scale.contiguous(), shift.contiguous())
    return self._folded[1], self._folded[2]


def _subsample2(x):
  """x[:, :, ::2, ::2] of an NHWC tensor -- the pixels a stride-2 convolution keeps (frcnn_subsample2)."""
  n, c, h, w = x.shape
  y = t.empty((n, c, (h + 1) // 2, (w + 1) // 2), dtype = t.float32, device = x.device, memory_format = t.channels_last)
  if y.numel() > 0:
    check(lib().frcnn_subsample2(ptr(x), ptr(y), n, h, w, c, stream()), "frcnn_subsample2")
    _lib.count()
  ops._copy_amax_hint(x, y)                                    # a subset of x's values
  return y


def _upsample2_zero(x, h, w):
  """The adjoint: an (n, c, h, w) NHWC tensor with x at the even pixels and zeros elsewhere (frcnn_upsample2_zero)."""
  n, c = x.shape[0], x.shape[1]
  assert x.shape[2] == (h + 1) // 2 and x.shape[3] == (w + 1) // 2
  y = t.empty((n, c, h, w), dtype = t.float32, device = x.device, memory_format = t.channels_last)
  if y.numel() > 0:
    check(lib().frcnn_upsample2_zero(ptr(x), ptr(y), n, h, w, c, stream()), "frcnn_upsample2_zero")
    _lib.count()
  ops._copy_amax_hint(x, y)
  return y


def _strided_on_tensor_cores(x, w, stride, pad):
  """The tcgen05 engine is stride-1.  The two stride-2 forms of a torchvision Bottleneck (resnet.py:79-81,109-118) are run on it as
    1x1, stride 2         conv1x1_s1(x[::2, ::2])          -- same products, same count
    3x3, stride 2, pad 1  conv3x3_s1(x)[::2, ::2]          -- 4x the products of three layers, still ~5x faster than the CUDA-core
                                                              kernel they ran on (profiles/r02_resnet_launches.md)
  and backwards as the stride-1 dgrad / wgrad of the zero-upsampled gradient (zeros add exact zeros: the sums are the strided form's).
  FRCNN_RESNET_S2_TC=0 restores the CUDA-core strided kernels."""
  if stride != 2 or os.environ.get("FRCNN_RESNET_S2_TC", "1") in ("", "0"):
    return None
  k = w.shape[2]
  if k == 1 and pad == 0:
    return "1x1"
  if k == 3 and pad == 1 and w.shape[3] == 3:
    return "3x3"
  return None


def _act_bwd_scale(dy, y, scale, want_dz):
  """(dz, dzs) of frcnn_act_bwd_scale: dz = dy * (y > 0) (None unless want_dz), dzs = dz * scale[c]."""
  n, c, h, w = dy.shape
  dzs = t.empty_like(dy)
  dz = t.empty_like(dy) if want_dz else None
  if dy.numel() > 0:
    check(lib().frcnn_act_bwd_scale(ptr(dy), ptr(y), ptr(scale), ptr(dz), ptr(dzs), n * h * w, c, stream()), "frcnn_act_bwd_scale")
    _lib.count()
  return dz, dzs


class _ConvBNAct(t.autograd.Function):
  """y = act(conv(x, w) * s + shift [+ residual]): the frozen BatchNorm is the GEMM's per-channel epilogue (factor s, offset shift), the
  residual add and the ReLU too.  The filter the GEMM reads is the parameter itself, so the operand split the fused optimizer carries
  with every weight (ops.weight_split_buffer) serves it, frozen filters keep theirs (ops.pin_split), and no scaled copy of the weights
  is made per step.  Backward: one pass gives dz = dy * (y > 0) (the residual branch's gradient) and dz * s (the convolution's)."""

  @staticmethod
  def forward(ctx, x, w, scale, shift, residual, stride, pad, act):
    xp = ops.as_nhwc(x.detach())
    wp = ops._phys_filter(w.detach())
    if not w.requires_grad and wp.data_ptr() == w.data_ptr():
      ops.pin_split(w)                                         # frozen filter (conv1, layer1): split once, not once per step
    res = ops.as_nhwc(residual.detach()) if residual is not None else None
    want_dw = bool(ctx.needs_input_grad[1])
    form = _strided_on_tensor_cores(xp, w, stride, pad)
    ctx.x_shape = tuple(xp.shape)
    if form == "1x1":
      xp = _subsample2(xp)                                     # saved instead of x: the filter gradient reads the same pixels
      y = ops.conv2d_fwd_raw(xp, wp, shift, 1, 0, act, scale = scale, residual = res, reuse_x = want_dw)
    elif form == "3x3":
      y = _subsample2(ops.conv2d_fwd_raw(xp, wp, shift, 1, 1, act, scale = scale, residual = res, reuse_x = want_dw))
    else:
      y = ops.conv2d_fwd_raw(xp, wp, shift, stride, pad, act, scale = scale, residual = res, reuse_x = want_dw)
    ctx.stride, ctx.pad, ctx.act, ctx.form = stride, pad, act, form
    ctx.w_shape = tuple(w.shape)
    ctx.w_key = id(w)
    ctx.has_residual = residual is not None
    ctx.save_for_backward(xp, wp, scale, y)
    return y

  @staticmethod
  def backward(ctx, dy):
    xp, wp, scale, y = ctx.saved_tensors
    dy = ops.as_nhwc(dy)
    want_dx, want_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    want_res = ctx.has_residual and ctx.needs_input_grad[4]
    relu = ctx.act == ops.ACT_RELU
    dz, dzs = _act_bwd_scale(dy, y if relu else None, scale, want_res and relu) if (want_dx or want_dw) else (None, None)
    if want_res and dz is None:
      if relu:                                                  # (only the residual branch needs a gradient)
        dz = t.empty_like(y)
        check(lib().frcnn_relu_bwd(ptr(dy), ptr(y), ptr(dz), y.numel(), stream()), "frcnn_relu_bwd")
        _lib.count()
      else:
        dz = dy
    dx = dw = None
    if ctx.form == "1x1":
      if want_dx:
        dx = _upsample2_zero(ops.conv2d_dgrad_raw(dzs, wp, tuple(xp.shape), 1, 0, reuse_dy = want_dw), ctx.x_shape[2], ctx.x_shape[3])
      if want_dw:
        dw = ops.conv2d_wgrad_raw(dzs, xp, ctx.w_shape, 1, 0, w_key = ctx.w_key)
    elif ctx.form == "3x3":
      dz_full = _upsample2_zero(dzs, ctx.x_shape[2], ctx.x_shape[3]) if (want_dx or want_dw) else None
      if want_dx:
        dx = ops.conv2d_dgrad_raw(dz_full, wp, ctx.x_shape, 1, 1, reuse_dy = want_dw)
      if want_dw:
        dw = ops.conv2d_wgrad_raw(dz_full, xp, ctx.w_shape, 1, 1, w_key = ctx.w_key)
    else:
      if want_dx:
        dx = ops.conv2d_dgrad_raw(dzs, wp, tuple(xp.shape), ctx.stride, ctx.pad, reuse_dy = want_dw)
      if want_dw:
        dw = ops.conv2d_wgrad_raw(dzs, xp, ctx.w_shape, ctx.stride, ctx.pad, w_key = ctx.w_key)
    return dx, dw, None, None, (dz if want_res else None), None, None, None


def conv_bn_act(x, conv, bn, stride, pad, act, residual = None):
  scale, shift = bn.folded()
  return _ConvBNAct.apply(x, conv.weight, scale, shift, residual, stride, pad, act)


class _SpatialMean(t.autograd.Function):
  @staticmethod
  def forward(ctx, x):
    xp = ops.as_nhwc(x.detach())
    n, c, h, w = xp.shape
    y = t.empty((n, c), dtype = t.float32, device = xp.device)
    if n > 0:
      check(lib().frcnn_spatial_mean_fwd(ptr(xp), ptr(y), n, h, w, c, stream()), "frcnn_spatial_mean_fwd")
      _lib.count()
    ctx.shape = (n, c, h, w)
    return y

  @staticmethod
  def backward(ctx, dy):
    n, c, h, w = ctx.shape
    dx = t.empty((n, c, h, w), dtype = t.float32, device = dy.device, memory_format = t.channels_last)
    if n > 0:
      check(lib().frcnn_spatial_mean_bwd(ptr(dy.contiguous()), ptr(dx), n, h * w, c, stream()), "frcnn_spatial_mean_bwd")
      _lib.count()
    return dx


def max_pool_3x3_s2(x):
  """nn.MaxPool2d(3, 2, 1) of the stem (resnet.py:42); the stem is frozen, so no gradient is defined."""
  assert not x.requires_grad, "the ResNet stem is frozen in the reference (resnet.py:48-50)"
  xp = ops.as_nhwc(x)
  n, c, h, w = xp.shape
  ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
  y = t.empty((n, c, ho, wo), dtype = t.float32, device = xp.device, memory_format = t.channels_last)
  check(lib().frcnn_maxpool3x3s2_fwd(ptr(xp), ptr(y), n, h, w, c, stream()), "frcnn_maxpool3x3s2_fwd")
  _lib.count()
  return y


def _make_conv(cin, cout, k):
  m = ConvParams(cin, cout, (k, k), bias = False)
  nn.init.kaiming_normal_(m.weight, mode = "fan_out", nonlinearity = "relu")          # torchvision's ResNet init
  return m


class Bottleneck(nn.Module):
  expansion = 4

  def __init__(self, inplanes, planes, stride, downsample):
    super().__init__()
    self.conv1, self.bn1 = _make_conv(inplanes, planes, 1), BNParams(planes)
    self.conv2, self.bn2 = _make_conv(planes, planes, 3), BNParams(planes)
    self.conv3, self.bn3 = _make_conv(planes, planes * 4, 1), BNParams(planes * 4)
    self.stride = stride
    if downsample:
      self.downsample = nn.Sequential(_make_conv(inplanes, planes * 4, 1), BNParams(planes * 4))
    else:
      self.downsample = None

  def forward(self, x):
    y = conv_bn_act(x, self.conv1, self.bn1, 1, 0, ops.ACT_RELU)
    y = conv_bn_act(y, self.conv2, self.bn2, self.stride, 1, ops.ACT_RELU)
    identity = x
    if self.downsample is not None:
      identity = conv_bn_act(x, self.downsample[0], self.downsample[1], self.stride, 0, ops.ACT_NONE)
    return conv_bn_act(y, self.conv3, self.bn3, 1, 0, ops.ACT_RELU, residual = identity)      # relu(bn3(conv3) + identity)


def _make_layer(inplanes, planes, blocks, stride):
  layers = [Bottleneck(inplanes, planes, stride, True)]
  for _ in range(1, blocks):
    layers.append(Bottleneck(planes * 4, planes, 1, False))
  return nn.Sequential(*layers)


def _freeze(module):
  for p in module.parameters():
    p.requires_grad = False


class FeatureExtractor(nn.Module):
  def __init__(self, blocks):
    super().__init__()
    self._feature_extractor = nn.Sequential(
      _make_conv(3, 64, 7),                       # 0  conv1
      BNParams(64),                               # 1  bn1
      nn.Identity(),                              # 2  relu    (fused into the conv epilogue)
      nn.Identity(),                              # 3  maxpool (ops kernel)
      _make_layer(64, 64, blocks[0], 1),          # 4  layer1
      _make_layer(256, 128, blocks[1], 2),        # 5  layer2
      _make_layer(512, 256, blocks[2], 2),        # 6  layer3
    )
    _freeze(self._feature_extractor[0])
    _freeze(self._feature_extractor[1])
    _freeze(self._feature_extractor[4])

  def forward(self, image_data):
    fe = self._feature_extractor
    y = conv_bn_act(image_data, fe[0], fe[1], 2, 3, ops.ACT_RELU)
    y = max_pool_3x3_s2(y)
    for idx in (4, 5, 6):
      for block in fe[idx]:
        y = block(y)
    return y


class PoolToFeatureVector(nn.Module):
  def __init__(self, blocks):
    super().__init__()
    self._layer4 = _make_layer(1024, 512, blocks[3], 2)

  def forward(self, rois):
    y = rois
    for block in self._layer4:
      y = block(y)
    return _SpatialMean.apply(y)                   # mean over W, then over H (resnet.py:117)


class ResNetBackbone(Backbone):
  def __init__(self, architecture):
    super().__init__()
    if not isinstance(architecture, Architecture) or architecture.value not in _BLOCKS:
      raise ValueError("Invalid ResNet architecture value: %s" % getattr(architecture, "value", architecture))
    self.feature_map_channels = 1024
    self.feature_pixels = 16
    self.feature_vector_size = 2048
    self.image_preprocessing_params = PreprocessingParams(channel_order = ChannelOrder.RGB, scaling = 1.0 / 255.0, means = [0.485, 0.456, 0.406], stds = [0.229, 0.224, 0.225])
    blocks = _BLOCKS[architecture.value]
    # The reference pre-loads IMAGENET1K_V1 weights through torchvision (resnet.py:145-149); they are
    # loaded here from a state dict (load_state_dict / --load-from), the default is torchvision's random init.
    self.feature_extractor = FeatureExtractor(blocks)
    self.pool_to_feature_vector = PoolToFeatureVector(blocks)

  def compute_feature_map_shape(self, image_shape):
    return (self.feature_map_channels, ceil(image_shape[-2] / self.feature_pixels), ceil(image_shape[-1] / self.feature_pixels))
