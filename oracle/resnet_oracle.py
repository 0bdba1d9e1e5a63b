"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's ResNet backbones
(pytorch/FasterRCNN/models/resnet.py): torchvision ResNet-50/101/152 (v1.5: the stride sits on
the 3x3 conv of each bottleneck), feature extractor = conv1, bn1, relu, maxpool, layer1..3
(resnet.py:38-46), RoI head = layer4 + spatial mean (resnet.py:94-118); every BatchNorm is frozen and
runs in eval mode (resnet.py:56-77,100-107); conv1/bn1/layer1 and all BN affine parameters have
requires_grad=False (resnet.py:48-55).  Parameters are addressed by the reference's state-dict keys.
"""
import torch as t
import torch.nn.functional as F

BLOCKS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
S1 = "_stage1_feature_extractor._feature_extractor."
S3 = "_stage3_detector_network._pool_to_feature_vector._layer4."


def layer_prefixes(arch):
  b = BLOCKS[arch]
  return [(S1 + "4.", 64, 64, b[0], 1), (S1 + "5.", 256, 128, b[1], 2), (S1 + "6.", 512, 256, b[2], 2), (S3, 1024, 512, b[3], 2)]


def _bn_shapes(prefix, c):
  return {prefix + "weight": (c,), prefix + "bias": (c,), prefix + "running_mean": (c,), prefix + "running_var": (c,), prefix + "num_batches_tracked": ()}


def param_shapes(arch, num_classes = 21):
  """Ordered {key: shape} exactly as FasterRCNNModel(backbone = ResNetBackbone(arch)).state_dict() lists them."""
  shapes = {}
  shapes[S1 + "0.weight"] = (64, 3, 7, 7)
  shapes.update(_bn_shapes(S1 + "1.", 64))

  def add_layer(prefix, cin, width, blocks):
    for i in range(blocks):
      p = "%s%d." % (prefix, i)
      inp = cin if i == 0 else width * 4
      shapes[p + "conv1.weight"] = (width, inp, 1, 1); shapes.update(_bn_shapes(p + "bn1.", width))
      shapes[p + "conv2.weight"] = (width, width, 3, 3); shapes.update(_bn_shapes(p + "bn2.", width))
      shapes[p + "conv3.weight"] = (width * 4, width, 1, 1); shapes.update(_bn_shapes(p + "bn3.", width * 4))
      if i == 0:
        shapes[p + "downsample.0.weight"] = (width * 4, inp, 1, 1); shapes.update(_bn_shapes(p + "downsample.1.", width * 4))
  layers = layer_prefixes(arch)
  for prefix, cin, width, blocks, _ in layers[:3]:
    add_layer(prefix, cin, width, blocks)
  s2 = "_stage2_region_proposal_network."
  shapes[s2 + "_rpn_conv1.weight"] = (1024, 1024, 3, 3); shapes[s2 + "_rpn_conv1.bias"] = (1024,)
  shapes[s2 + "_rpn_class.weight"] = (9, 1024, 1, 1); shapes[s2 + "_rpn_class.bias"] = (9,)
  shapes[s2 + "_rpn_boxes.weight"] = (36, 1024, 1, 1); shapes[s2 + "_rpn_boxes.bias"] = (36,)
  prefix, cin, width, blocks, _ = layers[3]
  add_layer(prefix, cin, width, blocks)
  s3 = "_stage3_detector_network."
  shapes[s3 + "_classifier.weight"] = (num_classes, 2048); shapes[s3 + "_classifier.bias"] = (num_classes,)
  shapes[s3 + "_regressor.weight"] = ((num_classes - 1) * 4, 2048); shapes[s3 + "_regressor.bias"] = ((num_classes - 1) * 4,)
  return shapes


def trainable_keys(params):
  """requires_grad=True: conv weights of layer2, layer3, layer4 + RPN + heads (resnet.py:48-55,96)."""
  out = []
  for k in params:
    if ".bn" in k or "downsample.1." in k or k.startswith(S1 + "0.") or k.startswith(S1 + "1.") or k.startswith(S1 + "4."):
      continue
    out.append(k)
  return out


def _bn(params, prefix, x):
  return F.batch_norm(x, params[prefix + "running_mean"], params[prefix + "running_var"], params[prefix + "weight"], params[prefix + "bias"], False, 0.1, 1e-5)


def _bottleneck(params, p, x, stride, has_down):
  y = F.relu(_bn(params, p + "bn1.", F.conv2d(x, params[p + "conv1.weight"])))
  y = F.relu(_bn(params, p + "bn2.", F.conv2d(y, params[p + "conv2.weight"], stride = stride, padding = 1)))
  y = _bn(params, p + "bn3.", F.conv2d(y, params[p + "conv3.weight"]))
  idt = x
  if has_down:
    idt = _bn(params, p + "downsample.1.", F.conv2d(x, params[p + "downsample.0.weight"], stride = stride))
  return F.relu(y + idt)


def _layer(params, prefix, x, blocks, stride):
  for i in range(blocks):
    x = _bottleneck(params, "%s%d." % (prefix, i), x, stride if i == 0 else 1, i == 0)
  return x


def features(params, image, arch):
  """resnet.py:38-46,79-81."""
  y = F.conv2d(image, params[S1 + "0.weight"], stride = 2, padding = 3)
  y = F.relu(_bn(params, S1 + "1.", y))
  y = F.max_pool2d(y, kernel_size = 3, stride = 2, padding = 1)
  for prefix, _, _, blocks, stride in layer_prefixes(arch)[:3]:
    y = _layer(params, prefix, y, blocks, stride)
  return y


def pool_to_feature_vector(params, rois, arch):
  """resnet.py:109-118: layer4 then mean over W, then over H."""
  prefix, _, _, blocks, stride = layer_prefixes(arch)[3]
  y = _layer(params, prefix, rois, blocks, stride)
  return y.mean(-1).mean(-1)
