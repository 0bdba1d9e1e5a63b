"""
Weights I/O for the B200 model: mirror of the reference's pytorch/FasterRCNN/state.py (load :221-272,
_load_vgg16_from_caffe_model :178-219, _load_vgg16_from_bart_keras_model :117-176, BestWeightsTracker :274-288).

The three accepted file formats and their probing order (Keras h5 -> Caffe VGG-16 .pth -> own checkpoint with a
"model_state_dict" key) are the reference's. Two deliberate differences, both listed in SURVEY.md section 8(f)4:

  * the Caffe/Keras maps of the reference target "_stage3_detector_network._fc1/_fc2" (state.py:158-168,197-198) while the live
    keys are "_stage3_detector_network._pool_to_feature_vector._fc1/_fc2" (detector.py:28, vgg16.py:106-107), so the reference's
    strict load_state_dict raises after copying the conv layers and the fc layers are silently lost.  Here the keys are mapped to
    the live names and the partial state is loaded non-strictly, so fc1/fc2 are initialised as the files intend.
  * tensors are read to host memory and copied by load_state_dict into the model's device parameters (channels-last filters are
    re-laid out by ConvParams on assignment); the file is never required to have been saved from a CUDA model.
"""
import numpy as np
import torch as t

from . import ops

_KERAS_CONV_LAYERS = [
  "block1_conv1", "block1_conv2", "block2_conv1", "block2_conv2", "block3_conv1", "block3_conv2", "block3_conv3",
  "block4_conv1", "block4_conv2", "block4_conv3", "block5_conv1", "block5_conv2", "block5_conv3",
]

_CAFFE_LAYERS = {
  "features.0": "_stage1_feature_extractor._block1_conv1",
  "features.2": "_stage1_feature_extractor._block1_conv2",
  "features.5": "_stage1_feature_extractor._block2_conv1",
  "features.7": "_stage1_feature_extractor._block2_conv2",
  "features.10": "_stage1_feature_extractor._block3_conv1",
  "features.12": "_stage1_feature_extractor._block3_conv2",
  "features.14": "_stage1_feature_extractor._block3_conv3",
  "features.17": "_stage1_feature_extractor._block4_conv1",
  "features.19": "_stage1_feature_extractor._block4_conv2",
  "features.21": "_stage1_feature_extractor._block4_conv3",
  "features.24": "_stage1_feature_extractor._block5_conv1",
  "features.26": "_stage1_feature_extractor._block5_conv2",
  "features.28": "_stage1_feature_extractor._block5_conv3",
  "classifier.0": "_stage3_detector_network._pool_to_feature_vector._fc1",
  "classifier.3": "_stage3_detector_network._pool_to_feature_vector._fc2",
}


def _keras_layer(hdf5_file, layer_name):
  """(kernel, bias) of one Keras layer as host tensors, or (None, None). Reference: state.py:14-82."""
  group = "model_weights/" + layer_name
  if group not in hdf5_file:
    return None, None
  for sub in hdf5_file[group].keys():
    if sub.startswith("conv") or sub.startswith("dense"):
      kernel = np.array(hdf5_file["/".join([group, sub, "kernel:0"])]).astype(np.float32)
      bias = np.array(hdf5_file["/".join([group, sub, "bias:0"])]).astype(np.float32)
      return t.from_numpy(kernel), t.from_numpy(bias)
  return None, None


def keras_vgg16_to_state(hdf5_file):
  """Key/layout conversion of a trzy/VGG16 Keras file (any mapping with h5py's interface). Reference: state.py:117-176."""
  state, missing = {}, []
  for name in _KERAS_CONV_LAYERS:
    kernel, bias = _keras_layer(hdf5_file, name)
    if kernel is None:
      missing.append(name)
      continue
    state["_stage1_feature_extractor._%s.weight" % name] = kernel.permute(3, 2, 0, 1).contiguous()   # (kh,kw,ci,co) -> (co,ci,kh,kw)
    state["_stage1_feature_extractor._%s.bias" % name] = bias
  kernel, bias = _keras_layer(hdf5_file, "fc1")
  if kernel is not None:
    # Keras flattens the RoI pool output as (7,7,512); the reference model flattens (512,7,7): state.py:146-157
    kernel = kernel.reshape(7, 7, 512, 4096).permute(2, 0, 1, 3).reshape(-1, 4096).permute(1, 0).contiguous()
    state["_stage3_detector_network._pool_to_feature_vector._fc1.weight"] = kernel
    state["_stage3_detector_network._pool_to_feature_vector._fc1.bias"] = bias
  else:
    missing.append("fc1")
  kernel, bias = _keras_layer(hdf5_file, "fc2")
  if kernel is not None:
    state["_stage3_detector_network._pool_to_feature_vector._fc2.weight"] = kernel.permute(1, 0).contiguous()
    state["_stage3_detector_network._pool_to_feature_vector._fc2.bias"] = bias
  else:
    missing.append("fc2")
  return state, missing


def caffe_vgg16_to_state(caffe):
  """Key conversion of the published Caffe VGG-16 state dict ("vgg16_caffe.pth"). Reference: state.py:178-219."""
  state = {}
  missing = set(_CAFFE_LAYERS.keys())
  for layer, ours in _CAFFE_LAYERS.items():
    if layer + ".weight" in caffe and layer + ".bias" in caffe:
      state[ours + ".weight"] = caffe[layer + ".weight"]
      state[ours + ".bias"] = caffe[layer + ".bias"]
      missing.discard(layer)
  if len(missing) == len(_CAFFE_LAYERS):
    raise ValueError("not a Caffe VGG-16 model")
  return state, sorted(missing)


def load(model, filepath):
  """
  Loads weights into `model` from a Keras VGG-16 h5 file, a Caffe VGG-16 .pth file or a complete checkpoint
  ({"epoch", "model_state_dict"}) written by save()/BestWeightsTracker or by the reference itself (same state-dict keys).
  Reference: state.py:221-272 (errors from load_state_dict are printed, not raised, as there).
  """
  state, partial = None, False
  try:
    import h5py
    with h5py.File(filepath, "r") as f:
      state, missing = keras_vgg16_to_state(f)
    partial = True
    print("Loaded initial VGG-16 layer weights from Keras model '%s'" % filepath)
  except Exception:
    state = None
  if state is None:
    try:
      state, missing = caffe_vgg16_to_state(t.load(filepath, map_location = "cpu"))
      partial = True
      print("Loaded initial VGG-16 layer weights from Caffe model '%s'" % filepath)
    except Exception:
      state = None
  if state is None:
    state = t.load(filepath, map_location = "cpu")
    if "model_state_dict" not in state:
      raise KeyError("Model state file '%s' is missing top-level key 'model_state_dict'" % filepath)
    state = state["model_state_dict"]
    missing = []
  if len(missing) > 0:
    print("Some layers were missing from '%s' and not loaded: %s" % (filepath, ", ".join(missing)))
  try:
    model.load_state_dict(state, strict = not partial)
    ops.invalidate_weight_splits()               # (load_state_dict bumps the version counters; explicit for loaders that write through .data)
    print("Loaded initial weights from '%s'" % filepath)
  except Exception as e:
    print(e)


def save(model, filepath, epoch = 0):
  """Per-epoch checkpoint in the reference's format (__main__.py:195-198). Optimizer state is not saved, as there."""
  t.save({"epoch": epoch, "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()}}, filepath)


class BestWeightsTracker:
  """Reference: state.py:274-288."""
  def __init__(self, filepath):
    self._filepath = filepath
    self._best_state = None
    self._best_mAP = 0

  def on_epoch_end(self, model, epoch, mAP):
    if mAP > self._best_mAP:
      self._best_mAP = mAP
      # snapshot to host: the live parameters keep training after this call
      self._best_state = {"epoch": epoch, "model_state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}}

  def save_best_weights(self, model):
    if self._best_state is not None:
      t.save(self._best_state, self._filepath)
      print("Saved best model weights (Mean Average Precision = %1.2f%%) to '%s'" % (self._best_mAP, self._filepath))
