"""
CPU tests of the host-side mirrors that sit either side of the hot path:
  * fasterrcnn_b200.statistics vs outputs of the reference's statistics.py (tests/golden/statistics.npz, made by
    oracle/make_golden.py running the unmodified reference on the seeded detections of oracle/golden_inputs.stats_case);
  * fasterrcnn_b200.state key/layout conversions vs the model's state-dict keys.
"""
import os
import types

import numpy as np
import pytest
import torch as t

from oracle import golden_inputs as gi
from oracle import frcnn_oracle as orc
from fasterrcnn_b200 import statistics, state

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _box(corners, cls):
  return types.SimpleNamespace(class_index = cls, class_name = str(cls), corners = corners)


@pytest.mark.parametrize("tag", list(gi.STATS_CASES))
def test_precision_recall_matches_reference(tag):
  g = np.load(os.path.join(GOLDEN, "statistics.npz"))
  calc = statistics.PrecisionRecallCurveCalculator()
  for gts, preds in gi.stats_case(tag):
    calc.add_image_results(scored_boxes_by_class_index = preds, gt_boxes = [_box(b, c) for b, c in gts])
  classes = sorted(calc._object_count_by_class_index.keys())
  assert classes == list(g[tag + "_classes"])
  tp = [sum(1 for p in calc._unsorted_predictions_by_class_index[c] if p[1]) for c in classes]
  npred = [len(calc._unsorted_predictions_by_class_index[c]) for c in classes]
  assert tp == list(g[tag + "_tp"])                        # the greedy matching visits pairs in the reference's order
  assert npred == list(g[tag + "_npred"])
  aps = calc.compute_class_average_precisions()
  np.testing.assert_array_equal(np.array([aps[c] for c in classes]), g[tag + "_ap"])       # same arithmetic -> same doubles
  assert calc.compute_mean_average_precision() == g[tag + "_map"]


def test_precision_recall_edge_cases():
  calc = statistics.PrecisionRecallCurveCalculator()
  # objects but no predictions at all: AP 0 for that class, counted in the mean
  calc.add_image_results(scored_boxes_by_class_index = {}, gt_boxes = [_box(np.array([0, 0, 10, 10], dtype = np.float32), 3)])
  # predictions for a class with no objects in the image: all false positives, class not counted in the mean
  calc.add_image_results(scored_boxes_by_class_index = {5: np.array([[0, 0, 10, 10, 0.9]], dtype = np.float32)}, gt_boxes = [])
  assert calc.compute_mean_average_precision() == 0.0
  assert list(calc._object_count_by_class_index.keys()) == [3]
  # a perfect detection
  calc = statistics.PrecisionRecallCurveCalculator()
  b = np.array([10, 10, 50, 60], dtype = np.float32)
  calc.add_image_results(scored_boxes_by_class_index = {1: np.concatenate([b, [0.8]])[None, :].astype(np.float32)}, gt_boxes = [_box(b, 1)])
  assert calc.compute_mean_average_precision() == 1.0


def test_training_statistics_running_means():
  st = statistics.TrainingStatistics()
  for i in range(3):
    st.on_training_step(types.SimpleNamespace(rpn_class = 1.0 + i, rpn_regression = 0.5, detector_class = 2.0 * i, detector_regression = 0.25))
  assert st.rpn_class_loss == 2.0 and st.detector_class_loss == 2.0
  assert st.get_progbar_postfix()["total_loss"] == "4.75"


def test_caffe_vgg16_key_map_targets_live_keys():
  shapes = orc.vgg16_param_shapes()
  caffe = {}
  for layer, ours in state._CAFFE_LAYERS.items():
    caffe[layer + ".weight"] = t.zeros(shapes[ours + ".weight"])
    caffe[layer + ".bias"] = t.zeros(shapes[ours + ".bias"])
  caffe["classifier.6.weight"] = t.zeros((1000, 4096))      # ImageNet head: ignored
  converted, missing = state.caffe_vgg16_to_state(caffe)
  assert missing == []
  assert set(converted.keys()) <= set(shapes.keys())         # every key exists in the model (the reference's fc keys do not)
  assert "_stage3_detector_network._pool_to_feature_vector._fc1.weight" in converted
  del caffe["features.28.weight"]
  _, missing = state.caffe_vgg16_to_state(caffe)
  assert missing == ["features.28"]
  with pytest.raises(ValueError):
    state.caffe_vgg16_to_state({"something.else": t.zeros(1)})


class _FakeH5(dict):
  """Just enough of h5py.File: nested groups addressed by '/'-joined paths."""
  def __contains__(self, path):
    try:
      self[path]
      return True
    except KeyError:
      return False

  def __getitem__(self, path):
    node = dict(self)
    for part in path.split("/"):
      node = node[part]
    return node


def test_keras_vgg16_layout_conversion():
  rng = np.random.RandomState(0)
  layers = {}
  shapes = orc.vgg16_param_shapes()
  for name in state._KERAS_CONV_LAYERS:
    co, ci, kh, kw = shapes["_stage1_feature_extractor._%s.weight" % name]
    layers[name] = {"conv2d": {"kernel:0": rng.randn(kh, kw, ci, co).astype(np.float32), "bias:0": rng.randn(co).astype(np.float32)}}
  fc1 = rng.randn(25088, 8).astype(np.float32)
  # a narrow fc1 keeps the test light; the conversion only relies on the (7,7,512,out) factorisation of the rows
  layers["fc2"] = {"dense": {"kernel:0": rng.randn(16, 8).astype(np.float32), "bias:0": rng.randn(8).astype(np.float32)}}
  f = _FakeH5({"model_weights": layers})
  converted, missing = state.keras_vgg16_to_state(f)
  assert missing == ["fc1"]
  k = layers["block3_conv2"]["conv2d"]["kernel:0"]
  w = converted["_stage1_feature_extractor._block3_conv2.weight"].numpy()
  assert w.shape == (256, 256, 3, 3)
  assert w[5, 7, 1, 2] == k[1, 2, 7, 5]
  assert converted["_stage3_detector_network._pool_to_feature_vector._fc2.weight"].shape == (8, 16)
  # fc1 row permutation: Keras row (y,x,c) -> reference column c*49 + y*7 + x   (state.py:146-157)
  kernel = t.from_numpy(fc1).reshape(7, 7, 512, 8).permute(2, 0, 1, 3).reshape(-1, 8).permute(1, 0)
  y, x, c, o = 3, 5, 100, 2
  assert kernel[o, c * 49 + y * 7 + x] == fc1[(y * 7 + x) * 512 + c, o]


# ---------------------------------------------------------------- data path: VOC iterator vs the reference's own output
def _voc_rows(ds, epochs):
  import hashlib
  sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
  names, rows, shapes = [], [], []
  for _ in range(epochs):
    for smp in ds:
      names.append(os.path.basename(smp.filepath))
      rows.append([sha(smp.image_data), sha(smp.anchor_map), sha(smp.anchor_valid_map), sha(smp.gt_rpn_map),
                   sha(np.asarray(smp.gt_rpn_object_indices, dtype = np.int64)), sha(np.asarray(smp.gt_rpn_background_indices, dtype = np.int64)),
                   sha(np.array([b.corners for b in smp.gt_boxes], dtype = np.float64)), sha(np.array([b.class_index for b in smp.gt_boxes], dtype = np.int64))])
      shapes.append(tuple(smp.image_data.shape))
  return names, rows, shapes


def _oracle_anchor_fns():
  from oracle import frcnn_oracle as orc

  def rpn_map(anchor_map, anchor_valid_map, gt_boxes):
    return orc.generate_rpn_map(anchor_map, anchor_valid_map, np.array([b.corners for b in gt_boxes], dtype = np.float32))
  return (lambda image_shape, feature_map_shape, feature_pixels: orc.generate_anchor_maps(image_shape, feature_map_shape, feature_pixels)), rpn_map


@pytest.mark.parametrize("prefetch", [0, 2])
def test_voc_dataset_matches_reference_golden(tmp_path, golden_dir, prefetch):
  """fasterrcnn_b200.datasets.voc.Dataset on the synthetic VOC tree: same sample order, flips, image tensors (bit-exact), ground-truth
  boxes and RPN maps as the UNMODIFIED reference produced for the same tree and seed (tests/golden/voc.npz, oracle/make_golden.py);
  the prefetching iterator returns the same stream.  (Anchor / RPN-map functions injected from the CPU oracle: no GPU here.)"""
  import random
  from oracle import golden_inputs as gi
  from fasterrcnn_b200.backbone import ChannelOrder, PreprocessingParams
  from fasterrcnn_b200.datasets import voc
  g = np.load(os.path.join(golden_dir, "voc.npz"))
  d = gi.make_voc_tree(str(tmp_path))
  params = PreprocessingParams(channel_order = ChannelOrder.BGR, scaling = 1.0, means = [103.939, 116.779, 123.680], stds = [1, 1, 1])
  shape_fn = lambda s: (512, s[-2] // 16, s[-1] // 16)
  for tag, augment, shuffle, epochs in (("vgg", True, True, 2), ("plain", False, False, 1)):
    random.seed(1234)
    ds = voc.Dataset(split = "trainval", image_preprocessing_params = params, compute_feature_map_shape_fn = shape_fn, feature_pixels = 16, dir = d,
                     augment = augment, shuffle = shuffle, cache = False, prefetch = prefetch, anchor_fns = _oracle_anchor_fns())
    assert ds.num_samples == int(g["num_samples"]) and ds.num_classes == 21 and ds.class_index_to_name[15] == "person"
    names, rows, shapes = _voc_rows(ds, epochs)
    assert names == list(g[tag + "_names"])
    assert [list(r) for r in rows] == [list(r) for r in g[tag + "_sha"]]
    for i, sh in enumerate(shapes):
      assert sh == tuple(g["%s_shape_%d" % (tag, i)]) and min(sh[1:]) == 600
  # caches, the `difficult` rule and the error path
  ds = voc.Dataset(split = "trainval", image_preprocessing_params = params, compute_feature_map_shape_fn = shape_fn, dir = d, augment = False, shuffle = False,
                   cache = True, anchor_fns = _oracle_anchor_fns())
  first = [s for s in ds]
  again = [s for s in ds]
  assert all(a is b for a, b in zip(first, again))
  assert len(first[1].gt_boxes) + 1 == len(voc.Dataset(split = "trainval", image_preprocessing_params = params, compute_feature_map_shape_fn = shape_fn, dir = d,
                                                        allow_difficult = True, anchor_fns = _oracle_anchor_fns())._gt_boxes_by_filepath[first[1].filepath])
  with pytest.raises(FileNotFoundError):
    voc.Dataset(split = "trainval", image_preprocessing_params = params, compute_feature_map_shape_fn = shape_fn, dir = str(tmp_path / "missing"))


# ---- bench.py contract (CPU-checkable part) ----------------------------------------------------
def test_bench_reference_arm_prints_the_contract_line():
  """`bench.py --impl reference` (the CPU port of the reference step on the host cores) prints ONE JSON line with the keys the driver
  reads; `bench.py` itself refuses to run without a CUDA device (no CPU path for the product arm)."""
  import json
  import subprocess
  import sys
  import torch as t
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output = True, text = True, timeout = 600, cwd = root)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
  assert d["metric"].startswith("images/sec fwd+bwd @ 1000x600")
  assert d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
  assert d["e2e"] == dict(value = d["value"], unit = "images/s", h2d_bytes_per_step = 0, d2h_bytes_per_step = 0)
  assert "workload" in d["config"] and "model" not in d["config"]
  if not t.cuda.is_available():
    ours = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output = True, text = True, timeout = 600, cwd = root)
    assert ours.returncode != 0 and "no CPU path" in (ours.stderr + ours.stdout)


# ---- host half of proposal sampling (faster_rcnn.py:526-561) --------------------------------------
def test_proposal_sampling_draws_match_the_restated_reference():
  """FasterRCNNModel._sample_proposal_indices (the one piece of train_step's data path that stays on the host, so that seeded runs
  draw the same samples as the reference) against the pinned restatement: same rows, same order, same RNG consumption -- on label
  vectors with many / few / no positives, fewer rows than the batch, and sampling switched off."""
  import fasterrcnn_b200 as f
  model = f.FasterRCNNModel(21, f.vgg16.VGG16Backbone(0.0))
  rng = np.random.RandomState(3)
  cases = [(2002, 0.03), (2002, 0.5), (300, 0.0), (300, 1.0), (40, 0.3), (128, 0.25), (5, 0.4)]
  for n, pos_rate in cases:
    cls = np.where(rng.rand(n) < pos_rate, rng.randint(1, 21, n), 0).astype(np.int32)
    onehot = t.zeros((n, 21)); onehot[t.arange(n), t.from_numpy(cls).long()] = 1.0
    rows = t.arange(n, dtype = t.float32).reshape(n, 1).repeat(1, 4)                     # "proposals" that carry their own row index
    deltas = t.zeros((n, 2, 80))
    t.manual_seed(n); ref_rows, _, _ = orc.sample_proposals(rows, onehot, deltas, 128, 0.25)
    after_ref = t.rand(1).item()                                                        # where the generator stands afterwards
    t.manual_seed(n); idx = model._sample_proposal_indices(cls, 128, 0.25)
    after = t.rand(1).item()
    assert np.array_equal(idx, ref_rows[:, 0].numpy().astype(np.int64)), (n, pos_rate)
    assert after == after_ref
  assert model._sample_proposal_indices(np.zeros((7,), np.int32), 0, 0.25) is None       # max_proposals <= 0: keep every row


def test_math_utils_host_helpers_equal_the_reference_bit_for_bit():
  """math_utils.intersection_over_union / convert_deltas_to_boxes (NumPy, host) against the oracle's restatement and -- where the
  reference tree exists -- against the reference's own functions on the same inputs, bit for bit, float32 and float64."""
  import importlib.util
  import os
  from fasterrcnn_b200 import math_utils
  from oracle import frcnn_oracle as orc
  from oracle import golden_inputs as gi
  rng = np.random.default_rng(5)
  ref = None
  path = "/root/reference/pytorch/FasterRCNN/models/math_utils.py"
  if os.path.exists(path):
    spec = importlib.util.spec_from_file_location("ref_math_utils", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
  for dtype in (np.float32, np.float64):
    b1 = gi.random_boxes(rng, 200, dtype = dtype)
    b2 = np.concatenate([gi.random_boxes(rng, 7, dtype = dtype), b1[:3], np.array([[5, 5, 5, 9], [0, 0, 0, 0]], dtype = dtype)])    # identical + degenerate boxes
    got = math_utils.intersection_over_union(b1, b2)
    assert np.array_equal(got, orc.iou_np(b1, b2))
    deltas = (rng.normal(0, 0.4, (200, 4))).astype(dtype)
    anchors = np.stack([rng.uniform(0, 600, 200), rng.uniform(0, 1000, 200), rng.uniform(16, 512, 200), rng.uniform(16, 512, 200)], axis = 1).astype(dtype)
    means, stds = np.array([0, 0, 0, 0], dtype = dtype), np.array([0.1, 0.1, 0.2, 0.2], dtype = dtype)
    boxes = math_utils.convert_deltas_to_boxes(deltas, anchors, means, stds)
    assert boxes.dtype == np.float64                                    # np.empty's default, as in the reference (predict relies on float64)
    assert np.array_equal(boxes, orc.deltas_to_boxes_np(deltas, anchors, means, stds))
    if ref is not None:
      assert np.array_equal(got, ref.intersection_over_union(b1, b2))
      assert np.array_equal(boxes, ref.convert_deltas_to_boxes(deltas, anchors, means, stds))


def test_resnet_frozen_bn_fold_and_stride2_form_selection(monkeypatch):
  """Host logic of resnet._ConvBNAct: (1) the frozen-BN fold equals BatchNorm2d.eval()'s affine map and is cached until a tensor changes
  (resnet.py:56-77: gamma / sqrt(var + eps), beta - mean * that); (2) which tensor-core form a strided convolution takes."""
  from fasterrcnn_b200 import resnet
  g = t.Generator().manual_seed(3)
  bn = resnet.BNParams(16)
  ref = t.nn.BatchNorm2d(16).eval()
  with t.no_grad():
    for name in ("weight", "bias", "running_mean"):
      v = t.randn((16,), generator = g)
      getattr(bn, name).copy_(v); getattr(ref, name).copy_(v)
    v = t.rand((16,), generator = g) + 0.5
    bn.running_var.copy_(v); ref.running_var.copy_(v)
  x = t.randn((2, 16, 5, 7), generator = g)
  scale, shift = bn.folded()
  np.testing.assert_allclose((x * scale[None, :, None, None] + shift[None, :, None, None]).numpy(), ref(x).detach().numpy(), rtol = 1e-5, atol = 1e-6)
  assert bn.folded()[0] is scale                                  # cached
  with t.no_grad():
    bn.weight.mul_(2.0)                                           # a version bump (what load_state_dict does) refolds
  assert bn.folded()[0] is not scale and t.allclose(bn.folded()[0], 2.0 * scale)

  w1, w3, w7 = t.empty((8, 4, 1, 1)), t.empty((8, 4, 3, 3)), t.empty((8, 3, 7, 7))
  assert resnet._strided_on_tensor_cores(None, w1, 2, 0) == "1x1"      # downsample: conv1x1(x[::2, ::2])
  assert resnet._strided_on_tensor_cores(None, w3, 2, 1) == "3x3"      # conv2: conv3x3(x)[::2, ::2]
  assert resnet._strided_on_tensor_cores(None, w3, 1, 1) is None       # stride 1: the engine's own geometry
  assert resnet._strided_on_tensor_cores(None, w7, 2, 3) is None       # the frozen 7x7 stem stays on the CUDA-core engine
  monkeypatch.setenv("FRCNN_RESNET_S2_TC", "0")
  assert resnet._strided_on_tensor_cores(None, w3, 2, 1) is None
