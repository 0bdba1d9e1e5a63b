"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE
(/root/reference, CPU, via oracle/ref_shim.py) and torchvision's own CPU ops on seeded inputs.

Run in the build container (the reference does not exist on the GPU box):
    python -m oracle.make_golden
Inputs are regenerated from seeds by tests (recipes in oracle/golden_inputs.py); the fixtures
hold the reference's OUTPUTS (small arrays) or SHA-256 digests of large bit-exact outputs.
"""
import hashlib
import os
import random
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, golden_inputs as gi          # noqa: E402
from oracle import frcnn_oracle as orc                    # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def statistics_golden(ref):
  """statistics.PrecisionRecallCurveCalculator of the reference on seeded detections (golden_inputs.stats_case)."""
  import importlib
  stats_mod = importlib.import_module("pytorch.FasterRCNN.statistics")
  out = {}
  for tag in gi.STATS_CASES:
    calc = stats_mod.PrecisionRecallCurveCalculator()
    for gts, preds in gi.stats_case(tag):
      boxes = [ref.Box(class_index = c, class_name = str(c), corners = b) for b, c in gts]
      calc.add_image_results(scored_boxes_by_class_index = preds, gt_boxes = boxes)
    classes = sorted(calc._object_count_by_class_index.keys())
    out[tag + "_classes"] = np.array(classes, dtype = np.int32)
    out[tag + "_ap"] = np.array([calc._compute_average_precision(class_index = c)[0] for c in classes], dtype = np.float64)
    out[tag + "_map"] = np.float64(calc.compute_mean_average_precision())
    out[tag + "_tp"] = np.array([sum(1 for p in calc._unsorted_predictions_by_class_index[c] if p[1]) for c in classes], dtype = np.int32)
    out[tag + "_npred"] = np.array([len(calc._unsorted_predictions_by_class_index[c]) for c in classes], dtype = np.int32)
  np.savez_compressed(os.path.join(OUT, "statistics.npz"), **out)


def voc_golden(ref):
  """datasets/voc.py of the reference on the synthetic VOC tree (golden_inputs.make_voc_tree): two shuffled, augmented epochs with the
  VGG-16 preprocessing and one un-augmented pass with the ResNet preprocessing -- per sample the file, the digest of every array of
  the TrainingSample and the ground-truth boxes."""
  import importlib
  import random
  import tempfile
  from PIL import Image
  sys.modules["imageio"].imread = lambda url, pilmode = "RGB": np.array(Image.open(url).convert(pilmode))       # imageio v2 semantics
  voc = importlib.import_module("pytorch.FasterRCNN.datasets.voc")
  vgg = ref.vgg16.VGG16Backbone(dropout_probability = 0.0)
  out = {}
  with tempfile.TemporaryDirectory() as tmp:
    d = gi.make_voc_tree(tmp)
    for tag, params, shape_fn, augment, shuffle, epochs in (
        ("vgg", vgg.image_preprocessing_params, vgg.compute_feature_map_shape, True, True, 2),
        ("plain", vgg.image_preprocessing_params, vgg.compute_feature_map_shape, False, False, 1)):
      random.seed(1234)
      ds = voc.Dataset(split = "trainval", image_preprocessing_params = params, compute_feature_map_shape_fn = shape_fn, feature_pixels = 16, dir = d,
                       augment = augment, shuffle = shuffle, cache = False)
      names, rows = [], []
      for _ in range(epochs):
        for smp in ds:
          names.append(os.path.basename(smp.filepath))
          rows.append([sha(smp.image_data), sha(smp.anchor_map), sha(smp.anchor_valid_map), sha(smp.gt_rpn_map),
                       sha(np.asarray(smp.gt_rpn_object_indices, dtype = np.int64)), sha(np.asarray(smp.gt_rpn_background_indices, dtype = np.int64)),
                       sha(np.array([b.corners for b in smp.gt_boxes], dtype = np.float64)), sha(np.array([b.class_index for b in smp.gt_boxes], dtype = np.int64))])
          out["%s_shape_%d" % (tag, len(names) - 1)] = np.array(smp.image_data.shape)
      out[tag + "_names"] = np.array(names)
      out[tag + "_sha"] = np.array(rows)
    out["num_samples"] = np.int64(ds.num_samples)
  np.savez_compressed(os.path.join(OUT, "voc.npz"), **out)
  print("voc golden:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and k.endswith(("_names", "_sha"))})


def main():
  os.makedirs(OUT, exist_ok = True)
  ref = ref_shim.load()
  if "--only-statistics" in sys.argv:
    statistics_golden(ref)
    return
  if "--only-voc" in sys.argv:
    voc_golden(ref)
    return
  tv = ref.torchvision
  t.set_num_threads(8)

  # ---- 1. geometry: anchors + RPN ground truth (models/anchors.py) -------------------------
  geo = {}
  for tag, (h, w) in gi.GEOMETRY_CASES.items():
    fm = (512, h // 16, w // 16)
    am, av = ref.anchors.generate_anchor_maps(image_shape = (3, h, w), feature_map_shape = fm, feature_pixels = 16)
    boxes = [ref.Box(class_index = c, class_name = str(c), corners = np.array(b, dtype = np.float32)) for b, c in gi.gt_boxes_for(h, w)]
    rm, obj, bg = ref.anchors.generate_rpn_map(anchor_map = am, anchor_valid_map = av, gt_boxes = boxes)
    geo[tag + "_anchor_sha"] = sha(am)
    geo[tag + "_valid_sha"] = sha(av)
    geo[tag + "_rpnmap_sha"] = sha(rm)
    geo[tag + "_obj"] = obj.astype(np.int16)
    geo[tag + "_nbg"] = np.int64(len(bg))
    geo[tag + "_bg_sha"] = sha(bg.astype(np.int64))
    if tag == "tiny":
      geo["tiny_anchor_map"] = am
      geo["tiny_valid_map"] = av
      geo["tiny_rpn_map"] = rm
  np.savez_compressed(os.path.join(OUT, "geometry.npz"), **geo)

  # ---- 2. torchvision CPU ops: nms + roi_pool --------------------------------------------
  ops = {}
  for tag in gi.NMS_CASES:
    boxes, scores, thr = gi.nms_case(tag)
    keep = tv.ops.nms(t.from_numpy(boxes), t.from_numpy(scores.astype(boxes.dtype)), thr).numpy()
    ops["nms_" + tag] = keep.astype(np.int32)
  for tag in gi.ROI_CASES:
    fm, rois = gi.roi_case(tag)
    fmt = t.from_numpy(fm).requires_grad_(True)
    out = tv.ops.roi_pool(fmt, t.from_numpy(rois), (7, 7), 1.0 / 16.0)
    go = gi.roi_grad(tag, out.shape)
    out.backward(t.from_numpy(go))
    ops["roi_%s_out_sha" % tag] = sha(out.detach().numpy())
    ops["roi_%s_gin_sha" % tag] = sha(fmt.grad.numpy())
    if tag == "small":
      ops["roi_small_out"] = out.detach().numpy()
      ops["roi_small_gin"] = fmt.grad.numpy()
  np.savez_compressed(os.path.join(OUT, "tv_ops.npz"), **ops)

  # ---- 3. RPN proposal stage in isolation (models/rpn.py:99-156) -------------------------
  rp = {}
  rpn_mod = ref.rpn.RegionProposalNetwork(feature_map_channels = 512, allow_edge_proposals = True)
  for tag in gi.RPN_CASES:
    c = gi.rpn_case(tag)
    am, av = ref.anchors.generate_anchor_maps(image_shape = c["image_shape"], feature_map_shape = (512,) + c["fm_hw"], feature_pixels = 16)
    # Re-run the reference's own proposal code on given maps: call forward's tail through a
    # stub feature path (conv layers replaced by the given maps).
    score_map, delta_map = t.from_numpy(c["score_map"]), t.from_numpy(c["delta_map"])
    props = _reference_rpn_tail(ref, rpn_mod, score_map, delta_map, am, av, c["image_shape"], c["pre_nms"], c["post_nms"])
    rp[tag + "_proposals"] = props.numpy()
  np.savez_compressed(os.path.join(OUT, "rpn_stage.npz"), **rp)

  # ---- 4. end-to-end on a small image: forward / predict / train_step --------------------
  from oracle import resnet_oracle
  ref_shim.patch_resnet_offline()
  for tag in gi.E2E_CASES:
    e2e = {}
    cfg = gi.E2E_CASES[tag]
    h, w = cfg["hw"]
    if cfg["backbone"] == "vgg16":
      params = orc.synth_params(orc.vgg16_param_shapes(), seed = cfg["weight_seed"], heads = cfg["heads"])
      backbone = ref.vgg16.VGG16Backbone(dropout_probability = 0.0)
    else:
      params = orc.synth_params(resnet_oracle.param_shapes(cfg["backbone"]), seed = cfg["weight_seed"], heads = cfg["heads"])
      backbone = ref.resnet.ResNetBackbone(architecture = {"resnet50": ref.resnet.Architecture.ResNet50, "resnet101": ref.resnet.Architecture.ResNet101}[cfg["backbone"]])
    model = ref.faster_rcnn.FasterRCNNModel(num_classes = 21, backbone = backbone, allow_edge_proposals = True)
    model.load_state_dict(params)
    smp = orc.synthetic_sample((h, w), seed = cfg["sample_seed"], backbone = cfg["backbone"])
    image = smp["image"]

    model.eval()
    with t.no_grad():
      props, classes, deltas = model(image_data = image)
    e2e[tag + "_fwd_proposals"] = props.numpy()
    e2e[tag + "_fwd_classes"] = classes.numpy()
    e2e[tag + "_fwd_deltas"] = deltas.numpy()
    pred = model.predict(image_data = image, score_threshold = cfg["score_threshold"])
    e2e[tag + "_pred_counts"] = np.array([pred[c].shape[0] for c in range(1, 21)], dtype = np.int32)
    e2e[tag + "_pred_boxes"] = np.concatenate([pred[c] for c in range(1, 21)], axis = 0)

    # train_step x2 with the reference's optimizer recipe (__main__.py:98-105)
    opt_params = []
    for key, value in dict(model.named_parameters()).items():
      if value.requires_grad and "weight" in key:
        opt_params += [{"params": [value], "weight_decay": 5e-4}]
    optimizer = t.optim.SGD(opt_params, lr = 1e-3, momentum = 0.9)
    boxes = [ref.Box(class_index = c, class_name = str(c), corners = b) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
    random.seed(cfg["sample_seed"]); np.random.seed(cfg["sample_seed"]); t.manual_seed(cfg["sample_seed"])
    losses = []
    for step in range(2):
      loss = model.train_step(
        optimizer = optimizer, image_data = image, anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
        gt_rpn_map = smp["gt_rpn_map"], gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
        gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
      losses.append([loss.rpn_class, loss.rpn_regression, loss.detector_class, loss.detector_regression, loss.total])
      if step == 0:
        for key, p in model.named_parameters():
          if p.grad is not None:
            g = p.grad
            e2e["%s_grad_norm/%s" % (tag, key)] = np.float64(g.double().norm().item())
            e2e["%s_grad_head/%s" % (tag, key)] = g.reshape(-1)[:64].numpy().copy()
    e2e[tag + "_losses"] = np.array(losses, dtype = np.float64)
    sd = model.state_dict()
    for key in sd:
      e2e["%s_w2_head/%s" % (tag, key)] = sd[key].reshape(-1)[:64].numpy().copy()
      e2e["%s_w2_norm/%s" % (tag, key)] = np.float64(sd[key].double().norm().item())
    np.savez_compressed(os.path.join(OUT, "e2e_%s.npz" % cfg["backbone"]), **e2e)
  statistics_golden(ref)
  print("golden vectors written to", OUT)
  for f in sorted(os.listdir(OUT)):
    print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


def _reference_rpn_tail(ref, rpn_mod, score_map, delta_map, anchor_map, anchor_valid_map, image_shape, pre_nms, post_nms):
  """Runs RegionProposalNetwork.forward (rpn.py:51-156) with its three conv layers replaced by
  functions that emit the given maps, so the reference's own decode/sort/clip/filter/NMS code runs."""
  class Emit(t.nn.Module):
    def __init__(self, value):
      super().__init__()
      self.value = value
    def forward(self, x):
      return self.value
  import torch.nn.functional as F
  saved = (rpn_mod._rpn_conv1, rpn_mod._rpn_class, rpn_mod._rpn_boxes)
  # score map is post-sigmoid: feed logit so that sigmoid(logit) round-trips exactly is not
  # guaranteed -> bypass sigmoid by patching t.sigmoid inside the rpn module for this call.
  rpn_mod._rpn_conv1 = Emit(t.zeros((1, 1, 1, 1)))
  rpn_mod._rpn_class = Emit(score_map.permute(0, 3, 1, 2).contiguous())
  rpn_mod._rpn_boxes = Emit(delta_map.permute(0, 3, 1, 2).contiguous())
  real_t = ref.rpn.t
  class NoSigmoid:
    def __getattr__(self, k):
      return getattr(real_t, k)
    @staticmethod
    def sigmoid(x):
      return x
  ref.rpn.t = NoSigmoid()
  try:
    _, _, props = rpn_mod(feature_map = t.zeros((1, 512, 1, 1)), image_shape = image_shape, anchor_map = anchor_map, anchor_valid_map = anchor_valid_map, max_proposals_pre_nms = pre_nms, max_proposals_post_nms = post_nms)
  finally:
    ref.rpn.t = real_t
    rpn_mod._rpn_conv1, rpn_mod._rpn_class, rpn_mod._rpn_boxes = saved
  return props.detach()


if __name__ == "__main__":
  main()
