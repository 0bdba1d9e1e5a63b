"""
GPU end-to-end parity: the drop-in model (fasterrcnn_b200.FasterRCNNModel) against the CPU
oracle on the same seeded weights and sample -- forward, predict and two train_steps.
Discontinuous stages (top-N cut, >=16 filter, IoU > 0.7) are compared exactly on indices where
the upstream floating-point agreement leaves a margin; floats within the stated tolerance.
"""
import os
import random

import numpy as np
import pytest
import torch as t

from oracle import frcnn_oracle as orc
from oracle import golden_inputs as gi

import _margins

pytestmark = pytest.mark.gpu


class Box:
  def __init__(self, corners, class_index):
    self.corners, self.class_index, self.class_name = corners, class_index, str(class_index)


# End-to-end bars = ~2x the worst value MEASURED on the B200 over every end-to-end case (profiles/r02_parity_margins.md: VGG-16 and
# ResNet-50/101, 384x512 .. 600x1000, batch 1 and 2).  The stage-isolated tests (test_kernels_gpu.py) hold the bit-exact / 1e-4 bars of
# the north star; these bound what 13-100 stacked fp32 convolutions in another summation order can move through exp() and the
# discontinuous selections.  The oracle's own decision margins are recorded next to them (min |IoU - 0.7| down to 5e-7, top-N score gap
# down to 1e-8): a decision inside the upstream floating-point agreement may legitimately flip, hence ONE unmatched row is tolerated.
PX_BAR = 1e-2                                  # a produced box and its oracle partner, px, max over the 4 coordinates   (measured <= 4.6e-3)
ROWS_BAR = 1                                   # |#rows - #oracle rows|                                                   (measured 0)
PRED_CLASSES_BAR = 19                          # classes (of 20) whose detection count equals the oracle's                (measured 20)
GRAD_BAR = 1.25e-3                             # relative L2 of every parameter gradient; full-size cases use 2x          (measured <= 5.5e-4 small, 1.2e-3 full size)
LOSS_RTOL_STEP1 = 1e-4                         # each of the five losses, first step                                      (measured <= 3.4e-5)
LOSS_RTOL_STEP2 = 3e-4                         # second step (after an SGD update, re-sampled RoIs)                       (measured <= 1.4e-4)
WEIGHT_ATOL = 1.2e-5                           # post-step weights, with rtol 1e-4                                        (measured <= 5.8e-6)


def UNMATCHED_BAR(rows):
  return max(1, int(0.004 * rows))             # rows with no oracle partner within PX_BAR                                (measured 0)


def _build(cfg):
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import resnet
  from oracle import resnet_oracle
  kind = cfg["backbone"]
  if kind == "vgg16":
    shapes = orc.vgg16_param_shapes()
    backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0)
  else:
    shapes = resnet_oracle.param_shapes(kind)
    backbone = resnet.ResNetBackbone({"resnet50": resnet.Architecture.ResNet50, "resnet101": resnet.Architecture.ResNet101}[kind])
  params = orc.synth_params(shapes, seed = cfg["weight_seed"], heads = cfg["heads"])
  model = f.FasterRCNNModel(num_classes = 21, backbone = backbone, allow_edge_proposals = True)
  model.load_state_dict(params)
  model = model.cuda()
  oracle = orc.OracleModel(params, backbone = kind)
  smp = orc.synthetic_sample(cfg["hw"], seed = cfg["sample_seed"], backbone = kind)
  return model, oracle, smp


@pytest.mark.parametrize("tag", list(gi.E2E_CASES))
def test_forward_and_predict_match_oracle(golden_dir, tag):
  cfg = gi.E2E_CASES[tag]
  model, oracle, smp = _build(cfg)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  taps = {}
  with t.no_grad():
    p_ref, c_ref, d_ref = oracle.forward(smp["image"], taps = taps)
  model.eval()
  with t.no_grad():
    fm = model._stage1_feature_extractor(image_data = smp["image"].cuda())
    props, classes, deltas = model(image_data = smp["image"].cuda())
  # stage 1: feature map within 1e-4 relative to its scale (fp32 accumulation order only)
  fm_ref = taps["feature_map"].numpy()
  np.testing.assert_allclose(fm.cpu().numpy(), fm_ref, rtol = 1e-4, atol = 1e-4 * float(np.abs(fm_ref).max()))
  # proposals: the discontinuous steps (top-N cut, >=16 filter, IoU > 0.7) may flip an isolated box when the
  # 13 stacked convs differ in the last bits, so rows are matched to their nearest oracle row; the measured deviation, the unmatched
  # rows and the oracle's own decision margins go to the margin report (profiles/r02_parity_margins.md), the bars are 2x the measured
  # values (the 1e-4 px bar is met stage-isolated, see test_kernels_gpu)
  pg, pr = props.cpu().numpy(), p_ref.numpy()
  partner, ok, m = _margins.proposal_margins(pg, pr, PX_BAR)
  fm_err = float(np.abs(fm.cpu().numpy() - fm_ref).max() / np.abs(fm_ref).max())
  cls_err = float(np.abs(classes.cpu().numpy()[ok] - c_ref.numpy()[partner[ok]]).max()) if ok.any() else 0.0
  dlt_err = float(np.abs(deltas.cpu().numpy()[ok] - d_ref.numpy()[partner[ok]]).max()) if ok.any() else 0.0
  _margins.record("forward_" + tag, feature_map_max_rel = fm_err, class_score_max_abs = cls_err, box_delta_max_abs = dlt_err,
                  **m, **_margins.decision_margins(taps, 6000, cfg["hw"]))
  assert abs(props.shape[0] - p_ref.shape[0]) <= ROWS_BAR
  assert m["unmatched_rows"] <= UNMATCHED_BAR(pg.shape[0]), m
  np.testing.assert_allclose(classes.cpu().numpy()[ok], c_ref.numpy()[partner[ok]], rtol = 0, atol = 1e-4)      # class scores within 1e-4
  np.testing.assert_allclose(deltas.cpu().numpy()[ok], d_ref.numpy()[partner[ok]], rtol = 0, atol = 1e-4)
  g = np.load(os.path.join(golden_dir, "e2e_%s.npz" % cfg["backbone"]))
  same = ok & (partner == np.arange(pg.shape[0])) if pg.shape[0] == g[tag + "_fwd_classes"].shape[0] else np.zeros(pg.shape[0], bool)
  if same.any():                                       # rows in identical position: compare with the unmodified reference's own output
    np.testing.assert_allclose(classes.cpu().numpy()[same], g[tag + "_fwd_classes"][same], rtol = 0, atol = 1e-4)

  pred = model.predict(image_data = smp["image"].cuda(), score_threshold = cfg["score_threshold"])
  ref = oracle.predict(smp["image"], cfg["score_threshold"])
  counts = np.array([pred[c].shape[0] for c in range(1, 21)])
  ref_counts = np.array([ref[c].shape[0] for c in range(1, 21)])
  assert np.array_equal(ref_counts, g[tag + "_pred_counts"])                     # oracle == reference
  worst_px, worst_score, unmatched = 0.0, 0.0, 0
  for c in range(1, 21):
    if counts[c - 1] > 0 and ref_counts[c - 1] > 0:
      partner_c, dist_c = _margins.match_rows(pred[c][:, :4], ref[c][:, :4])
      okc = dist_c <= PX_BAR
      unmatched += int((~okc).sum())
      if okc.any():
        worst_px = max(worst_px, float(dist_c[okc].max()))
        worst_score = max(worst_score, float(np.abs(pred[c][okc, 4] - ref[c][partner_c[okc], 4]).max()))
    else:
      unmatched += int(counts[c - 1])
  _margins.record("predict_" + tag, classes_with_equal_count = int((counts == ref_counts).sum()), boxes = int(counts.sum()), ref_boxes = int(ref_counts.sum()),
                  unmatched_boxes = unmatched, max_px_matched = worst_px, max_score_abs = worst_score)
  assert (counts == ref_counts).sum() >= PRED_CLASSES_BAR and abs(int(counts.sum()) - int(ref_counts.sum())) <= ROWS_BAR
  assert unmatched <= UNMATCHED_BAR(int(counts.sum())), unmatched
  assert worst_score <= 1e-4


@pytest.mark.parametrize("tag", list(gi.E2E_CASES))
def test_train_step_matches_oracle(golden_dir, tag):
  cfg = gi.E2E_CASES[tag]
  model, oracle, smp = _build(cfg)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  params = [{"params": [p], "weight_decay": 5e-4} for k, p in model.named_parameters() if p.requires_grad and "weight" in k]
  optimizer = t.optim.SGD(params, lr = 1e-3, momentum = 0.9)                     # __main__.py:98-105
  boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  g = np.load(os.path.join(golden_dir, "e2e_%s.npz" % cfg["backbone"]))

  losses, ref_losses = [], []
  for who in ("ref", "gpu"):
    random.seed(0); np.random.seed(0); t.manual_seed(0)
    for step in range(2):
      if who == "ref":
        l = oracle.train_step(smp["image"], smp["anchor_map"], smp["anchor_valid_map"], smp["gt_rpn_map"], smp["gt_rpn_object_indices"],
                              smp["gt_rpn_background_indices"], smp["gt_corners"], smp["gt_class_idxs"])
        ref_losses.append([l.rpn_class, l.rpn_regression, l.detector_class, l.detector_regression, l.total])
        if step == 0:
          ref_grads = {k: v.grad.clone() for k, v in oracle.params.items() if v.grad is not None}
      else:
        l = model.train_step(optimizer = optimizer, image_data = smp["image"].cuda(), anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                             gt_rpn_map = smp["gt_rpn_map"].cuda(), gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                             gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
        losses.append([l.rpn_class, l.rpn_regression, l.detector_class, l.detector_regression, l.total])
        if step == 0:
          grads = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
  np.testing.assert_allclose(np.array(ref_losses), g[tag + "_losses"], rtol = 1e-5, atol = 1e-6)   # oracle == reference
  np.testing.assert_allclose(np.array(losses[0]), np.array(ref_losses[0]), rtol = LOSS_RTOL_STEP1, atol = 1e-6)
  np.testing.assert_allclose(np.array(losses[1]), np.array(ref_losses[1]), rtol = LOSS_RTOL_STEP2, atol = 1e-6)   # (after an SGD update, re-sampled RoIs)
  assert set(grads) == set(ref_grads)
  rels, gm = _margins.grad_margins(grads, ref_grads)
  w_err = max(float((p.detach().cpu() - oracle.params[k].detach()).abs().max()) for k, p in model.named_parameters())
  _margins.record("train_step_" + tag, loss_rel_step1 = float(np.max(np.abs(np.array(losses[0]) - np.array(ref_losses[0])) / np.abs(np.array(ref_losses[0])))),
                  loss_rel_step2 = float(np.max(np.abs(np.array(losses[1]) - np.array(ref_losses[1])) / np.abs(np.array(ref_losses[1])))),
                  post_step_weight_max_abs = w_err, optimizer = "torch.optim.SGD", **gm)
  for k, rel in rels.items():
    assert rel < GRAD_BAR, (k, rel)                           # every parameter gradient (isolated ReLU / RoI-argmax flips)
  for k, p in model.named_parameters():                       # post-step weights
    ref_w = oracle.params[k].detach()
    np.testing.assert_allclose(p.detach().cpu().numpy(), ref_w.numpy(), rtol = 1e-4, atol = WEIGHT_ATOL)


def test_empty_and_ragged_inputs():
  """No proposals survive / tiny image: the reference's edge behaviour (empty tensors, asserts)."""
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import ops
  dev = "cuda"
  # empty RoI list through the detector head ops
  fm = t.randn((1, 512, 10, 12), device = dev)
  out = ops.roi_pool(fm, t.zeros((0, 4), device = dev))
  assert out.shape == (0, 512, 7, 7)
  y = ops.linear_act(out.reshape(0, 25088), t.zeros((16, 25088), device = dev), t.zeros((16,), device = dev), ops.ACT_RELU)
  assert y.shape == (0, 16)
  assert ops.nms(t.zeros((0, 4), device = dev), t.zeros((0,), device = dev), 0.5).numel() == 0
  # all proposals smaller than 16 px: strongly negative size deltas -> empty proposal set
  fh, fw = 10, 12
  scores = t.rand((1, fh, fw, 9), device = dev)
  deltas = t.zeros((1, fh, fw, 36), device = dev)
  deltas.view(-1, 4)[:, 2:4] = -8.0
  props = ops.rpn_proposals(scores, deltas, (3, 160, 192), 16, 6000, 300)
  assert props.shape == (0, 4)
  # batch size must be 1 (faster_rcnn.py:108)
  model = f.FasterRCNNModel(21, f.vgg16.VGG16Backbone(0.0)).cuda()
  with pytest.raises(AssertionError):
    model(image_data = t.zeros((2, 3, 64, 64), device = dev))


def _forward_vs_oracle(name, model, oracle, smp, pre_nms = 6000):
  """forward() against the oracle with the margin record; returns (partner, ok) of the proposal rows."""
  taps = {}
  with t.no_grad():
    p_ref, c_ref, d_ref = oracle.forward(smp["image"], taps = taps)
  model.eval()
  with t.no_grad():
    props, classes, deltas = model(image_data = smp["image"].cuda())
  pg, pr = props.cpu().numpy(), p_ref.numpy()
  partner, ok, m = _margins.proposal_margins(pg, pr, PX_BAR)
  row_c = np.abs(classes.cpu().numpy()[ok] - c_ref.numpy()[partner[ok]]).max(axis = 1)
  row_d = np.abs(deltas.cpu().numpy()[ok] - d_ref.numpy()[partner[ok]]).max(axis = 1)
  _margins.record(name, class_score_max_abs = float(row_c.max()), box_delta_max_abs = float(row_d.max()), rows_over_1e4 = int(((row_c > 1e-4) | (row_d > 1e-4)).sum()),
                  **m, **_margins.decision_margins(taps, pre_nms, tuple(smp["image"].shape[2:])))
  assert abs(props.shape[0] - p_ref.shape[0]) <= ROWS_BAR
  assert m["unmatched_rows"] <= UNMATCHED_BAR(pg.shape[0]), m
  # a proposal within PX_BAR of its partner can still quantise to another RoIPool cell (round(x / 16) at a .5 boundary): such an isolated
  # row gets different pooled features, so rows -- not elements -- are counted
  assert ((row_c > 1e-4) | (row_d > 1e-4)).sum() <= max(1, int(0.01 * ok.sum())), (row_c.max(), row_d.max())
  return partner, ok


def _train_step_vs_oracle(name, model, oracle, smp, optimizer, seed, loss_rtol = LOSS_RTOL_STEP1, check_weights = False):
  """One train_step against the oracle on the same RNG streams: losses, every parameter gradient, optionally post-step weights."""
  boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  random.seed(seed); np.random.seed(seed); t.manual_seed(seed)
  taps = {}
  ref = oracle.train_step(smp["image"], smp["anchor_map"], smp["anchor_valid_map"], smp["gt_rpn_map"], smp["gt_rpn_object_indices"],
                          smp["gt_rpn_background_indices"], smp["gt_corners"], smp["gt_class_idxs"], apply_update = check_weights, taps = taps)
  ref_grads = {k: v.grad.clone() for k, v in oracle.params.items() if v.grad is not None}
  random.seed(seed); np.random.seed(seed); t.manual_seed(seed)
  got = model.train_step(optimizer = optimizer, image_data = smp["image"].cuda(), anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                         gt_rpn_map = smp["gt_rpn_map"].cuda(), gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                         gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
  a = np.array([got.rpn_class, got.rpn_regression, got.detector_class, got.detector_regression, got.total])
  b = np.array([ref.rpn_class, ref.rpn_regression, ref.detector_class, ref.detector_regression, ref.total])
  grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
  rels, gm = _margins.grad_margins(grads, ref_grads)
  extra = {}
  if check_weights:
    extra["post_step_weight_max_abs"] = max(float((p.detach().cpu() - oracle.params[k].detach()).abs().max()) for k, p in model.named_parameters())
  # Did both sides train the detector on the SAME RoIs?  The proposal set feeds a seeded sampler, so one flipped NMS / top-N decision
  # upstream (the oracle's own margins go down to |IoU - 0.7| = 5e-7, score gaps to 1e-8 -- inside the fp32 agreement of two summation
  # orders) re-draws the sample: the step is then a different, equally valid one.  That case is recognised and held to the wide bars only
  # if the oracle's margins really are that thin; with the same RoIs the tight bars (2x the measured values) apply.
  sp_got, sp_ref = model.last_step_info["sampled_proposals"].detach().cpu().numpy(), taps["sampled_proposals"].numpy()
  same_rois = sp_got.shape == sp_ref.shape and bool((_margins.match_rows(sp_got, sp_ref)[1] <= PX_BAR).all())
  dm = _margins.decision_margins(taps, 12000, tuple(smp["image"].shape[2:]))
  _margins.record(name, loss_rel = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6))), num_rois = model.last_step_info["num_rois"],
                  optimizer = type(optimizer).__name__, same_sampled_rois = same_rois, **gm, **extra, **dm)
  assert set(grads) == set(ref_grads)
  if not same_rois:
    assert min(dm.get("min_iou_margin", 1.0), dm["topn_cut_gap"], dm["min_adjacent_gap"] + 1e-30) < 1e-4, ("RoI sets differ although every decision margin is wide", dm)
    np.testing.assert_allclose(a, b, rtol = 5e-3, atol = 1e-4)
    for k, rel in rels.items():
      assert rel < 5e-2, (k, rel)
    return got
  np.testing.assert_allclose(a, b, rtol = loss_rtol, atol = 1e-5)
  for k, rel in rels.items():
    assert rel < 2 * GRAD_BAR, (k, rel)
  if check_weights:
    for k, p in model.named_parameters():
      np.testing.assert_allclose(p.detach().cpu().numpy(), oracle.params[k].detach().numpy(), rtol = 1e-4, atol = WEIGHT_ATOL, err_msg = k)
  return got


def _reference_optimizer(model):
  return t.optim.SGD([{"params": [p], "weight_decay": 5e-4} for k, p in model.named_parameters() if p.requires_grad and "weight" in k], lr = 1e-3, momentum = 0.9)


def test_config1_full_size_forward_600x800():
  """BASELINE config 1: single 600x800 image, VGG-16, forward-only (RPN -> RoI -> NMS correctness) vs the CPU oracle."""
  cfg = dict(backbone = "vgg16", weight_seed = 3, heads = "spread", hw = (600, 800), sample_seed = 3)
  model, oracle, smp = _build(cfg)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  _forward_vs_oracle("config1_forward_600x800_vgg16", model, oracle, smp)


def test_config2_full_size_train_step_600x1000():
  """BASELINE config 2: VGG-16 fwd+bwd at 1000x600, batch 1: losses and every parameter gradient vs the CPU oracle."""
  cfg = dict(backbone = "vgg16", weight_seed = 5, heads = "reference", hw = (600, 1000), sample_seed = 5)
  model, oracle, smp = _build(cfg)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  _train_step_vs_oracle("config2_train_step_600x1000_vgg16", model, oracle, smp, _reference_optimizer(model), seed = 1)
  assert model.last_step_info["num_rois"] == 128


def test_config2_train_steps_with_the_fused_optimizer():
  """The optimizer the benchmark uses (optim.FusedSGD: update + carried fp16 operand split of the new weights in one kernel) end to end:
  TWO steps at 600x1000 against the oracle's SGD, so the second step's GEMMs read the splits the first step's optimizer kernel wrote."""
  from fasterrcnn_b200 import optim
  cfg = dict(backbone = "vgg16", weight_seed = 5, heads = "reference", hw = (600, 1000), sample_seed = 5)
  model, oracle, smp = _build(cfg)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  optimizer = optim.create_optimizer(model, 1e-3, 0.9, 5e-4, fused = True)
  assert isinstance(optimizer, optim.FusedSGD)
  boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  losses = {"ref": [], "gpu": []}
  for who in ("ref", "gpu"):
    random.seed(4); np.random.seed(4); t.manual_seed(4)
    for _ in range(2):
      if who == "ref":
        l = oracle.train_step(smp["image"], smp["anchor_map"], smp["anchor_valid_map"], smp["gt_rpn_map"], smp["gt_rpn_object_indices"],
                              smp["gt_rpn_background_indices"], smp["gt_corners"], smp["gt_class_idxs"])
      else:
        l = model.train_step(optimizer = optimizer, image_data = smp["image"].cuda(), anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                             gt_rpn_map = smp["gt_rpn_map"].cuda(), gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                             gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
      losses[who].append([l.rpn_class, l.rpn_regression, l.detector_class, l.detector_regression, l.total])
  a, b = np.array(losses["gpu"]), np.array(losses["ref"])
  w_err = {k: float((p.detach().cpu() - oracle.params[k].detach()).abs().max()) for k, p in model.named_parameters()}
  worst = max(w_err, key = w_err.get)
  _margins.record("config2_two_steps_fused_sgd", loss_rel_step1 = float(np.max(np.abs(a[0] - b[0]) / np.abs(b[0]))), loss_rel_step2 = float(np.max(np.abs(a[1] - b[1]) / np.abs(b[1]))),
                  post_step_weight_max_abs = w_err[worst], worst_weight = worst)
  np.testing.assert_allclose(a[0], b[0], rtol = LOSS_RTOL_STEP1, atol = 1e-6)
  np.testing.assert_allclose(a[1], b[1], rtol = LOSS_RTOL_STEP2, atol = 1e-6)
  for k, p in model.named_parameters():
    np.testing.assert_allclose(p.detach().cpu().numpy(), oracle.params[k].detach().numpy(), rtol = 1e-4, atol = WEIGHT_ATOL, err_msg = k)


def test_out_of_band_weight_write_needs_invalidate_weight_splits():
  """the fused optimizer carries each weight's fp16 operand split in a side buffer validated by address + torch's
  version counter, which writes through ``p.data`` bypass.  The documented contract: call ops.invalidate_weight_splits() after such a
  write.  After it, the model must behave exactly like a fresh model loaded with the same state dict."""
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import ops, optim
  cfg = gi.E2E_CASES["small"]
  model, _, smp = _build(cfg)
  optimizer = optim.create_optimizer(model, 1e-3, 0.9, 5e-4, fused = True)
  boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  random.seed(0); np.random.seed(0); t.manual_seed(0)
  model.train_step(optimizer = optimizer, image_data = smp["image"].cuda(), anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                   gt_rpn_map = smp["gt_rpn_map"].cuda(), gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                   gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])          # the carried splits now exist
  with t.no_grad():
    for k, p in model.named_parameters():
      if "_block5_conv" in k or "_fc2" in k:
        p.data.mul_(0.5)                                        # out-of-band: no version bump
  ops.invalidate_weight_splits()
  fresh = f.FasterRCNNModel(num_classes = 21, backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0), allow_edge_proposals = True)
  fresh.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()})
  fresh = fresh.cuda()
  model.eval(); fresh.eval()
  with t.no_grad():
    a = model(image_data = smp["image"].cuda())
    b = fresh(image_data = smp["image"].cuda())
  for x, y in zip(a, b):
    assert x.shape == y.shape and t.equal(x, y)


def _build_resnet(kind, hw, seed, bn3_scale):
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import resnet
  from oracle import resnet_oracle
  params = orc.synth_params(resnet_oracle.param_shapes(kind), seed = seed, heads = "spread")
  for k in params:
    if k.endswith("bn3.weight"):
      params[k] = params[k] * bn3_scale                # deep residual stacks with Kaiming-scale branches overflow the synthetic init otherwise
  arch = {"resnet50": resnet.Architecture.ResNet50, "resnet101": resnet.Architecture.ResNet101, "resnet152": resnet.Architecture.ResNet152}[kind]
  model = f.FasterRCNNModel(num_classes = 21, backbone = resnet.ResNetBackbone(arch), allow_edge_proposals = True)
  model.load_state_dict(params)
  return model.cuda(), orc.OracleModel(params, backbone = kind), orc.synthetic_sample(hw, seed = seed, backbone = kind)


def test_config4_resnet101_full_size_600x1000():
  """BASELINE config 4's per-GPU work at its real size: ResNet-101, one 600x1000 image (feature map 1024 x 38 x 63, 21,546 anchors) --
  forward and one train step against the CPU oracle (the 8-GPU half of the config is tests/test_dp_gpu.py + bench.py --backbone resnet101)."""
  model, oracle, smp = _build_resnet("resnet101", (600, 1000), 6, 0.3)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  _forward_vs_oracle("config4_forward_600x1000_resnet101", model, oracle, smp)
  _train_step_vs_oracle("config4_train_step_600x1000_resnet101", model, oracle, smp, _reference_optimizer(model), seed = 2)


def _build_batch(kind, hw, roi_op, proposal_batch_size, weight_seed):
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import resnet
  from oracle import resnet_oracle
  if kind == "vgg16":
    shapes, backbone = orc.vgg16_param_shapes(), f.vgg16.VGG16Backbone(dropout_probability = 0.0)
  else:
    shapes = resnet_oracle.param_shapes(kind)
    backbone = resnet.ResNetBackbone({"resnet50": resnet.Architecture.ResNet50, "resnet101": resnet.Architecture.ResNet101}[kind])
  params = orc.synth_params(shapes, seed = weight_seed, heads = "spread")
  model = f.FasterRCNNModel(num_classes = 21, backbone = backbone, proposal_batch_size = proposal_batch_size, roi_op = roi_op, roi_sampling_ratio = 2, roi_aligned = False)
  model.load_state_dict(params)
  oracle = orc.OracleModel(params, backbone = kind, proposal_batch_size = proposal_batch_size)
  # two different images / ground truths of one size
  smps = [orc.synthetic_sample(hw, seed = 21, backbone = kind),
          orc.synthetic_sample(hw, seed = 22, backbone = kind, gt = [((30.0, 40.0, 200.0, 260.0), 3), ((120.0, 250.0, 330.0, 480.0), 12), ((10.0, 300.0, 150.0, 400.0), 9)])]
  return model.cuda(), oracle, smps


def _match_rows(pg, pr, tol = PX_BAR):
  dist = np.abs(pg[:, None, :] - pr[None, :, :]).max(axis = 2)
  partner = dist.argmin(axis = 1)
  return partner, dist[np.arange(pg.shape[0]), partner] <= tol


@pytest.mark.parametrize("kind,roi_op", [("vgg16", "pool"), ("resnet50", "align")])
def test_batch2_forward_matches_oracle_and_single_image_path(kind, roi_op):
  """EXTENSION, BASELINE config 3 shape (batch 2 per GPU, RoIAlign on ResNet-50) and the RoIPool/VGG-16 twin: forward_batch
  against the oracle's batch restatement, and (RoIPool) against this package's own single-image forward per image."""
  model, oracle, smps = _build_batch(kind, (384, 512), roi_op, 128, weight_seed = 2)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  images = t.cat([s["image"] for s in smps], dim = 0)
  with t.no_grad():
    ref = oracle.forward_batch(images, roi_op = roi_op)
  model.eval()
  with t.no_grad():
    got = model.forward_batch(images.cuda())
    single = [model(image_data = images[b:b + 1].cuda()) for b in range(2)] if roi_op == "pool" else None
  assert len(got) == 2
  for b in range(2):
    pg, cg, dg = [x.cpu().numpy() for x in got[b]]
    pr, cr, dr = [x.numpy() for x in ref[b]]
    assert abs(pg.shape[0] - pr.shape[0]) <= ROWS_BAR
    partner, ok = _match_rows(pg, pr)
    # a proposal within PX_BAR of its partner can still quantise to another RoIPool cell (round(x / 16) at a .5 boundary): such an
    # isolated row gets different pooled features, so rows -- not elements -- are counted
    row_c = np.abs(cg[ok] - cr[partner[ok]]).max(axis = 1)
    row_d = np.abs(dg[ok] - dr[partner[ok]]).max(axis = 1)
    _margins.record("batch2_forward_%s_%s_image%d" % (kind, roi_op, b), rows = int(pg.shape[0]), ref_rows = int(pr.shape[0]), unmatched_rows = int((~ok).sum()),
                    class_score_max_abs = float(row_c.max()), box_delta_max_abs = float(row_d.max()), rows_over_1e4 = int(((row_c > 1e-4) | (row_d > 1e-4)).sum()))
    assert (~ok).sum() <= UNMATCHED_BAR(pg.shape[0]), (b, (~ok).sum())
    assert ((row_c > 1e-4) | (row_d > 1e-4)).sum() <= max(1, int(0.01 * ok.sum())), (b, (row_c > 1e-4).sum(), (row_d > 1e-4).sum())
    if single is not None:                     # same kernels, same weights: the batch path must reproduce the per-image path
      ps, cs, ds = [x.cpu().numpy() for x in single[b]]
      partner, ok = _match_rows(pg, ps, tol = 1e-3)
      assert ok.mean() >= 0.99
      assert (np.abs(cg[ok] - cs[partner[ok]]).max(axis = 1) <= 2e-5).mean() >= 0.99


@pytest.mark.parametrize("kind,roi_op,rois,hw", [("resnet50", "align", 300, (384, 512)), ("vgg16", "pool", 128, (384, 512)), ("resnet50", "align", 300, (600, 1000))])
def test_batch2_train_step_matches_oracle(kind, roi_op, rois, hw):
  """EXTENSION, BASELINE config 3: ResNet-50, batch 2, 300 RoIs per image through RoIAlign -- losses, every parameter gradient and
  the post-step weights against the oracle's batch restatement (same RNG order), at a small size and at the config's own 600x1000
  (feature maps 2 x 1024 x 38 x 63, 600 RoIs through layer4); VGG-16 / RoIPool as the second case."""
  model, oracle, smps = _build_batch(kind, hw, roi_op, rois, weight_seed = 4)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  images = t.cat([s["image"] for s in smps], dim = 0)
  params = [{"params": [p], "weight_decay": 5e-4} for k, p in model.named_parameters() if p.requires_grad and "weight" in k]
  optimizer = t.optim.SGD(params, lr = 1e-3, momentum = 0.9)
  random.seed(3); np.random.seed(3); t.manual_seed(3)
  ref = oracle.train_step_batch(images, smps, roi_op = roi_op)
  ref_grads = {k: v.grad.clone() for k, v in oracle.params.items() if v.grad is not None}
  samples = [dict(anchor_map = s["anchor_map"], anchor_valid_map = s["anchor_valid_map"], gt_rpn_map = s["gt_rpn_map"].cuda(),
                  gt_rpn_object_indices = s["gt_rpn_object_indices"], gt_rpn_background_indices = s["gt_rpn_background_indices"],
                  gt_boxes = [Box(b, c) for b, c in zip(s["gt_corners"], s["gt_class_idxs"])]) for s in smps]
  random.seed(3); np.random.seed(3); t.manual_seed(3)
  got = model.train_step_batch(optimizer, images.cuda(), samples)
  a = np.array([got.rpn_class, got.rpn_regression, got.detector_class, got.detector_regression, got.total])
  b = np.array([ref.rpn_class, ref.rpn_regression, ref.detector_class, ref.detector_regression, ref.total])
  np.testing.assert_allclose(a, b, rtol = LOSS_RTOL_STEP1, atol = 1e-6)
  assert model.last_step_info["rois_per_image"] == oracle.last_batch_rois
  grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
  assert set(grads) == set(ref_grads)
  rels, gm = _margins.grad_margins(grads, ref_grads)
  w_err = max(float((p.detach().cpu() - oracle.params[k].detach()).abs().max()) for k, p in model.named_parameters())
  _margins.record("batch2_train_step_%s_%s_%dx%d" % (kind, roi_op, hw[0], hw[1]), loss_rel = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6))),
                  rois_per_image = model.last_step_info["rois_per_image"], post_step_weight_max_abs = w_err, **gm)
  for k, rel in rels.items():
    assert rel < 2 * GRAD_BAR, (k, rel)
  for k, p in model.named_parameters():
    np.testing.assert_allclose(p.detach().cpu().numpy(), oracle.params[k].detach().numpy(), rtol = 1e-4, atol = WEIGHT_ATOL)


def test_resnet101_forward_and_train_step_match_oracle():
  """BASELINE config 4's backbone at a small size (320x416; the full-size case is test_config4_resnet101_full_size_600x1000): forward
  and one train step against the CPU restatement -- 23 bottlenecks in layer3, frozen BN folded into the filters, stride-2 / 7x7 convs on
  the CUDA-core engine, everything else on the tensor cores."""
  model, oracle, smp = _build_resnet("resnet101", (320, 416), 6, 0.3)
  t.set_num_threads(min(16, os.cpu_count() or 8))
  _forward_vs_oracle("resnet101_forward_320x416", model, oracle, smp)
  _train_step_vs_oracle("resnet101_train_step_320x416", model, oracle, smp, _reference_optimizer(model), seed = 2)


def test_checkpoint_round_trip_and_caffe_partial_load(tmp_path):
  """state.save / state.load (reference state.py:221-288): own-format round trip is bit-exact; a Caffe VGG-16 file initialises
  the 13 convs AND fc1/fc2 (the reference loses the fc layers to a key mismatch) and leaves the heads untouched."""
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import state
  cfg = gi.E2E_CASES["small"]
  model, _, _ = _build(cfg)
  path = str(tmp_path / "ckpt.pth")
  state.save(model, path, epoch = 3)
  fresh = f.FasterRCNNModel(num_classes = 21, backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0)).cuda()
  state.load(fresh, path)
  a, b = model.state_dict(), fresh.state_dict()
  assert list(a.keys()) == list(b.keys())
  for k in a:
    assert t.equal(a[k], b[k]), k
  assert t.load(path)["epoch"] == 3

  caffe = {}
  rng = t.Generator().manual_seed(5)
  for layer, ours in state._CAFFE_LAYERS.items():
    caffe[layer + ".weight"] = t.randn(a[ours + ".weight"].shape, generator = rng)
    caffe[layer + ".bias"] = t.randn(a[ours + ".bias"].shape, generator = rng)
  cpath = str(tmp_path / "vgg16_caffe.pth")
  t.save(caffe, cpath)
  head_before = fresh.state_dict()["_stage2_region_proposal_network._rpn_class.weight"].clone()
  state.load(fresh, cpath)
  sd = fresh.state_dict()
  for layer, ours in state._CAFFE_LAYERS.items():
    assert t.equal(sd[ours + ".weight"].cpu(), caffe[layer + ".weight"]), ours
  assert t.equal(sd["_stage2_region_proposal_network._rpn_class.weight"], head_before)

  tracker = state.BestWeightsTracker(str(tmp_path / "best.pth"))
  tracker.on_epoch_end(model = model, epoch = 1, mAP = 10.0)
  tracker.on_epoch_end(model = fresh, epoch = 2, mAP = 5.0)          # worse: ignored
  tracker.save_best_weights(model)
  best = t.load(str(tmp_path / "best.pth"))
  assert best["epoch"] == 1 and t.equal(best["model_state_dict"]["_stage1_feature_extractor._block1_conv1.weight"], a["_stage1_feature_extractor._block1_conv1.weight"].cpu())


def test_voc_dataset_with_device_anchor_kernels_matches_reference_golden(tmp_path, golden_dir):
  """The data path end to end on the GPU box: fasterrcnn_b200.datasets.voc.Dataset with its default anchor / RPN-target generation
  (frcnn_rpn_decode's anchor generator + frcnn_rpn_targets) reproduces, digest for digest, what the unmodified reference produced
  for the same synthetic VOC tree and seed; one sample then drives a train_step through the reference's call sequence
  (__main__.py:170-184)."""
  import hashlib
  import fasterrcnn_b200 as f
  from fasterrcnn_b200.datasets import voc
  sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
  g = np.load(os.path.join(golden_dir, "voc.npz"))
  d = gi.make_voc_tree(str(tmp_path))
  backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0)
  random.seed(1234)
  ds = voc.Dataset(split = "trainval", image_preprocessing_params = backbone.image_preprocessing_params, compute_feature_map_shape_fn = backbone.compute_feature_map_shape,
                   feature_pixels = backbone.feature_pixels, dir = d, augment = True, shuffle = True, cache = False, prefetch = 2)
  names, rows, samples = [], [], []
  for _ in range(2):
    for smp in ds:
      names.append(os.path.basename(smp.filepath))
      rows.append([sha(smp.image_data), sha(smp.anchor_map), sha(smp.anchor_valid_map), sha(smp.gt_rpn_map),
                   sha(np.asarray(smp.gt_rpn_object_indices, dtype = np.int64)), sha(np.asarray(smp.gt_rpn_background_indices, dtype = np.int64)),
                   sha(np.array([b.corners for b in smp.gt_boxes], dtype = np.float64)), sha(np.array([b.class_index for b in smp.gt_boxes], dtype = np.int64))])
      samples.append(smp)
  assert names == list(g["vgg_names"])
  # every digest but the RPN map's equals the reference's; the map's labels / indices / (ty, tx) are identical too, its (th, tw) = log(gt / anchor) are
  # correctly rounded on the device while NumPy's SIMD float32 log is a <= 1 ulp approximation -- compared against the restatement within 1 ulp
  for got_row, want_row in zip(rows, g["vgg_sha"]):
    assert [x for i, x in enumerate(got_row) if i != 3] == [str(x) for i, x in enumerate(want_row) if i != 3]
  for smp in samples[:4]:
    ref_map, ref_obj, ref_bg = orc.generate_rpn_map(smp.anchor_map, smp.anchor_valid_map, np.array([b.corners for b in smp.gt_boxes], dtype = np.float32))
    assert np.array_equal(smp.gt_rpn_map[..., 0:4], ref_map[..., 0:4])
    np.testing.assert_allclose(smp.gt_rpn_map[..., 4:6], ref_map[..., 4:6], rtol = 2.4e-7, atol = 1e-8)   # 2 ulp: NumPy's float32 log near 1
    assert np.array_equal(smp.gt_rpn_object_indices, ref_obj) and np.array_equal(smp.gt_rpn_background_indices, ref_bg)
  model = f.FasterRCNNModel(num_classes = 21, backbone = backbone).cuda()
  from fasterrcnn_b200 import optim
  optimizer = optim.create_optimizer(model, 1e-3, 0.9, 5e-4)
  smp = samples[0]
  loss = model.train_step(optimizer = optimizer, image_data = t.from_numpy(smp.image_data).unsqueeze(dim = 0).cuda(), anchor_map = smp.anchor_map,
                          anchor_valid_map = smp.anchor_valid_map, gt_rpn_map = t.from_numpy(smp.gt_rpn_map).unsqueeze(dim = 0).cuda(),
                          gt_rpn_object_indices = [smp.gt_rpn_object_indices], gt_rpn_background_indices = [smp.gt_rpn_background_indices], gt_boxes = [smp.gt_boxes])
  assert np.isfinite(loss.total) and loss.total > 0


def test_device_feeder_delivers_every_step_its_own_inputs():
  """datasets.feeder.DeviceFeeder (the end-to-end leg of bench.py): the copy of step i + 1 runs on the copy stream while step i computes;
  every take() must return exactly the tensors submitted for that step, also when the host buffers are rewritten right after submit()
  of the following step (two device slots, two pinned source buffers alternating as a loader would)."""
  from fasterrcnn_b200.datasets.feeder import DeviceFeeder
  dev = t.device("cuda", t.cuda.current_device())
  feeder = DeviceFeeder(dev)
  hosts = [(t.empty((1, 3, 64, 96)).pin_memory(), t.empty((1, 4, 6, 9, 6)).pin_memory()) for _ in range(2)]
  burn = t.randn((2048, 2048), device = dev)
  hosts[0][0].fill_(0.0); hosts[0][1].fill_(100.0)
  feeder.submit(*hosts[0])
  for step in range(6):
    image, gmap = feeder.take()
    nxt = hosts[(step + 1) % 2]
    t.cuda.current_stream().synchronize()                    # (a loader refills a pinned buffer only after its previous copy has been consumed)
    nxt[0].fill_(float(step + 1)); nxt[1].fill_(100.0 + step + 1)
    feeder.submit(*nxt)
    for _ in range(3):
      burn = burn @ burn * 1e-3                              # compute-stream work the next copy overlaps with
    assert float(image.min()) == float(image.max()) == float(step), (step, float(image.min()), float(image.max()))
    assert float(gmap.min()) == float(gmap.max()) == 100.0 + step
  assert DeviceFeeder.bytes_per_step(*hosts[0]) == (3 * 64 * 96 + 4 * 6 * 9 * 6) * 4
