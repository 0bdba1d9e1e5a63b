"""
GPU parity tests: every kernel is called through the C ABI (fasterrcnn_b200.ops -> ctypes ->
libfrcnn_sm100.so) and compared with the CPU oracle on the same seeded inputs.
Bit-exact for index / integer work; floating point within the tolerance written in each test.
"""
import os

import numpy as np
import pytest
import torch as t
import torch.nn.functional as F

from oracle import frcnn_oracle as orc
from oracle import golden_inputs as gi

import _margins

pytestmark = pytest.mark.gpu


@pytest.fixture(scope = "module")
def ops():
  from fasterrcnn_b200 import ops as o
  return o


def _cuda(x):
  return t.from_numpy(np.ascontiguousarray(x)).cuda()


# ---------------------------------------------------------------- conv / linear (K1, K2, K8)
CONV_CASES = [
  # name, N, H, W, Cin, Cout, KH, stride, pad
  ("rgb_stem", 1, 37, 45, 3, 64, 3, 1, 1),
  ("vgg_64", 1, 40, 56, 64, 64, 3, 1, 1),
  ("vgg_512_splitk", 1, 19, 23, 256, 512, 3, 1, 1),
  ("head_1x1_9", 1, 19, 23, 512, 9, 1, 1, 0),
  ("head_1x1_36", 1, 19, 23, 512, 36, 1, 1, 0),
  ("resnet_7x7_s2", 1, 61, 77, 3, 64, 7, 2, 3),
  ("resnet_3x3_s2", 2, 28, 28, 128, 128, 3, 2, 1),
  ("resnet_1x1_s2", 2, 14, 14, 256, 512, 1, 2, 0),
]


def _int_tensor(g, shape, lo, hi):
  return t.randint(lo, hi + 1, shape, generator = g).float()


@pytest.mark.parametrize("case", CONV_CASES, ids = [c[0] for c in CONV_CASES])
@pytest.mark.parametrize("engine", ["simt", "tf32", "auto"])
def test_conv_fwd_dgrad_wgrad_vs_torch_fp32(ops, case, engine):
  """Linear part (no activation): fp32 tolerance = accumulation-order noise only (1e-4 of the output scale)."""
  _, n, h, w, cin, cout, k, stride, pad = case
  ops.set_engine(engine)
  g = t.Generator().manual_seed(7)
  x = t.randn((n, cin, h, w), generator = g)
  wt = t.randn((cout, cin, k, k), generator = g) * (2.0 / (cin * k * k)) ** 0.5
  b = t.randn((cout,), generator = g) * 0.1
  xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
  yr = F.conv2d(xr, wr, br, stride = stride, padding = pad)
  gy = t.randn(yr.shape, generator = g)
  yr.backward(gy)
  xc, wc, bc = x.cuda().requires_grad_(True), wt.cuda().contiguous(memory_format = t.channels_last).requires_grad_(True), b.cuda().requires_grad_(True)
  y = ops.conv2d_act(xc, wc, bc, stride, pad, ops.ACT_NONE)
  assert tuple(y.shape) == tuple(yr.shape)
  np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().numpy(), rtol = 1e-4, atol = 1e-4)
  y.backward(gy.cuda())
  np.testing.assert_allclose(xc.grad.cpu().numpy(), xr.grad.numpy(), rtol = 1e-4, atol = 2e-4)
  scale = float(wr.grad.abs().max())
  np.testing.assert_allclose(wc.grad.cpu().numpy(), wr.grad.numpy(), rtol = 1e-4, atol = 1e-5 * max(scale, 1.0) + 1e-4)
  np.testing.assert_allclose(bc.grad.cpu().numpy(), br.grad.numpy(), rtol = 1e-4, atol = 1e-3)
  ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))


@pytest.mark.parametrize("case", CONV_CASES, ids = [c[0] for c in CONV_CASES])
@pytest.mark.parametrize("engine", ["simt", "tf32", "auto"])
def test_conv_relu_exact_on_integer_data(ops, case, engine):
  """Small-integer operands: every product and partial sum is exact in fp32 (and in the 3xTF32
  split), so the result is independent of the accumulation order -> BIT-EXACT vs torch, including
  the ReLU mask and both gradients."""
  _, n, h, w, cin, cout, k, stride, pad = case
  ops.set_engine(engine)
  g = t.Generator().manual_seed(13)
  x = _int_tensor(g, (n, cin, h, w), -2, 2)
  wt = _int_tensor(g, (cout, cin, k, k), -1, 1)
  b = _int_tensor(g, (cout,), -3, 3)
  xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
  yr = F.relu(F.conv2d(xr, wr, br, stride = stride, padding = pad))
  gy = _int_tensor(g, tuple(yr.shape), -1, 1)
  yr.backward(gy)
  xc, wc, bc = x.cuda().requires_grad_(True), wt.cuda().contiguous(memory_format = t.channels_last).requires_grad_(True), b.cuda().requires_grad_(True)
  y = ops.conv2d_act(xc, wc, bc, stride, pad, ops.ACT_RELU)
  assert np.array_equal(y.detach().cpu().numpy(), yr.detach().numpy())
  y.backward(gy.cuda())
  assert np.array_equal(xc.grad.cpu().numpy(), xr.grad.numpy())
  assert np.array_equal(wc.grad.cpu().numpy(), wr.grad.numpy())
  assert np.array_equal(bc.grad.cpu().numpy(), br.grad.numpy())
  ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))


def test_conv_pool_fused_exact_on_integer_data(ops):
  g = t.Generator().manual_seed(3)
  x = _int_tensor(g, (1, 64, 37, 51), -2, 2)             # odd sizes: floor-mode pooling drops the last row/col
  wt = _int_tensor(g, (128, 64, 3, 3), -1, 1)
  b = _int_tensor(g, (128,), -3, 3)
  xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
  yr = F.max_pool2d(F.relu(F.conv2d(xr, wr, b, padding = 1)), 2, 2)
  gy = _int_tensor(g, tuple(yr.shape), -1, 1)
  yr.backward(gy)
  xc, wc = x.cuda().requires_grad_(True), wt.cuda().requires_grad_(True)
  y = ops.conv2d_act(xc, wc, b.cuda(), 1, 1, ops.ACT_RELU, pool = True)
  assert np.array_equal(y.detach().cpu().numpy(), yr.detach().numpy())
  y.backward(gy.cuda())
  # max-pool ties (equal integers) are broken by the first maximum in (kh,kw) scan order on both sides
  assert np.array_equal(xc.grad.cpu().numpy(), xr.grad.numpy())
  assert np.array_equal(wc.grad.cpu().numpy(), wr.grad.numpy())


@pytest.mark.parametrize("m,k,n", [(128, 25088, 4096), (300, 4096, 4096), (128, 4096, 21), (7, 4096, 80), (0, 4096, 21)])
def test_linear_vs_torch_fp32(ops, m, k, n):
  g = t.Generator().manual_seed(11)
  x = t.randn((m, k), generator = g)
  wt = t.randn((n, k), generator = g) * (1.0 / k) ** 0.5
  b = t.randn((n,), generator = g) * 0.1
  xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
  yr = F.linear(xr, wr, br)
  gy = t.randn(yr.shape, generator = g)
  yr.backward(gy)
  xc, wc, bc = x.cuda().requires_grad_(True), wt.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
  y = ops.linear_act(xc, wc, bc, ops.ACT_NONE)
  np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().numpy(), rtol = 1e-4, atol = 1e-4)
  y.backward(gy.cuda())
  np.testing.assert_allclose(xc.grad.cpu().numpy(), xr.grad.numpy(), rtol = 1e-4, atol = 1e-4)
  np.testing.assert_allclose(wc.grad.cpu().numpy(), wr.grad.numpy(), rtol = 1e-4, atol = 1e-4)
  np.testing.assert_allclose(bc.grad.cpu().numpy(), br.grad.numpy(), rtol = 1e-4, atol = 1e-4)


def test_linear_relu_exact_on_integer_data(ops):
  g = t.Generator().manual_seed(17)
  x = _int_tensor(g, (300, 4096), -2, 2)
  wt = _int_tensor(g, (4096, 4096), -1, 1)
  b = _int_tensor(g, (4096,), -3, 3)
  xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
  yr = F.relu(F.linear(xr, wr, b))
  gy = _int_tensor(g, tuple(yr.shape), -1, 1)
  yr.backward(gy)
  xc, wc = x.cuda().requires_grad_(True), wt.cuda().requires_grad_(True)
  y = ops.linear_act(xc, wc, b.cuda(), ops.ACT_RELU)
  assert np.array_equal(y.detach().cpu().numpy(), yr.detach().numpy())
  y.backward(gy.cuda())
  assert np.array_equal(xc.grad.cpu().numpy(), xr.grad.numpy())
  assert np.array_equal(wc.grad.cpu().numpy(), wr.grad.numpy())


# ---------------------------------------------------------------- anchors + RPN proposal path (K5, K6)
@pytest.mark.parametrize("tag", list(gi.GEOMETRY_CASES))
def test_anchor_generation_bit_exact(ops, tag):
  h, w = gi.GEOMETRY_CASES[tag]
  am, av = orc.generate_anchor_maps((3, h, w), (512, h // 16, w // 16), 16)
  a_dev, v_dev = ops.generate_anchors_device((3, h, w), (h // 16, w // 16), 16)
  assert np.array_equal(a_dev.cpu().numpy(), am)
  assert np.array_equal(v_dev.cpu().numpy(), av)


@pytest.mark.parametrize("tag", list(gi.GEOMETRY_CASES))
def test_rpn_ground_truth_map_on_device(golden_dir, tag):
  """frcnn_rpn_targets (anchors.generate_rpn_map, reference models/anchors.py:137-262) against the restatement pinned on the reference's
  digests: trainable / object flags, (ty, tx) and both index lists bit-exact; (th, tw) = log(gt / anchor) within 2 ulp (correctly rounded
  here, NumPy's SIMD float32 log on the CPU)."""
  from fasterrcnn_b200 import anchors

  class B:
    def __init__(self, c):
      self.corners = np.asarray(c, dtype = np.float32)
  h, w = gi.GEOMETRY_CASES[tag]
  fm = (512, h // 16, w // 16)
  am, av = orc.generate_anchor_maps((3, h, w), fm, 16)
  gt = np.array([b for b, _ in gi.gt_boxes_for(h, w)], dtype = np.float32)
  ref_map, ref_obj, ref_bg = orc.generate_rpn_map(am, av, gt)
  geo = np.load(os.path.join(golden_dir, "geometry.npz"))
  assert np.array_equal(ref_obj.astype(np.int16), geo[tag + "_obj"]) and len(ref_bg) == int(geo[tag + "_nbg"])       # restatement == reference
  am_dev, av_dev = anchors.generate_anchor_maps((3, h, w), fm, 16)
  got_map, got_obj, got_bg = anchors.generate_rpn_map(am_dev, av_dev, [B(c) for c in gt])
  assert np.array_equal(got_map[..., 0:4], ref_map[..., 0:4])
  np.testing.assert_allclose(got_map[..., 4:6], ref_map[..., 4:6], rtol = 2.4e-7, atol = 1e-8)
  assert np.array_equal(got_obj, ref_obj) and np.array_equal(got_bg, ref_bg)


@pytest.mark.parametrize("tag", gi.RPN_CASES)
def test_rpn_proposal_stage_matches_reference_golden(ops, golden_dir, tag):
  g = np.load(os.path.join(golden_dir, "rpn_stage.npz"))
  c = gi.rpn_case(tag)
  am, av = orc.generate_anchor_maps(c["image_shape"], (512,) + c["fm_hw"], 16)
  taps = {}
  ref = orc.rpn_proposals(t.from_numpy(c["score_map"]), t.from_numpy(c["delta_map"]), am, av, c["image_shape"], c["pre_nms"], c["post_nms"], taps = taps)
  assert np.array_equal(ref.numpy(), g[tag + "_proposals"])                       # oracle == reference (pinned)
  props, dbg = ops.rpn_proposals(_cuda(c["score_map"]), _cuda(c["delta_map"]), c["image_shape"], 16, c["pre_nms"], c["post_nms"], return_debug = True)
  # ordering: bit-exact indices (scores are distinct by construction)
  assert np.array_equal(dbg["order"].cpu().numpy().astype(np.int64), taps["order"])
  # decode: <= 1e-4 px on clipped coordinates.  The one transcendental is exp(): correctly rounded on the device, but ATen's CPU exp is
  # a <= 1 ulp approximation whose BITS depend on the host's SIMD dispatch (DEFAULT / AVX2 / AVX512 kernels differ in the last place:
  # ATEN_CPU_CAPABILITY=default vs avx2 changes 2198 -> 2148 of 200000 last-bit disagreements with the correctly rounded value on the
  # same inputs).  That is what round 1 saw as a "host glitch" on one gpurun box.  The bar therefore allows, on top of 1e-4 px, the one
  # ulp of exp() scaled to the box: 2^-23 * the UNCLIPPED side length (fp64 decode of the same deltas), which only matters for boxes
  # larger than ~800 px -- everything else is held to 1e-4 as before.  No retry.
  got, want = dbg["boxes_sorted"].cpu().numpy(), taps["pre_nms_boxes"]
  a64 = am.reshape(-1, 4).astype(np.float64)
  d64 = c["delta_map"].reshape(-1, 4).astype(np.float64)
  side = (a64[:, 2:4] * np.exp(d64[:, 2:4]))[taps["order"]][taps["size_ok"]]            # (rows, [h, w]) of the surviving boxes
  tol = 1e-4 + 2.0 ** -23 * np.concatenate([side, side], axis = 1)                        # columns y1,x1,y2,x2 <- h,w,h,w
  _margins.record("rpn_stage_" + tag, decode_max_abs_px = float(np.abs(got - want).max()), decode_rows = int(got.shape[0]),
                  largest_unclipped_side_px = float(side.max()), rows_over_1e4_px = int((np.abs(got - want) > 1e-4).any(axis = 1).sum()))
  bad = np.argwhere(np.abs(got - want) > tol)
  assert bad.shape[0] == 0, "decode mismatch at (row, col) %s: got %s want %s" % (bad[:4].tolist(), got[bad[:4, 0]].tolist(), want[bad[:4, 0]].tolist())
  assert props.shape == ref.shape
  np.testing.assert_allclose(props.cpu().numpy(), ref.numpy(), rtol = 0, atol = 1e-4)


def test_math_utils_device_helpers_match_the_torch_expressions():
  """math_utils.t_intersection_over_union / t_convert_deltas_to_boxes (one kernel each) against the reference's torch expressions as the
  oracle restates them (frcnn_oracle.iou_t / deltas_to_boxes_t) on CPU: IoU bit for bit (only +, -, *, / with explicit roundings);
  decoded boxes within 1e-4 px + one ulp of exp() scaled to the box (see test_rpn_proposal_stage_matches_reference_golden)."""
  from fasterrcnn_b200 import math_utils
  rng = np.random.default_rng(11)
  b1 = gi.random_boxes(rng, 2000)
  b2 = np.concatenate([gi.random_boxes(rng, 5), b1[:2], np.array([[5, 5, 5, 9], [0, 0, 0, 0]], dtype = np.float32)])
  got = math_utils.t_intersection_over_union(_cuda(b1), _cuda(b2)).cpu().numpy()
  want = orc.iou_t(t.from_numpy(b1), t.from_numpy(b2)).numpy()
  assert np.array_equal(got, want)
  assert math_utils.t_intersection_over_union(_cuda(b1[:0]), _cuda(b2)).shape == (0, 9)
  deltas = rng.normal(0, 0.4, (3000, 4)).astype(np.float32)
  anchors = np.stack([rng.uniform(0, 600, 3000), rng.uniform(0, 1000, 3000), rng.uniform(16, 512, 3000), rng.uniform(16, 512, 3000)], axis = 1).astype(np.float32)
  for means, stds in (([0, 0, 0, 0], [1, 1, 1, 1]), ([0.0, 0.0, 0.0, 0.0], [0.1, 0.1, 0.2, 0.2]), ([0.01, -0.02, 0.03, 0.0], [0.5, 0.5, 0.25, 0.25])):
    got = math_utils.t_convert_deltas_to_boxes(_cuda(deltas), _cuda(anchors), t.tensor(means, dtype = t.float32), t.tensor(stds, dtype = t.float32)).cpu().numpy()
    want = orc.deltas_to_boxes_t(t.from_numpy(deltas), t.from_numpy(anchors), t.tensor(means, dtype = t.float32), t.tensor(stds, dtype = t.float32)).numpy()
    side = anchors[:, 2:4].astype(np.float64) * np.exp(deltas[:, 2:4].astype(np.float64) * np.array(stds[2:4]) + np.array(means[2:4]))
    tol = 1e-4 + 2.0 ** -23 * np.concatenate([side, side], axis = 1)
    assert (np.abs(got - want) <= tol).all(), float(np.abs(got - want).max())


@pytest.mark.parametrize("tag", gi.NMS_CASES)
def test_nms_bit_exact_vs_oracle_and_golden(ops, golden_dir, tag):
  boxes, scores, thr = gi.nms_case(tag)                     # float64 boxes (faster_rcnn.py:216-220) run frcnn_nms_sorted_f64
  gold = np.load(os.path.join(golden_dir, "tv_ops.npz"))["nms_" + tag].astype(np.int64)
  keep = ops.nms(_cuda(boxes), _cuda(scores), thr).cpu().numpy()
  assert np.array_equal(keep, gold)
  assert np.array_equal(keep, orc.nms(boxes, scores, thr))


def test_nms_fp64_standalone_vs_oracle(ops):
  """The per-class call site's dtype (float64 boxes, float32 scores) at the per-class size (300) and at BASELINE config 5's (6000), plus
  boxes whose IoU sits within an fp32 ulp of the threshold: decided differently in float and double, so the fp64 path must not round."""
  rng = np.random.default_rng(7)
  for n, thr in ((300, 0.3), (6000, 0.3), (2000, 0.7)):
    b = gi.random_boxes(rng, n, dtype = np.float64)
    s = rng.permutation(n).astype(np.float32) / n
    keep = ops.nms(_cuda(b), _cuda(s), thr).cpu().numpy()
    assert np.array_equal(keep, orc.nms(b, s, thr)), (n, thr)
  # pairs at IoU = 0.5 exactly and one double-ulp either side of it (never distinguishable in fp32)
  eps = 2.0 ** -40
  b = np.array([[0, 0, 10, 10], [0, 0, 10, 5], [20, 20, 30, 30], [20, 20, 30, 25 + eps * 25], [40, 40, 50, 50], [40, 40, 50, 45 - eps * 45]], dtype = np.float64)
  s = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.4], dtype = np.float32)
  keep = ops.nms(_cuda(b), _cuda(s), 0.5).cpu().numpy()
  assert np.array_equal(keep, orc.nms(b, s, 0.5))
  assert list(keep) == [0, 1, 2, 4, 5]                         # 0.5 is not > 0.5; 0.5 + ulp is; 0.5 - ulp is not


def test_nms_full_size_properties(ops):
  """6000 x 1 and 12000 x 1 at BASELINE sizes: idempotence and the pairwise-IoU invariant."""
  rng = np.random.default_rng(99)
  for n in (6000, 12000):
    b = gi.random_boxes(rng, n)
    s = rng.permutation(n).astype(np.float32) / n
    keep = ops.nms(_cuda(b), _cuda(s), 0.7).cpu().numpy()
    assert np.array_equal(keep, orc.nms(b, s, 0.7))
    kb, ks = b[keep], s[keep]
    again = ops.nms(_cuda(kb), _cuda(ks), 0.7).cpu().numpy()
    assert np.array_equal(again, np.arange(len(keep)))             # idempotent: survivors do not suppress each other
    assert np.all(np.diff(ks) <= 0)                                  # descending scores


def test_nms_batched_matches_per_problem_oracle(ops):
  """Config-5 shape (6000 boxes x 20 classes, thr 0.3) and a small ragged-tie case: the batched entry returns, per problem,
  exactly the single-problem oracle's indices."""
  rng = np.random.default_rng(7)
  for bsz, n, thr, ties in ((20, 6000, 0.3, False), (3, 777, 0.5, True), (2, 64, 0.7, True)):
    b = np.stack([gi.random_boxes(rng, n) for _ in range(bsz)])
    if ties:
      s = np.stack([rng.integers(0, 50, n).astype(np.float32) / 50 for _ in range(bsz)])     # many equal scores
    else:
      s = np.stack([rng.permutation(n).astype(np.float32) / n for _ in range(bsz)])
    keep, counts = ops.nms_batched(_cuda(b), _cuda(s), thr)
    keep, counts = keep.cpu().numpy(), counts.cpu().numpy()
    for z in range(bsz):
      ref = orc.nms(b[z], s[z], thr)
      assert counts[z] == len(ref)
      assert np.array_equal(keep[z, :counts[z]], ref)
      assert np.all(keep[z, counts[z]:] == -1)
  # max_keep cut: first k of the full answer
  keep, counts = ops.nms_batched(_cuda(b), _cuda(s), thr, max_keep = 5)
  for z in range(b.shape[0]):
    ref = orc.nms(b[z], s[z], thr)[:5]
    assert np.array_equal(keep[z, :counts[z]].cpu().numpy(), ref)


def test_roi_pool_scalar_path_when_channels_not_multiple_of_4(ops):
  """C = 6: the one-channel-per-lane kernel (the 128-bit path needs C % 4 == 0)."""
  rng = np.random.default_rng(11)
  fm = rng.normal(0, 1, (1, 6, 12, 17)).astype(np.float32)
  _, rois = gi.roi_case("small")
  out_ref, _ = orc.roi_pool_forward(fm, rois)
  props = np.stack([rois[:, 2], rois[:, 1], rois[:, 4], rois[:, 3]], axis = 1)
  out = ops.roi_pool(_cuda(fm), _cuda(props), (7, 7), 1.0 / 16.0)
  assert np.array_equal(out.cpu().numpy(), out_ref)


# ---------------------------------------------------------------- RoI pooling (K7)
@pytest.mark.parametrize("tag", gi.ROI_CASES)
def test_roi_pool_fwd_bit_exact_bwd_matches(ops, tag):
  fm, rois = gi.roi_case(tag)
  out_ref, arg_ref = orc.roi_pool_forward(fm, rois)
  props = np.stack([rois[:, 2], rois[:, 1], rois[:, 4], rois[:, 3]], axis = 1)    # (b,x1,y1,x2,y2) -> (y1,x1,y2,x2)
  fmc = _cuda(fm).requires_grad_(True)
  out = ops.roi_pool(fmc, _cuda(props), (7, 7), 1.0 / 16.0)
  assert np.array_equal(out.detach().cpu().numpy(), out_ref)
  arg = out.grad_fn.saved_tensors[0].cpu().numpy()                                  # (K, 49, C) bin-major
  assert np.array_equal(arg.transpose(0, 2, 1).reshape(arg_ref.shape), arg_ref)     # same argmax cell (or -1) as the library op, every bin
  go = gi.roi_grad(tag, out_ref.shape)
  out.backward(_cuda(go))
  gin_ref = orc.roi_pool_backward(go, arg_ref, rois, fm.shape)
  np.testing.assert_allclose(fmc.grad.cpu().numpy(), gin_ref, rtol = 0, atol = 1e-5)


def test_roi_pool_microbench_shape_property(ops):
  """6000 RoIs (config 5 shape): every output equals the max over its bin computed by the oracle on a sample."""
  rng = np.random.default_rng(5)
  fm = np.maximum(rng.normal(0, 1, (1, 512, 37, 62)), 0).astype(np.float32)
  b = gi.random_boxes(rng, 6000)
  out = ops.roi_pool(_cuda(fm), _cuda(b), (7, 7), 1.0 / 16.0).cpu().numpy()
  pick = rng.choice(6000, 64, replace = False)
  rois = np.stack([np.zeros(64, np.float32), b[pick, 1], b[pick, 0], b[pick, 3], b[pick, 2]], axis = 1)
  ref, _ = orc.roi_pool_forward(fm, rois)
  assert np.array_equal(out[pick], ref)
  assert out.max() <= fm.max() and out.min() >= 0.0


@pytest.mark.parametrize("aligned", [False, True])
def test_roi_align_extension_vs_oracle(ops, aligned):
  """EXTENSION (no reference counterpart): RoIAlign forward / deterministic backward vs the torchvision-pinned oracle."""
  for tag, s_ratio in (("small", 2), ("small", 3), ("vgg_b128", 2)):
    fm, rois = gi.roi_case(tag)
    props = np.stack([rois[:, 2], rois[:, 1], rois[:, 4], rois[:, 3]], axis = 1)
    ref = orc.roi_align_forward(fm, rois, (7, 7), 1.0 / 16.0, s_ratio, aligned)
    fmc = _cuda(fm).requires_grad_(True)
    out = ops.roi_align(fmc, _cuda(props), (7, 7), 1.0 / 16.0, s_ratio, aligned)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol = 1e-6, atol = 1e-6)
    go = gi.roi_grad(tag, ref.shape)
    out.backward(_cuda(go))
    gin = orc.roi_align_backward(go, rois, fm.shape, 1.0 / 16.0, s_ratio, aligned)
    np.testing.assert_allclose(fmc.grad.cpu().numpy(), gin, rtol = 1e-5, atol = 2e-5)


# ---------------------------------------------------------------- labelling, losses, post-processing
def test_label_proposals_vs_oracle(ops):
  rng = np.random.default_rng(1)
  props = gi.random_boxes(rng, 500)
  gt = np.array([[100, 150, 400, 600], [50, 650, 500, 850], [120, 160, 380, 590]], dtype = np.float32)
  cls = [7, 15, 3]
  allp = np.vstack([props, gt])
  p_ref, onehot_ref, packed_ref = orc.label_proposals(t.from_numpy(props), gt, cls)
  best, cidx, onehot, packed = ops.label_proposals(_cuda(allp), _cuda(gt), t.tensor(cls, dtype = t.int32).cuda(), 21)
  assert np.array_equal(onehot.cpu().numpy(), onehot_ref.numpy())                  # labels bit-exact
  np.testing.assert_array_equal(packed[:, 0, :].cpu().numpy(), packed_ref[:, 0, :].numpy())
  np.testing.assert_allclose(packed[:, 1, :].cpu().numpy(), packed_ref[:, 1, :].numpy(), rtol = 1e-5, atol = 1e-5)


def test_rpn_and_detector_losses_and_grads_vs_oracle(ops):
  smp = orc.synthetic_sample((384, 512), seed = 0)
  fh, fw = 24, 32
  g = t.Generator().manual_seed(5)
  scores = t.rand((1, fh, fw, 9), generator = g) * 0.98 + 0.01
  deltas = t.randn((1, fh, fw, 36), generator = g) * 0.5
  import random
  random.seed(0)
  mb = orc.sample_rpn_minibatch(smp["gt_rpn_map"], smp["gt_rpn_object_indices"], smp["gt_rpn_background_indices"])
  sr, dr = scores.clone().requires_grad_(True), deltas.clone().requires_grad_(True)
  l1, l2 = orc.rpn_class_loss(sr, mb), orc.rpn_regression_loss(dr, mb)
  (l1 + l2).backward()
  sc, dc = scores.cuda().requires_grad_(True), deltas.cuda().requires_grad_(True)
  out = ops.rpn_losses(sc, dc, mb.cuda())
  out.sum().backward()
  np.testing.assert_allclose(out.detach().cpu().numpy(), [l1.item(), l2.item()], rtol = 1e-5)
  np.testing.assert_allclose(sc.grad.cpu().numpy(), sr.grad.numpy(), rtol = 1e-4, atol = 1e-8)
  np.testing.assert_allclose(dc.grad.cpu().numpy(), dr.grad.numpy(), rtol = 1e-4, atol = 1e-8)

  n, c = 128, 21
  logits = t.randn((n, c), generator = g)
  dl = t.randn((n, 80), generator = g) * 0.5
  yc = F.one_hot(t.randint(0, c, (n,), generator = g), c).float()
  yd = t.zeros((n, 2, 80))
  yd[:, 1, :] = t.randn((n, 80), generator = g)
  yd[:, 0, :] = t.repeat_interleave(yc, 4, dim = 1)[:, 4:]
  lr_, dlr = logits.clone().requires_grad_(True), dl.clone().requires_grad_(True)
  pr = F.softmax(lr_, dim = 1)
  k1, k2 = orc.detector_class_loss(pr, yc), orc.detector_regression_loss(dlr, yd)
  (k1 + k2).backward()
  lc, dlc = logits.cuda().requires_grad_(True), dl.cuda().requires_grad_(True)
  pc = ops.softmax_rows(lc)
  np.testing.assert_allclose(pc.detach().cpu().numpy(), pr.detach().numpy(), rtol = 1e-5, atol = 1e-7)
  out = ops.detector_losses(pc, dlc, yc.cuda(), yd.cuda())
  out.sum().backward()
  np.testing.assert_allclose(out.detach().cpu().numpy(), [k1.item(), k2.item()], rtol = 1e-5)
  np.testing.assert_allclose(lc.grad.cpu().numpy(), lr_.grad.numpy(), rtol = 1e-4, atol = 1e-7)
  np.testing.assert_allclose(dlc.grad.cpu().numpy(), dlr.grad.numpy(), rtol = 1e-4, atol = 1e-8)


def test_detect_postprocess_vs_oracle(ops):
  rng = np.random.default_rng(21)
  n = 300
  props = gi.random_boxes(rng, n, 600.0, 800.0)
  classes = rng.dirichlet(np.ones(21) * 0.3, n).astype(np.float32)
  deltas = rng.normal(0, 0.5, (n, 80)).astype(np.float32)
  got = ops.detect_postprocess(_cuda(props), _cuda(classes), _cuda(deltas), (600, 800), 0.05, 0.3)
  pa = np.empty(props.shape)
  pa[:, 0] = 0.5 * (props[:, 0] + props[:, 2]); pa[:, 1] = 0.5 * (props[:, 1] + props[:, 3]); pa[:, 2:4] = props[:, 2:4] - props[:, 0:2]
  for c in range(1, 21):
    boxes = orc.deltas_to_boxes_np(deltas[:, (c - 1) * 4:(c - 1) * 4 + 4], pa, [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2])
    boxes[:, 0::2] = np.clip(boxes[:, 0::2], 0, 599); boxes[:, 1::2] = np.clip(boxes[:, 1::2], 0, 799)
    sel = np.where(classes[:, c] > 0.05)[0]
    keep = orc.nms(boxes[sel], classes[sel, c], 0.3)
    ref = np.hstack([boxes[sel][keep], classes[sel, c][keep][:, None]])
    assert got[c].shape == ref.shape, c
    np.testing.assert_allclose(got[c], ref, rtol = 0, atol = 1e-9)


def test_sgd_kernel_vs_oracle_update_rule(ops):
  g = t.Generator().manual_seed(2)
  p = t.randn((1000003,), generator = g); gr = t.randn((1000003,), generator = g)
  pr, buf_r = p.clone(), None
  pc, bc = p.cuda(), t.zeros_like(p).cuda()
  for step in range(3):
    gg = gr + 5e-4 * pr
    buf_r = gg.clone() if buf_r is None else buf_r * 0.9 + gg
    pr = pr - 1e-3 * buf_r
    ops.sgd_step(pc, gr.cuda(), bc, 1e-3, 0.9, 5e-4, 1.0, first_step = (step == 0))
  np.testing.assert_allclose(pc.cpu().numpy(), pr.numpy(), rtol = 1e-6, atol = 1e-7)


# ---------------------------------------------------------------- tcgen05 engine vs the exact fp32 engine
TC_CASES = [
  ("vgg_b5_512_37x62", 1, 37, 62, 512, 512, 3, 1),
  ("vgg_b3_256_150x250", 1, 150, 250, 128, 256, 3, 1),
  ("vgg_b1_64_600x1000", 1, 600, 1000, 64, 64, 3, 1),
  ("rpn_1x1_like_512", 1, 37, 62, 512, 128, 1, 0),
]


@pytest.mark.parametrize("case", TC_CASES, ids = [c[0] for c in TC_CASES])
def test_tcgen05_conv_fwd_matches_fp32_engine(ops, case):
  """3xTF32 on the tensor cores vs exact-fp32 CUDA cores at the BASELINE layer shapes: <= 2e-5 of the output scale."""
  _, n, h, w, cin, cout, k, pad = case
  g = t.Generator().manual_seed(23)
  x = ops.as_nhwc(t.randn((n, cin, h, w), generator = g).cuda())
  wt = (t.randn((cout, cin, k, k), generator = g) * (2.0 / (cin * k * k)) ** 0.5).cuda().contiguous(memory_format = t.channels_last)
  b = (t.randn((cout,), generator = g) * 0.1).cuda()
  ops.set_engine("simt")
  y_ref = ops.conv2d_fwd_raw(x, wt, b, 1, pad, ops.ACT_RELU)
  ops.set_engine("tc")
  y = ops.conv2d_fwd_raw(x, wt, b, 1, pad, ops.ACT_RELU)
  ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))
  scale = float(y_ref.abs().max())
  err = float((y - y_ref).abs().max())
  assert err <= 2e-5 * scale, (err, scale)


def test_tcgen05_linear_fwd_splitk_matches_fp32_engine(ops):
  g = t.Generator().manual_seed(29)
  x = t.randn((128, 25088), generator = g).cuda()
  wt = (t.randn((4096, 25088), generator = g) * (1.0 / 25088) ** 0.5).cuda()
  b = (t.randn((4096,), generator = g) * 0.1).cuda()
  # K = 25088: the fp32 accumulation error of EITHER engine is ~1e-4, so both are judged against fp64
  y64 = t.relu(x.double().cpu() @ wt.double().cpu().t() + b.double().cpu())
  ops.set_engine("simt")
  y_simt = ops.linear_act(x, wt, b, ops.ACT_RELU)
  ops.set_engine("tc")
  y = ops.linear_act(x, wt, b, ops.ACT_RELU)
  ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))
  scale = float(y64.abs().max())
  err_tc = float((y.double().cpu() - y64).abs().max())
  err_simt = float((y_simt.double().cpu() - y64).abs().max())
  assert err_simt <= 1e-5 * scale, (err_simt, scale)
  assert err_tc <= 1e-5 * scale, (err_tc, err_simt, scale)          # 3xTF32 + out-of-core accumulation: fp32-grade


TC_BWD_CASES = [
  ("vgg_b5_512_37x62", 1, 37, 62, 512, 512, 3, 1),
  ("vgg_b3_128to256_150x250", 1, 150, 250, 128, 256, 3, 1),
  ("vgg_b4_256to512_75x125", 1, 75, 125, 256, 512, 3, 1),
  ("pointwise_512to128_37x62", 1, 37, 62, 512, 128, 1, 0),
]


@pytest.mark.parametrize("case", TC_BWD_CASES, ids = [c[0] for c in TC_BWD_CASES])
def test_tcgen05_dgrad_wgrad_match_fp32_engine(ops, case):
  """Data / filter gradients on the tensor cores (MN-major operand descriptors) vs the exact-fp32 engine."""
  _, n, h, w, cin, cout, k, pad = case
  g = t.Generator().manual_seed(31)
  x = ops.as_nhwc(t.randn((n, cin, h, w), generator = g).cuda())
  dy = ops.as_nhwc(t.randn((n, cout, h, w), generator = g).cuda())
  wt = (t.randn((cout, cin, k, k), generator = g) * (2.0 / (cin * k * k)) ** 0.5).cuda().contiguous(memory_format = t.channels_last)
  ops.set_engine("simt")
  dx_ref = ops.conv2d_dgrad_raw(dy, wt, (n, cin, h, w), 1, pad)
  dw_ref = ops.conv2d_wgrad_raw(dy, x, (cout, cin, k, k), 1, pad)
  ops.set_engine("tc")
  dx = ops.conv2d_dgrad_raw(dy, wt, (n, cin, h, w), 1, pad)
  dw = ops.conv2d_wgrad_raw(dy, x, (cout, cin, k, k), 1, pad)
  ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))
  assert float((dx - dx_ref).abs().max()) <= 2e-5 * float(dx_ref.abs().max())
  assert float((dw - dw_ref).abs().max()) <= 5e-5 * float(dw_ref.abs().max())


def test_tcgen05_linear_backward_matches_fp64(ops):
  g = t.Generator().manual_seed(37)
  x = t.randn((128, 25088), generator = g)
  wt = t.randn((4096, 25088), generator = g) * (1.0 / 25088) ** 0.5
  gy = t.randn((128, 4096), generator = g)
  dx64 = gy.double() @ wt.double()
  dw64 = gy.double().t() @ x.double()
  ops.set_engine("tc")
  xc, wc = x.cuda().requires_grad_(True), wt.cuda().requires_grad_(True)
  y = ops.linear_act(xc, wc, None, ops.ACT_NONE)
  y.backward(gy.cuda())
  ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))
  assert float((xc.grad.double().cpu() - dx64).abs().max()) <= 1e-4 * float(dx64.abs().max())
  assert float((wc.grad.double().cpu() - dw64).abs().max()) <= 1e-4 * float(dw64.abs().max())


# ---------------------------------------------------------------- fp16 tcgen05 engine (FRCNN_ENGINE_TC_3XF16)
def test_f16_split_reconstructs_fp32(ops):
  """x * 2^e = hi + lo / 2048 with e from the tensor's amax: max |hi| in [2^13, 2^14); reconstruction error <= 2^-22 of |x| for
  elements within 2^-27 of the maximum (both halves in fp16's normal range), and <= 2^-36 * 2^-e absolutely below that."""
  from fasterrcnn_b200._lib import lib, ptr, check, stream
  L = lib()
  g = t.Generator().manual_seed(41)
  for n, scale in ((1000003, 1.0), (4096 * 64, 3.7e-6), (12345, 8.1e4)):
    x = (t.randn((n,), generator = g) * scale).cuda()
    x[::7] *= 1e-6                                                   # wide dynamic range
    buf = t.empty((L.frcnn_f16_split_bytes(n),), dtype = t.uint8, device = "cuda")
    check(L.frcnn_f16_split(ptr(x), n, ptr(buf), stream()), "frcnn_f16_split")
    hdr = buf[:8].view(t.int32).cpu()
    assert int(hdr[0]) == int(x.abs().max().view(t.int32))
    e = int(hdr[1])
    half_bytes = (2 * n + 1023) // 1024 * 1024
    hi = buf[4096:4096 + 2 * n].view(t.float16).double()
    lo = buf[4096 + half_bytes:4096 + half_bytes + 2 * n].view(t.float16).double()
    assert 2.0 ** 13 <= float(hi.abs().max()) <= 2.0 ** 14
    rec = (hi + lo / 2048.0) * 2.0 ** (-e)
    err = (rec - x.double()).abs()
    bound = t.maximum(x.double().abs() * 2.0 ** -21, t.full_like(err, 2.0 ** -35 * 2.0 ** (-e)))
    assert bool((err <= bound).all()), float((err / bound).max())


F16_CASES = [
  ("vgg_b5_512_37x62", 1, 37, 62, 512, 512, 3, 1),
  ("vgg_b3_128to256_150x250", 1, 150, 250, 128, 256, 3, 1),
  ("vgg_b2_64to128_60x100", 1, 60, 100, 64, 128, 3, 1),
  ("pointwise_512to128_37x62", 1, 37, 62, 512, 128, 1, 0),
  ("batch2_256_19x23", 2, 19, 23, 256, 256, 3, 1),
]


@pytest.mark.parametrize("case", F16_CASES, ids = [c[0] for c in F16_CASES])
def test_f16_engine_fwd_dgrad_wgrad_match_fp32_engine(ops, case):
  """The fp16 tensor-core engine (scaled hi/lo fp16 splits, three kind::f16 products) against the exact-fp32 CUDA-core engine:
  same bars as the tf32 engine (2e-5 / 5e-5 of the output scale); operands deliberately far from unit scale."""
  name, n, h, w, cin, cout, k, pad = case
  g = t.Generator().manual_seed(43)
  x = ops.as_nhwc((t.randn((n, cin, h, w), generator = g) * 37.0).cuda())
  dy = ops.as_nhwc((t.randn((n, cout, h, w), generator = g) * 2.5e-4).cuda())
  wt = (t.randn((cout, cin, k, k), generator = g) * (2.0 / (cin * k * k)) ** 0.5).cuda().contiguous(memory_format = t.channels_last)
  b = (t.randn((cout,), generator = g) * 0.1).cuda()
  ops.set_engine("simt")
  y_ref = ops.conv2d_fwd_raw(x, wt, b, 1, pad, ops.ACT_RELU)
  dx_ref = ops.conv2d_dgrad_raw(dy, wt, (n, cin, h, w), 1, pad)
  dw_ref = ops.conv2d_wgrad_raw(dy, x, (cout, cin, k, k), 1, pad)
  ops.set_engine("f16")
  try:
    assert ops._uses_tc(0, (n, h, w, cin, cout, k, k, 1, pad)) and ops._uses_tc(1, (n, h, w, cin, cout, k, k, 1, pad)) and ops._uses_tc(2, (n, h, w, cin, cout, k, k, 1, pad))
    y = ops.conv2d_fwd_raw(x, wt, b, 1, pad, ops.ACT_RELU)
    dx = ops.conv2d_dgrad_raw(dy, wt, (n, cin, h, w), 1, pad)
    dw = ops.conv2d_wgrad_raw(dy, x, (cout, cin, k, k), 1, pad)
    # the epilogue leaves the per-CTA maxima of |output| behind: their maximum IS the tensor's, and the split built from them
    # equals the split built by the amax pass
    hints = 0
    for out in (y, dx):
      hnt = ops._amax_hint(out)
      if hnt is not None:
        hints += 1
        assert int(hnt[0][16:16 + hnt[1]].max()) == int(out.abs().max().view(t.int32))
        a = ops.tf32_split(out, cache = False)
        ops._amax_hints.clear()
        bfull = ops.tf32_split(out, cache = False)
        nb, half = 2 * out.numel(), (2 * out.numel() + 1023) // 1024 * 1024          # (padding bytes behind each half are not written)
        assert t.equal(a[4:8], bfull[4:8]) and t.equal(a[4096:4096 + nb], bfull[4096:4096 + nb]) and t.equal(a[4096 + half:4096 + half + nb], bfull[4096 + half:4096 + half + nb])
    assert hints >= 1 or name.startswith("batch2")
  finally:
    ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))
  assert float((y - y_ref).abs().max()) <= 2e-5 * float(y_ref.abs().max())
  assert float((dx - dx_ref).abs().max()) <= 2e-5 * float(dx_ref.abs().max())
  assert float((dw - dw_ref).abs().max()) <= 5e-5 * float(dw_ref.abs().max())


def test_f16_engine_exact_on_integer_data_and_long_k(ops):
  """Small integers: every fp16 product and fp32 partial sum is exact -> bit-exact vs torch through conv + ReLU and both
  gradients; K = 25088 linear layer (fc1) forward / backward against fp64 at the tf32 engine's bar."""
  ops.set_engine("f16")
  try:
    g = t.Generator().manual_seed(47)
    x = _int_tensor(g, (1, 64, 40, 56), -2, 2)
    wt = _int_tensor(g, (128, 64, 3, 3), -1, 1)
    b = _int_tensor(g, (128,), -3, 3)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wr, br, padding = 1))
    gy = _int_tensor(g, tuple(yr.shape), -1, 1)
    yr.backward(gy)
    xc, wc, bc = x.cuda().requires_grad_(True), wt.cuda().contiguous(memory_format = t.channels_last).requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.conv2d_act(xc, wc, bc, 1, 1, ops.ACT_RELU)
    assert np.array_equal(y.detach().cpu().numpy(), yr.detach().numpy())
    y.backward(gy.cuda())
    assert np.array_equal(xc.grad.cpu().numpy(), xr.grad.numpy())
    assert np.array_equal(wc.grad.cpu().numpy(), wr.grad.numpy())
    assert np.array_equal(bc.grad.cpu().numpy(), br.grad.numpy())

    x = t.randn((128, 25088), generator = g)
    wt = t.randn((4096, 25088), generator = g) * (1.0 / 25088) ** 0.5
    gy = t.randn((128, 4096), generator = g)
    y64 = x.double() @ wt.double().t()
    dx64 = gy.double() @ wt.double()
    dw64 = gy.double().t() @ x.double()
    xc, wc = x.cuda().requires_grad_(True), wt.cuda().requires_grad_(True)
    y = ops.linear_act(xc, wc, None, ops.ACT_NONE)
    y.backward(gy.cuda())
  finally:
    ops.set_engine(os.environ.get("FRCNN_ENGINE", "auto"))
  assert float((y.detach().double().cpu() - y64).abs().max()) <= 1e-5 * float(y64.abs().max())
  assert float((xc.grad.double().cpu() - dx64).abs().max()) <= 1e-4 * float(dx64.abs().max())
  assert float((wc.grad.double().cpu() - dw64).abs().max()) <= 1e-4 * float(dw64.abs().max())


@pytest.mark.parametrize("rows,c,act", [(37 * 45, 256, "relu"), (128, 4096, "relu"), (1000, 64, "none")])
def test_f16_fused_split_producers_match_the_plain_split(ops, rows, c, act):
  """frcnn_act_bwd_fused_f16 (exponent from max |dy|) and frcnn_sgd_step_split_f16 (exponent kept from the seeding split) write
  operand splits that reconstruct dz / the updated weights to the format's accuracy."""
  from fasterrcnn_b200._lib import lib, ptr, check, stream, workspace
  L = lib()
  g = t.Generator().manual_seed(rows * 3 + c)
  n = rows * c
  half_bytes = (2 * n + 1023) // 1024 * 1024

  def unpack(buf):
    e = int(buf[4:8].view(t.int32).cpu())
    hi = buf[4096:4096 + 2 * n].view(t.float16).double()
    lo = buf[4096 + half_bytes:4096 + half_bytes + 2 * n].view(t.float16).double()
    return (hi + lo / 2048.0) * 2.0 ** (-e), e, float(hi.abs().max())

  dy = (t.randn((rows, c), generator = g) * 3e-3).cuda()
  y = t.randn((rows, c), generator = g).cuda()
  code = ops.ACT_RELU if act == "relu" else ops.ACT_NONE
  split = t.empty((L.frcnn_f16_split_bytes(n),), dtype = t.uint8, device = "cuda")
  db = t.empty((c,), dtype = t.float32, device = "cuda")
  ws, ws_n = workspace(L.frcnn_act_bwd_fused_workspace_bytes(rows, c), slot = 1)
  check(L.frcnn_act_bwd_fused_f16(ptr(dy), ptr(y), code, None, ptr(split), ptr(db), rows, c, None, 0, ws, ws_n, stream()), "frcnn_act_bwd_fused_f16")
  want = (t.where(y > 0, dy, t.zeros_like(dy)) if act == "relu" else dy).double().reshape(-1)
  rec, e, hmax = unpack(split)
  assert hmax < 2.0 ** 14 and float(dy.abs().max()) * 2.0 ** e >= 2.0 ** 13
  assert float((rec - want).abs().max()) <= 2.0 ** -21 * float(dy.abs().max())
  assert float((db.double() - want.reshape(rows, c).sum(0)).abs().max()) <= 1e-5 * float(want.abs().reshape(rows, c).sum(0).max())

  p = (t.randn((n,), generator = g) * 0.02).cuda(); gr = t.randn((n,), generator = g).cuda(); buf = t.zeros_like(p)
  ps = t.empty((L.frcnn_f16_split_bytes(n),), dtype = t.uint8, device = "cuda")
  check(L.frcnn_f16_split(ptr(p), n, ptr(ps), stream()), "frcnn_f16_split")
  p_ref = p.clone(); buf_ref = buf.clone()
  check(L.frcnn_sgd_step(ptr(p_ref), ptr(gr), ptr(buf_ref), n, 1e-3, 0.9, 5e-4, 1.0, 1, stream()), "frcnn_sgd_step")
  check(L.frcnn_sgd_step_split_f16(ptr(p), ptr(gr), ptr(buf), n, 1e-3, 0.9, 5e-4, 1.0, 1, ptr(ps), stream()), "frcnn_sgd_step_split_f16")
  assert t.equal(p, p_ref) and t.equal(buf, buf_ref)
  rec, e, hmax = unpack(ps)
  assert float((rec - p.double()).abs().max()) <= 2.0 ** -21 * float(p.abs().max())


@pytest.mark.parametrize("rows,c,act", [(37 * 45, 256, "relu"), (2294, 512, "relu"), (128, 4096, "relu"), (1000, 64, "none"), (77, 2048, "none")])
def test_act_bwd_fused_matches_separate_passes(ops, rows, c, act):
  """frcnn_act_bwd_fused = relu-backward + tf32 operand split + bias row-sum in one pass: dz bit-exact, hi + lo == dz exactly with
  hi tf32-representable (10 explicit mantissa bits), bias gradient within fp32 summation-order noise; bit-exact on integers."""
  from fasterrcnn_b200._lib import lib, ptr, check, stream, workspace
  L = lib()
  assert L.frcnn_act_bwd_fused_supported(rows, c) == 1
  assert L.frcnn_act_bwd_fused_supported(rows, 21) == 0 and L.frcnn_act_bwd_fused_supported(rows, 96) == 0
  g = t.Generator().manual_seed(rows + c)
  for integer in (False, True):
    dy = (_int_tensor(g, (rows, c), -3, 3) if integer else t.randn((rows, c), generator = g)).cuda()
    y = t.randn((rows, c), generator = g).cuda()
    code = ops.ACT_RELU if act == "relu" else ops.ACT_NONE
    dz = t.empty_like(dy)
    split = t.empty((L.frcnn_tf32_split_bytes(dy.numel()),), dtype = t.uint8, device = "cuda")
    db = t.empty((c,), dtype = t.float32, device = "cuda")
    ws, ws_n = workspace(L.frcnn_act_bwd_fused_workspace_bytes(rows, c), slot = 1)
    check(L.frcnn_act_bwd_fused(ptr(dy), ptr(y), code, ptr(dz), ptr(split), ptr(db), rows, c, ws, ws_n, stream()), "frcnn_act_bwd_fused")
    want = t.where(y > 0, dy, t.zeros_like(dy)) if act == "relu" else dy
    assert t.equal(dz, want)
    lo_off = (dy.numel() * 4 + 1023) // 1024 * 1024
    hi = split[:dy.numel() * 4].view(t.float32).view(rows, c)
    lo = split[lo_off:lo_off + dy.numel() * 4].view(t.float32).view(rows, c)
    assert t.equal(hi + lo, want)
    assert int((hi.view(t.int32) & 0x1FFF).abs().max()) == 0              # low 13 mantissa bits clear: exactly a tf32 value
    ref = want.double().sum(0)
    if integer:
      assert t.equal(db.double(), ref)
    else:
      assert float((db.double() - ref).abs().max()) <= 1e-5 * float(want.abs().sum(0).max())
    # split-only and bias-only variants leave the other outputs untouched
    db2 = t.full((c,), 7.0, device = "cuda")
    check(L.frcnn_act_bwd_fused(ptr(dy), ptr(y), code, None, ptr(split), None, rows, c, None, 0, stream()), "frcnn_act_bwd_fused")
    assert t.equal(split[:dy.numel() * 4].view(t.float32).view(rows, c), hi) and float(db2[0]) == 7.0


@pytest.mark.parametrize("m,k,n1,n2,act1", [(23 * 37, 512, 9, 36, "sigmoid"), (128, 4096, 21, 80, "none"), (5, 2048, 21, 80, "none"), (300, 4096, 21, 80, "none")])
def test_two_heads_vs_torch(ops, m, k, n1, n2, act1):
  """frcnn_heads_fwd / _bwd (the RPN's two 1x1 convs, the detector's two linears) against torch fp32: forward, dx, dw, db within
  summation-order noise on random data, bit-exact on small integers (linear heads)."""
  g = t.Generator().manual_seed(m + k)
  code = ops.ACT_SIGMOID if act1 == "sigmoid" else ops.ACT_NONE
  for integer in (False, True):
    if integer:
      x = _int_tensor(g, (m, k), -2, 2); w1 = _int_tensor(g, (n1, k), -1, 1); w2 = _int_tensor(g, (n2, k), -1, 1)
      b1 = _int_tensor(g, (n1,), -3, 3); b2 = _int_tensor(g, (n2,), -3, 3)
      g1 = _int_tensor(g, (m, n1), -1, 1); g2 = _int_tensor(g, (m, n2), -1, 1)
    else:
      x = t.randn((m, k), generator = g); w1 = t.randn((n1, k), generator = g) * k ** -0.5; w2 = t.randn((n2, k), generator = g) * k ** -0.5
      b1 = t.randn((n1,), generator = g) * 0.1; b2 = t.randn((n2,), generator = g) * 0.1
      g1 = t.randn((m, n1), generator = g); g2 = t.randn((m, n2), generator = g)
    ref = [v.clone().requires_grad_(True) for v in (x, w1, b1, w2, b2)]
    r1 = F.linear(ref[0], ref[1], ref[2]); r2 = F.linear(ref[0], ref[3], ref[4])
    if act1 == "sigmoid":
      r1 = t.sigmoid(r1)
    (r1 * g1).sum().backward(retain_graph = True); (r2 * g2).sum().backward()
    dev = [v.cuda().requires_grad_(True) for v in (x, w1, b1, w2, b2)]
    y1, y2 = ops.two_heads(dev[0], dev[1], dev[2], code, dev[3], dev[4], ops.ACT_NONE)
    t.autograd.backward([y1, y2], [g1.cuda(), g2.cuda()])
    exact = integer and act1 != "sigmoid"
    pairs = [(y1, r1), (y2, r2)] + [(d.grad, r.grad) for d, r in zip(dev, ref)]
    for got, want in pairs:
      got, want = got.detach().cpu(), want.detach()
      if exact:
        assert t.equal(got, want)
      else:
        scale = max(float(want.abs().max()), 1e-6)
        assert float((got - want).abs().max()) <= 2e-5 * scale + 1e-6


# ---------------------------------------------------------------- stride-2 convolutions of the ResNet bottlenecks on the tensor cores
@pytest.mark.parametrize("n,h,w,c", [(1, 38, 63, 128), (3, 7, 7, 64), (2, 9, 12, 6), (1, 1, 1, 4)])
def test_subsample2_upsample2_zero_bit_exact(ops, n, h, w, c):
  """frcnn_subsample2 == x[:, :, ::2, ::2]; frcnn_upsample2_zero == its adjoint (odd extents, vector and scalar channel counts)."""
  from fasterrcnn_b200 import resnet
  g = t.Generator().manual_seed(5)
  x = t.randn((n, c, h, w), generator = g).cuda().contiguous(memory_format = t.channels_last)
  sub = resnet._subsample2(x)
  assert sub.is_contiguous(memory_format = t.channels_last) or sub.numel() == sub.shape[1]
  assert t.equal(sub, x[:, :, ::2, ::2])
  up = resnet._upsample2_zero(sub, h, w)
  want = t.zeros_like(x)
  want[:, :, ::2, ::2] = x[:, :, ::2, ::2]
  assert t.equal(up, want)


S2_CASES = [
  # k, n, h, w, cin, cout     (layer3.0 at 600x1000 scale, layer4.0 on 7x7 RoIs, odd maps)
  (1, 1, 38, 63, 128, 256),
  (3, 1, 38, 63, 128, 128),
  (3, 6, 7, 7, 128, 128),
  (1, 6, 7, 7, 128, 256),
  (3, 1, 37, 61, 64, 64),
]


@pytest.mark.parametrize("case", S2_CASES, ids = ["%dx%d_n%d_%dx%d_%d_%d" % (c[0], c[0], c[1], c[2], c[3], c[4], c[5]) for c in S2_CASES])
def test_stride2_conv_bn_relu_vs_torch_fp64(ops, case, monkeypatch):
  """resnet.conv_bn_act(stride = 2): the tensor-core forms (1x1 on the subsampled input; 3x3 at full resolution, subsampled; backward
  through the zero-upsampled gradient) against torch's strided conv2d in fp64 -- output, dx, dw within 1e-5 of each tensor's scale --
  and against the CUDA-core strided kernels they replace (FRCNN_RESNET_S2_TC=0) at the same bar."""
  from fasterrcnn_b200 import resnet
  k, n, h, w, cin, cout = case
  pad = 1 if k == 3 else 0
  g = t.Generator().manual_seed(11)
  x = t.randn((n, cin, h, w), generator = g)
  wt = t.randn((cout, cin, k, k), generator = g) * (2.0 / (cin * k * k)) ** 0.5
  scale = t.rand((cout,), generator = g) + 0.5
  shift = t.randn((cout,), generator = g) * 0.1
  xr, wr = x.double().requires_grad_(True), wt.double().requires_grad_(True)
  yr = F.relu(F.conv2d(xr, wr * scale.double()[:, None, None, None], shift.double(), stride = 2, padding = pad))
  gy = t.randn(yr.shape, generator = g)
  yr.backward(gy.double())

  def run():
    xc = x.cuda().contiguous(memory_format = t.channels_last).requires_grad_(True)
    wc = wt.cuda().contiguous(memory_format = t.channels_last).requires_grad_(True)
    ops.begin_step()
    y = resnet._ConvBNAct.apply(xc, wc, scale.cuda(), shift.cuda(), None, 2, pad, ops.ACT_RELU)
    y.backward(gy.cuda())
    return y.detach().cpu().double(), xc.grad.cpu().double(), wc.grad.cpu().double()

  def close(a, b, what):
    tol = 1e-5 * max(float(b.abs().max()), 1.0)
    assert tuple(a.shape) == tuple(b.shape), what
    assert float((a - b).abs().max()) <= tol, (what, float((a - b).abs().max()), tol)

  y, dx, dw = run()
  close(y, yr.detach(), "y"); close(dx, xr.grad, "dx"); close(dw, wr.grad, "dw")
  monkeypatch.setenv("FRCNN_RESNET_S2_TC", "0")
  y0, dx0, dw0 = run()
  close(y0, yr.detach(), "y simt"); close(dx0, xr.grad, "dx simt"); close(dw0, wr.grad, "dw simt")
