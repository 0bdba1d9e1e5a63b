"""
GPU numerics of the data-parallel path: two ranks, each with its own 600x1000 image, identical VGG-16 replicas.
After train steps through each data-parallel optimizer,
  * the post-all-reduce gradient equals the SUM of the two per-image oracle gradients (the fused SGD applies 1 / world),
  * the post-step weights equal an oracle SGD step on the MEAN of the two per-image oracle gradients (1e-4),
  * the replicas are bit-identical.
Optimizers: DataParallel(FusedSGD) -- bucketed NCCL all-reduce on the gradient arena -- and NvlsShardedSGD -- the sharded step: fused
reduce-scatter + SGD + all-gather kernel over NVLink peer memory (bench.py's default at N > 1), the same through multimem instructions,
and with NCCL doing the two transfers.
Needs two devices (gpurun --gpus 2); skipped on a one-GPU box.
"""
import os
import random
import socket

import numpy as np
import pytest
import torch as t

from oracle import frcnn_oracle as orc

import _margins

pytestmark = pytest.mark.gpu

HW = (600, 1000)
STEPS = 2


class Box:
  def __init__(self, corners, class_index):
    self.corners, self.class_index, self.class_name = corners, class_index, str(class_index)


def _sample(rank, kind = "vgg16", hw = HW):
  gt = None if rank == 0 else [((60.0, 80.0, 300.0, 420.0), 3), ((200.0, 500.0, 560.0, 900.0), 12), ((20.0, 700.0, 180.0, 960.0), 9)]
  if hw != HW and gt is not None:
    gt = [(tuple(c * hw[0] / HW[0] if i % 2 == 0 else c * hw[1] / HW[1] for i, c in enumerate(b)), k) for b, k in gt]
  return orc.synthetic_sample(hw, seed = 100 + rank, gt = gt, backbone = kind)


def _params_and_backbone(kind):
  """(synthetic state dict, oracle backbone name, package backbone factory) for the two-rank tests."""
  if kind == "vgg16":
    import fasterrcnn_b200 as f
    return orc.synth_params(orc.vgg16_param_shapes(), seed = 5, heads = "reference"), lambda: f.vgg16.VGG16Backbone(dropout_probability = 0.0)
  from fasterrcnn_b200 import resnet
  from oracle import resnet_oracle
  params = orc.synth_params(resnet_oracle.param_shapes(kind), seed = 6, heads = "spread")
  for k in params:
    if k.endswith("bn3.weight"):
      params[k] = params[k] * 0.3                      # as in tests/test_model_gpu.py: keeps the deep residual stack in range
  return params, lambda: resnet.ResNetBackbone({"resnet50": resnet.Architecture.ResNet50, "resnet101": resnet.Architecture.ResNet101}[kind])


def _oracle_reference(world, params, kind = "vgg16", hw = HW):
  """The oracle's data-parallel step: per-image train_step (no update) on every emulated rank's own RNG streams, SGD on the mean gradient."""
  oracle = orc.OracleModel(params, backbone = kind)
  smps = [_sample(r, kind, hw) for r in range(world)]
  rng = []
  for r in range(world):
    random.seed(r); t.manual_seed(r)
    rng.append((random.getstate(), t.get_rng_state()))
  summed, scale = None, {}
  for _ in range(STEPS):
    grads = []
    for r in range(world):
      random.setstate(rng[r][0]); t.set_rng_state(rng[r][1])
      s = smps[r]
      oracle.train_step(s["image"], s["anchor_map"], s["anchor_valid_map"], s["gt_rpn_map"], s["gt_rpn_object_indices"], s["gt_rpn_background_indices"],
                        s["gt_corners"], s["gt_class_idxs"], apply_update = False)
      rng[r] = (random.getstate(), t.get_rng_state())
      grads.append({k: v.grad.clone() for k, v in oracle.params.items() if v.grad is not None})
    summed = {k: sum(g[k] for g in grads) for k in grads[0]}
    for k, g in summed.items():
      scale[k] = max(scale.get(k, 0.0), float(g.double().norm()))      # the largest this tensor's gradient has been over the steps
    for k, v in oracle.params.items():
      if k in summed:
        v.grad = summed[k] / world
    oracle.sgd_step(1e-3, 0.9, 5e-4)
  return oracle, summed, scale


def _worker(rank, world, port, which, out, kind = "vgg16", hw = HW):
  import hashlib
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
  t.cuda.set_device(rank)
  dist.init_process_group("nccl", rank = rank, world_size = world, device_id = t.device("cuda", rank))
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import optim
  t.set_num_threads(max(1, min(16, (os.cpu_count() or 8) // world)))
  params, make_backbone = _params_and_backbone(kind)
  oracle, summed, scale = _oracle_reference(world, params, kind, hw)               # every rank computes the same expectation
  model = f.FasterRCNNModel(num_classes = 21, backbone = make_backbone(), allow_edge_proposals = True)
  model.load_state_dict(params)
  model = model.cuda()
  named = list(model.named_parameters())
  if which == "nccl":
    optimizer = optim.DataParallel(optim.create_optimizer(model, 1e-3, 0.9, 5e-4, fused = True), named_params = named)
  else:
    optimizer = optim.NvlsShardedSGD(optim.optimizer_param_groups(model, 5e-4), lr = 1e-3, momentum = 0.9, named_params = named,
                                     exchange = {"nvls_multicast": "multicast", "nvls_peer": "peer", "nccl_sharded": "nccl"}[which])
  smp = _sample(rank, kind, hw)
  boxes = [Box(b, c) for b, c in zip(smp["gt_corners"], smp["gt_class_idxs"])]
  random.seed(rank); t.manual_seed(rank)
  for _ in range(STEPS):
    model.train_step(optimizer = optimizer, image_data = smp["image"].cuda(), anchor_map = smp["anchor_map"], anchor_valid_map = smp["anchor_valid_map"],
                     gt_rpn_map = smp["gt_rpn_map"].cuda(), gt_rpn_object_indices = [smp["gt_rpn_object_indices"]],
                     gt_rpn_background_indices = [smp["gt_rpn_background_indices"]], gt_boxes = [boxes])
  t.cuda.synchronize()
  dist.barrier()
  worst_w, worst_w_key, worst_g, digest = 0.0, "", 0.0, hashlib.sha256()
  grad_rels = []
  for k, p in named:
    w = p.detach().float().cpu()
    digest.update(np.ascontiguousarray(w.numpy()).tobytes())
    ref = oracle.params[k].detach()
    err = float(((w - ref).abs() / (1e-4 * ref.abs() + 2e-5)).max())        # in units of the bar: rtol 1e-4, atol 2e-5
    if err > worst_w:
      worst_w, worst_w_key = err, k
    if which == "nccl" and k in summed and p.grad is not None and p.requires_grad and "weight" in k:
      # an optimizer-owned tensor: its .grad is the arena view holding the all-reduced SUM of the last step (biases are not reduced)
      a, b = p.grad.detach().cpu().double(), summed[k].double()
      # relative to the gradient's norm, floored at 1e-4 of the largest norm this tensor's gradient had: with the reference's head
      # initialisation the softmax saturates after the first update and the classifier's true gradient collapses from 277 to 1e-7 --
      # pure cancellation noise on both sides, which a plain relative error would compare digit by digit
      rel = float((a - b).norm() / (b.norm() + 1e-4 * scale[k] + 1e-12))
      grad_rels.append((rel, k, float(a.norm()), float(b.norm())))
      worst_g = max(worst_g, rel)
  bucket_order = [optimizer.arena.bucket_of[id(optimizer.arena.params[i])] for i in optimizer.hook_order]
  out[rank] = dict(worst_weight_err_in_bars = worst_w, worst_weight = worst_w_key, worst_summed_grad_rel_l2 = worst_g, digest = digest.hexdigest(),
                   buckets = len(optimizer.arena.buckets), bytes_reduced = int(optimizer.bytes_reduced_last_step), hook_order = bucket_order,
                   multicast = bool(getattr(optimizer, "use_multicast", False)), worst_grads = str(sorted(grad_rels, reverse = True)[:4]))
  optimizer.remove_hooks()
  dist.destroy_process_group()


@pytest.mark.parametrize("which", ["nccl", "nvls_multicast", "nvls_peer", "nccl_sharded"])
def test_two_rank_step_matches_oracle_step_on_mean_gradient(which):
  import torch.multiprocessing as mp
  if t.cuda.device_count() < 2:
    pytest.skip("needs two GPUs (gpurun --gpus 2)")
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  mgr = mp.Manager(); out = mgr.dict()
  mp.spawn(_worker, args = (2, port, which, out), nprocs = 2, join = True)
  r0, r1 = out[0], out[1]
  _margins.record("dp2_" + which, **{k: v for k, v in r0.items() if k not in ("digest", "hook_order")}, replicas_identical = r0["digest"] == r1["digest"],
                  gradients_arrive_in_bucket_order = r0["hook_order"] == sorted(r0["hook_order"]), bucket_of_each_hook = str(r0["hook_order"]))
  assert r0["digest"] == r1["digest"]                                       # replicas bit-identical
  assert r0["hook_order"] == sorted(r0["hook_order"]), r0["hook_order"]     # backward produces the buckets in arena order: each is launched the moment it is complete
  assert r0["worst_weight_err_in_bars"] <= 1.0 and r1["worst_weight_err_in_bars"] <= 1.0, (r0, r1)
  assert 5.4e8 < r0["bytes_reduced"] < 5.6e8                                # all 16 optimizer tensors (136.78 M elements, 547 MB) crossed the wire
  if which == "nccl":
    assert r0["worst_summed_grad_rel_l2"] < 2e-2 and r1["worst_summed_grad_rel_l2"] < 2e-2
  if which == "nvls_multicast":
    assert r0["multicast"], "this box has NVSwitch multicast: the multimem path must be the one that ran"


@pytest.mark.parametrize("which", ["nccl", "nvls_peer"])
def test_two_rank_resnet50_step_matches_oracle_step_on_mean_gradient(which):
  """The same check on a ResNet backbone (BASELINE config 4's data-parallel path: frozen BatchNorm as the GEMM epilogue, stride-2
  bottleneck convolutions on the tensor cores, filter gradients written straight into the gradient arena), ResNet-50 at 384x512."""
  import torch.multiprocessing as mp
  if t.cuda.device_count() < 2:
    pytest.skip("needs two GPUs (gpurun --gpus 2)")
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  mgr = mp.Manager(); out = mgr.dict()
  mp.spawn(_worker, args = (2, port, which, out, "resnet50", (384, 512)), nprocs = 2, join = True)
  r0, r1 = out[0], out[1]
  _margins.record("dp2_resnet50_" + which, **{k: v for k, v in r0.items() if k not in ("digest", "hook_order")}, replicas_identical = r0["digest"] == r1["digest"],
                  gradients_arrive_in_bucket_order = r0["hook_order"] == sorted(r0["hook_order"]))
  assert r0["digest"] == r1["digest"]                                       # replicas bit-identical
  assert r0["worst_weight_err_in_bars"] <= 1.0 and r1["worst_weight_err_in_bars"] <= 1.0, (r0, r1)
  assert r0["bytes_reduced"] > 5e7                                          # the trainable filters of layer2-4 + heads crossed the wire
  if which == "nccl":
    assert r0["worst_summed_grad_rel_l2"] < 2e-2 and r1["worst_summed_grad_rel_l2"] < 2e-2
