"""CPU tests of the N>1 host logic: world_size-2 gloo, gradient all-reduce through the optimizer wrapper."""
import os
import socket

import torch as t
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank = rank, world_size = world)
  from fasterrcnn_b200.optim import DataParallel
  t.manual_seed(0)
  w = t.nn.Parameter(t.ones(4, 3)); b = t.nn.Parameter(t.zeros(4))
  inner = t.optim.SGD([{"params": [w], "weight_decay": 0.0}], lr = 0.5, momentum = 0.0)     # weights only, like the reference
  opt = DataParallel(inner)
  x = t.full((2, 3), float(rank + 1))                       # per-rank "image"
  opt.zero_grad()
  loss = ((x @ w.t() + b) * float(rank + 1)).sum()
  loss.backward()
  local_grad = w.grad.clone()
  opt.step()
  out[rank] = (w.detach().clone(), local_grad, b.grad.clone(), opt.bytes_reduced_last_step)
  dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
  world, port = 2, _free_port()
  mgr = mp.Manager(); out = mgr.dict()
  mp.spawn(_worker, args = (world, port, out), nprocs = world, join = True)
  w0, g0, bg0, n0 = out[0]; w1, g1, bg1, n1 = out[1]
  mean_grad = (g0 + g1) / 2                                  # post-all-reduce gradient = mean of the per-image gradients
  expected = t.ones(4, 3) - 0.5 * mean_grad
  assert t.allclose(w0, expected) and t.allclose(w1, expected)         # replicas stay identical
  assert not t.allclose(bg0, bg1)                            # tensors outside the optimizer are not reduced
  assert n0 == n1 == 4 * 3 * 4


def test_sharded_sgd_arena_layout_on_cpu_tensors():
  """Host logic of optim.NvlsShardedSGD that needs no GPU: parameters re-pointed into a flat arena keep values and strides (channels-last
  filters included), gradients copied into the twin arena land at the same flat positions, and ONE elementwise SGD over the two flat
  arenas -- what csrc/dp_sgd.cu does shard by shard -- equals torch.optim.SGD applied tensor by tensor."""
  from fasterrcnn_b200.optim import NvlsShardedSGD
  t.manual_seed(0)
  conv = t.nn.Parameter(t.randn(8, 4, 3, 3).contiguous(memory_format = t.channels_last))
  lin = t.nn.Parameter(t.randn(5, 7))
  one = t.nn.Parameter(t.randn(6, 4, 1, 1))
  params = [conv, lin, one]
  offs, total = [], 0
  for p in params:
    offs.append(total); total += (p.numel() + 3) // 4 * 4
  world = 2
  shard = (total + 4 * world - 1) // (4 * world) * 4
  W, G = t.zeros(shard * world), t.zeros(shard * world)
  before = [p.detach().clone() for p in params]
  with t.no_grad():
    for p, off in zip(params, offs):
      v = NvlsShardedSGD._view(W, off, p); v.copy_(p); p.data = v
  for p, b in zip(params, before):
    assert t.equal(p, b) and p.stride() == b.stride()
  gv = {id(p): NvlsShardedSGD._view(G, off, p) for p, off in zip(params, offs)}

  def hook(p):
    gv[id(p)].copy_(p.grad)
  for p in params:
    p.register_post_accumulate_grad_hook(hook)
  x = t.randn(2, 4, 6, 6)
  y = t.nn.functional.conv2d(x, conv, padding = 1)
  (y.sum() + t.nn.functional.conv2d(y[:, :4], one).sum() + (lin * 2).sum()).backward()
  phys = lambda a, p: a.permute(0, 2, 3, 1).reshape(-1) if (p.dim() == 4 and not p.is_contiguous()) else a.reshape(-1)
  for p, off in zip(params, offs):
    assert t.equal(W[off:off + p.numel()], phys(p.detach(), p)) and t.equal(G[off:off + p.numel()], phys(p.grad, p))
  ref = [t.nn.Parameter(b.clone()) for b in before]
  opt = t.optim.SGD([{"params": [r], "weight_decay": 5e-4} for r in ref], lr = 1e-3, momentum = 0.9)
  for r, p in zip(ref, params):
    r.grad = p.grad.clone()
  opt.step()
  w = W.clone()
  W.copy_(w - 1e-3 * (G + 5e-4 * w))                          # first step: momentum buffer = gradient (+ weight decay)
  for r, p in zip(ref, params):
    assert t.allclose(p.detach(), r.detach(), atol = 1e-7)


def test_grad_arena_bucket_plan_for_vgg16():
  """Bucket layout of the data-parallel gradient arena on VGG-16's optimizer tensors (meta tensors: no memory): backward order (RPN
  branch, detector head, backbone in reverse), fc1 alone, a small final bucket, every bucket a whole number of float4s per rank."""
  from fasterrcnn_b200.optim import GradArena, backward_order
  from oracle import frcnn_oracle as orc
  shapes = orc.vgg16_param_shapes()
  trainable = set(orc.trainable_keys_vgg16(shapes))
  named = [(k, t.nn.Parameter(t.empty(shape, device = "meta"))) for k, shape in shapes.items() if k in trainable and k.endswith(".weight")]
  assert len(named) == 16                                                      # __main__.py:98-105: 16 tensors, 136.78 M elements
  order = backward_order(named)
  names = {id(p): k for k, p in named}
  got = [names[id(p)] for p in order]
  assert got[0].startswith("_stage2") and got[2].endswith("_rpn_conv1.weight")
  assert got[3].endswith("_regressor.weight") and got[6].endswith("_fc1.weight")
  assert got[7].endswith("_block5_conv3.weight") and got[-1].endswith("_block3_conv1.weight")
  for world in (2, 8):
    arena = GradArena(order, world, allocate = lambda n: t.empty((n,), device = "meta"))
    sizes = [b / 2 ** 20 for b in arena.payload_bytes]
    assert len(arena.buckets) == 4 and 70 < sizes[0] < 85 and 380 < sizes[1] < 400 and 45 < sizes[2] < 56 and sizes[3] < 8, sizes
    assert [names[id(p)] for p in order[arena.buckets[1][0]:arena.buckets[1][1]]] == ["_stage3_detector_network._pool_to_feature_vector._fc1.weight"]
    for (f0, f1, begin, end) in arena.buckets:
      assert begin % (32 * world) == 0 and end % (32 * world) == 0 and (end - begin) // world % 4 == 0
    assert all(o % 32 == 0 for o in arena.offsets)
    assert arena.total * 4 < 1.001 * sum(p.numel() for p in order) * 4 + 4096 * world


def _bucket_worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank = rank, world_size = world)
  from fasterrcnn_b200.optim import DataParallel
  t.manual_seed(0)
  a = t.nn.Parameter(t.randn(40, 30)); b = t.nn.Parameter(t.randn(50, 40)); c = t.nn.Parameter(t.randn(8, 50)); unused = t.nn.Parameter(t.randn(6, 6))
  named = [("_stage1.a.weight", a), ("_stage1.b.weight", b), ("_stage3.c.weight", c), ("_stage2.unused.weight", unused)]
  inner = t.optim.SGD([{"params": [p], "weight_decay": 0.0} for _, p in named], lr = 0.1, momentum = 0.9)
  opt = DataParallel(inner, named_params = named, bucket_bytes = 1500)        # several buckets: [unused, c] [b] [a] (backward order)
  launched = []
  orig = opt._launch_bucket
  opt._launch_bucket = lambda k: (launched.append(k), orig(k))[1]
  before = [p.detach().clone() for _, p in named]
  local = []
  for step in range(2):
    opt.zero_grad()
    x = t.full((3, 30), float(rank + 1 + step))
    y = (x @ a.t()) @ b.t()
    # rank 1 does not use c in step 0: its hook never fires there, yet the bucket must be reduced by both ranks, in order
    loss = (y @ c.t()).sum() if (rank == 0 or step == 1) else y.sum()
    loss.backward()
    local.append([None if p.grad is None else p.grad.detach().clone() for _, p in named])
    opt.step()
  out[rank] = (before, local, [p.detach().clone() for _, p in named], launched, len(opt.arena.buckets))
  dist.destroy_process_group()


def test_bucketed_allreduce_in_order_with_missing_gradients_gloo():
  world, port = 2, _free_port()
  mgr = mp.Manager(); out = mgr.dict()
  mp.spawn(_bucket_worker, args = (world, port, out), nprocs = world, join = True)
  before, g0, w0, launched0, nb = out[0]
  _, g1, w1, launched1, _ = out[1]
  assert nb >= 3
  assert launched0 == launched1 == list(range(nb)) * 2                          # same bucket order on both ranks, both steps
  # reference: SGD with momentum on the mean gradient (missing gradient = zeros)
  ref = [p.clone() for p in before]
  bufs = [None] * 4
  for step in range(2):
    for i in range(4):
      ga = g0[step][i] if g0[step][i] is not None else t.zeros_like(ref[i])
      gb = g1[step][i] if g1[step][i] is not None else t.zeros_like(ref[i])
      g = (ga + gb) / 2
      bufs[i] = g.clone() if bufs[i] is None else bufs[i] * 0.9 + g
      ref[i] = ref[i] - 0.1 * bufs[i]
  for i in range(4):
    assert t.allclose(w0[i], ref[i], atol = 1e-5) and t.equal(w0[i], w1[i]), i   # replicas stay identical
