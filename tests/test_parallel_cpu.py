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
