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
