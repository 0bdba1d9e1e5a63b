"""
Optimizer side of the hot path.

FusedSGD       torch.optim.SGD as the reference configures it (__main__.py:98-105: momentum,
               L2 weight decay, no dampening / nesterov) in ONE kernel per tensor (K10): the update
               reads param, grad, momentum and writes param, momentum (5 accesses x 4 B/elem),
               with the 1/world_size gradient scale of data-parallel runs folded in.
DataParallel   one process per GPU: wraps any optimizer; gradient all-reduce
               (NCCL over NVLink, sum, fp32) is launched per tensor from a post-accumulate hook the
               moment autograd finishes that gradient, so the 411 MB fc1 reduction overlaps the conv
               backward still running on the compute stream; ``step()`` waits for the handles and
               applies the update with grad_scale = 1/world_size.
Only tensors owned by the wrapped optimizer are reduced (the reference never steps biases).
               While reductions are in flight (first gradient hook .. step()) the persistent tcgen05 GEMMs can be told to leave
               ``sm_reserve`` SMs to NCCL's resident CTAs (frcnn_set_sm_reserve; FRCNN_DP_SM_RESERVE, default 0 = off): a one-CTA-per-SM
               grid that finds SMs taken runs a second, nearly empty wave, which is what the 2-GPU run loses.
"""
import os

import torch as t
import torch.distributed as dist

from . import ops


class FusedSGD(t.optim.Optimizer):
  """eager (EXPERIMENT, off by default; FRCNN_EAGER_SGD=1): tensors of at least ``eager_min_numel`` elements (VGG-16: fc1, fc2 = 87 % of the
  optimizer's bytes) are updated from a post-accumulate-grad hook on a side stream, as soon as their gradient exists, while the
  tensor-pipe-bound convolution backward still runs on the compute stream -- the update is HBM-bound and its 256-thread / 32-register CTAs
  fit beside a GEMM CTA.  ``step()`` then covers the remaining tensors and joins the side stream.  Same arithmetic, same result; only the
  schedule differs.  Single-GPU only (with DataParallel the gradient is not final in the hook).  Unmeasured in round 1."""

  def __init__(self, params, lr = 1e-3, momentum = 0.9, weight_decay = 0.0, eager = None, eager_min_numel = 1 << 23, eager_ctas_per_sm = None):
    super().__init__(params, dict(lr = lr, momentum = momentum, weight_decay = weight_decay))
    self.grad_scale = 1.0
    if eager is None:
      eager = os.environ.get("FRCNN_EAGER_SGD", "0") not in ("", "0")
    self.eager = bool(eager)
    self.eager_ctas_per_sm = int(os.environ.get("FRCNN_EAGER_SGD_CTAS", "1")) if eager_ctas_per_sm is None else int(eager_ctas_per_sm)
    self._side, self._eager_done, self._eager_hooks, self._group_of = None, set(), [], {}
    if self.eager:
      for group in self.param_groups:
        for p in group["params"]:
          self._group_of[id(p)] = group
          if p.requires_grad and p.numel() >= eager_min_numel:
            self._eager_hooks.append(p.register_post_accumulate_grad_hook(self._eager_update))

  def _entry(self, p, group):
    state = self.state[p]
    first = "momentum_buffer" not in state
    if first:
      state["momentum_buffer"] = t.empty_like(p)          # preserves the channels_last strides of filters
    g = p.grad
    if g.stride() != p.stride():
      g = g.contiguous(memory_format = t.channels_last) if p.dim() == 4 and not p.is_contiguous() else g.contiguous()
    # matrices / filters feed the tcgen05 GEMMs next step: their operand split is produced by the same kernel
    return (p, g, state["momentum_buffer"], group["lr"], group["momentum"], group["weight_decay"], first, p.dim() >= 2 and p.numel() >= 4096)

  @t.no_grad()
  def _eager_update(self, p):
    if self.grad_scale != 1.0 or not p.is_cuda or p.grad is None:
      return                                                 # data-parallel run: the gradient still has to be reduced
    main = t.cuda.current_stream()
    if self._side is None:
      self._side = t.cuda.Stream()
    ready = t.cuda.Event()
    ready.record(main)                                       # behind this layer's dgrad (last reader of the weights) and wgrad
    entry = self._entry(p, self._group_of[id(p)])            # allocations happen on the compute stream's pool
    with t.cuda.stream(self._side):
      self._side.wait_event(ready)
      ops.sgd_step_multi([entry], self.grad_scale, self.eager_ctas_per_sm)
    self._eager_done.add(id(p))

  @t.no_grad()
  def step(self, closure = None):
    assert closure is None
    entries = []
    for group in self.param_groups:
      for p in group["params"]:
        if p.grad is None or id(p) in self._eager_done:
          continue
        entries.append(self._entry(p, group))
    ops.sgd_step_multi(entries, self.grad_scale)                 # every parameter group in one C call
    if self._eager_done:
      t.cuda.current_stream().wait_stream(self._side)            # the next forward reads the eagerly updated weights
      self._eager_done.clear()


class DataParallel:
  """Optimizer wrapper: overlapped gradient all-reduce + (optionally fused) update.  Quacks like the
  optimizer ``FasterRCNNModel.train_step`` expects (zero_grad / step / param_groups)."""

  def __init__(self, optimizer, process_group = None, sm_reserve = None):
    self.optimizer = optimizer
    self.group = process_group
    self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
    if sm_reserve is None:
      sm_reserve = int(os.environ.get("FRCNN_DP_SM_RESERVE", "0"))
    self.sm_reserve = int(sm_reserve) if self.world_size > 1 else 0
    self._reserved = False
    self._handles = []
    self._hooks = []
    self.bytes_reduced_last_step = 0
    if self.world_size > 1:
      for group in optimizer.param_groups:
        for p in group["params"]:
          if p.requires_grad:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
    if isinstance(optimizer, FusedSGD):
      optimizer.grad_scale = 1.0 / self.world_size

  @property
  def param_groups(self):
    return self.optimizer.param_groups

  def _on_grad_ready(self, p):
    if self.sm_reserve > 0 and not self._reserved:
      from . import _lib
      _lib.set_sm_reserve(self.sm_reserve)                      # GEMMs launched from here on leave room for NCCL's CTAs
      self._reserved = True
    # the gradient was produced on the current (compute) stream; NCCL orders itself after it
    self._handles.append((p, dist.all_reduce(p.grad, op = dist.ReduceOp.SUM, group = self.group, async_op = True)))

  def zero_grad(self, set_to_none = True):
    self._handles = []
    self.optimizer.zero_grad(set_to_none = set_to_none)

  def step(self):
    nbytes = 0
    for p, h in self._handles:
      h.wait()                                                  # compute stream waits for the reduction
      nbytes += p.grad.numel() * p.grad.element_size()
    self.bytes_reduced_last_step = nbytes
    if self._reserved:
      from . import _lib
      _lib.set_sm_reserve(0)                                    # the reductions are behind the compute stream now: all SMs again
      self._reserved = False
    if self.world_size > 1 and not isinstance(self.optimizer, FusedSGD):
      for p, _ in self._handles:
        p.grad.div_(self.world_size)
    self._handles = []
    self.optimizer.step()

  def remove_hooks(self):
    for h in self._hooks:
      h.remove()
    self._hooks = []


def create_optimizer(model, learning_rate = 1e-3, momentum = 0.9, weight_decay = 5e-4, fused = True):
  """The reference's recipe (__main__.py:98-105): one group per tensor that requires grad and has
  "weight" in its name."""
  params = []
  for key, value in dict(model.named_parameters()).items():
    if not value.requires_grad:
      continue
    if "weight" in key:
      params += [{"params": [value], "weight_decay": weight_decay}]
  if fused:
    return FusedSGD(params, lr = learning_rate, momentum = momentum)
  return t.optim.SGD(params, lr = learning_rate, momentum = momentum)
