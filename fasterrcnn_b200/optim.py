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


class NvlsShardedSGD:
  """EXPERIMENT (opt-in, written after round 1's GPU budget had ended -- unmeasured; bench.py: FRCNN_DP_FUSED=1 at N > 1).

  The data-parallel optimizer step as ONE hand-written kernel per rank over NVLink / NVSwitch instead of "NCCL all-reduce, then
  optimizer.step()": reduce-scatter of the weight gradients (multimem.ld_reduce: summed inside the switch), torch.optim.SGD on this
  rank's 1 / world shard (the momentum buffer exists only for the shard), all-gather of the updated weights (multimem.st) -- see
  csrc/dp_sgd.cu.  The optimizer's tensors are moved into two flat symmetric-memory arenas (torch.distributed._symmetric_memory gives
  the peer / multicast mappings and the stream-ordered cross-rank barrier): ``W`` -- the parameters become views of it -- and ``G``, into
  which each gradient is copied from a post-accumulate hook the moment autograd has produced it.
  Quacks like the optimizer FasterRCNNModel.train_step expects (zero_grad / step / param_groups).  Hyper-parameters must be uniform
  over the groups (they are in the reference's recipe, __main__.py:98-105)."""

  def __init__(self, params, lr = 1e-3, momentum = 0.9, process_group = None, use_multicast = None, ctas_per_sm = 0):
    os.environ.setdefault("TORCH_SYMMMEM_IMPLICIT_POOL", "0")   # one allocation per arena: the mappings then start at the tensor (offset 0)
    import torch.distributed._symmetric_memory as symm
    assert dist.is_initialized(), "NvlsShardedSGD needs an initialised process group (one process per GPU)"
    group = process_group if process_group is not None else dist.group.WORLD
    self.group, self.world_size, self.rank = group, dist.get_world_size(group), dist.get_rank(group)
    assert 1 <= self.world_size <= 8, "one NVSwitch domain: at most 8 ranks"
    self.param_groups = [dict(g) for g in params]
    for g in self.param_groups:
      g.setdefault("lr", lr); g.setdefault("momentum", momentum); g.setdefault("weight_decay", 0.0)
    hp = {(g["lr"], g["momentum"], g["weight_decay"]) for g in self.param_groups}
    assert len(hp) == 1, "NvlsShardedSGD: lr / momentum / weight_decay must be the same for every group"
    self.lr, self.momentum, self.weight_decay = hp.pop()
    self.params = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
    assert self.params and all(p.is_cuda and p.dtype == t.float32 for p in self.params)
    dev = self.params[0].device
    # flat layout: every tensor starts on a 16-byte boundary; the whole space is cut into world equal shards of whole float4s
    self.offsets, total = [], 0
    for p in self.params:
      assert p.is_contiguous() or p.is_contiguous(memory_format = t.channels_last), "parameters must be dense"
      self.offsets.append(total)
      total += (p.numel() + 3) // 4 * 4
    self.shard = (total + 4 * self.world_size - 1) // (4 * self.world_size) * 4
    self.total = self.shard * self.world_size
    self.W = symm.empty(self.total, dtype = t.float32, device = dev)
    self.G = symm.empty(self.total, dtype = t.float32, device = dev)
    self.W.zero_(); self.G.zero_()
    symm.enable_symm_mem_for_group(group.group_name)
    self.hW, self.hG = symm.rendezvous(self.W, group), symm.rendezvous(self.G, group)
    with t.no_grad():
      for p, off in zip(self.params, self.offsets):
        view = self._view(self.W, off, p)
        view.copy_(p)
        p.data = view                                            # same shape, same strides; storage = this rank's weight arena
    self._gviews = {id(p): self._view(self.G, off, p) for p, off in zip(self.params, self.offsets)}
    self.momentum_shard = t.zeros((self.shard,), dtype = t.float32, device = dev)
    self._first = True
    self._seen = set()
    self.ctas_per_sm = int(ctas_per_sm) if ctas_per_sm else int(os.environ.get("FRCNN_DP_FUSED_CTAS", "0"))   # 0: the kernel's default (4 CTAs of 256 threads per SM)
    self.sm_reserve = 0
    self.bytes_reduced_last_step = 0
    # mappings: multicast (in-switch reduction / broadcast) when the fabric offers it, else every rank's arena through peer pointers
    def mapping(handle, tensor):
      """(peer pointers of `tensor` on every rank, multicast pointer or 0): the handle describes an allocation block, the tensor sits at
      handle.offset inside it (0 with one allocation per arena) -- checked against the one address known for sure, this rank's."""
      get = lambda name: (lambda v: v() if callable(v) else v)(getattr(handle, name))     # properties in current torch
      base = [int(x) for x in get("buffer_ptrs")]
      off = int(get("offset"))
      if base[self.rank] + off != tensor.data_ptr():
        assert base[self.rank] == tensor.data_ptr(), "symmetric-memory mapping does not contain the tensor where expected"
        off = 0
      mc = int(get("multicast_ptr") or 0)
      return [b + off for b in base], (mc + off if mc else 0)
    peers_g, mc_g = mapping(self.hG, self.G)
    peers_w, mc_w = mapping(self.hW, self.W)
    if use_multicast is None:
      use_multicast = os.environ.get("FRCNN_DP_FUSED_MULTICAST", "1") not in ("", "0")
    self.use_multicast = bool(use_multicast) and mc_w != 0 and mc_g != 0
    self._mc = (mc_g, mc_w) if self.use_multicast else (None, None)
    import ctypes
    vp = ctypes.c_void_p * self.world_size
    self._peers = (vp(*peers_g), vp(*peers_w))
    self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad_ready) for p in self.params]
    dist.barrier(group)

  @staticmethod
  def _view(buf, off, p):
    return buf[off:off + p.numel()].as_strided(p.shape, p.stride())

  @t.no_grad()
  def _on_grad_ready(self, p):
    self._gviews[id(p)].copy_(p.grad)                            # on the compute stream, right behind the kernel that produced the gradient
    self._seen.add(id(p))
    self.bytes_reduced_last_step += p.grad.numel() * 4

  def zero_grad(self, set_to_none = True):
    self.bytes_reduced_last_step = 0
    self._seen = set()
    for p in self.params:
      if set_to_none:
        p.grad = None
      elif p.grad is not None:
        p.grad.zero_()

  @t.no_grad()
  def step(self):
    from . import _lib
    for p in self.params:
      if id(p) not in self._seen:
        self._gviews[id(p)].zero_()                              # no gradient on this rank this step (e.g. an empty RoI sample): contribute zero, not last step's values
    self.hG.barrier(channel = 0)                                 # every rank's gradients are in its arena
    _lib.check(_lib.lib().frcnn_dp_sgd_fused(self._mc[0], self._mc[1], self._peers[0], self._peers[1], self.world_size, _lib.ptr(self.W), _lib.ptr(self.momentum_shard),
                                             self.rank * self.shard, self.shard, float(self.lr), float(self.momentum), float(self.weight_decay),
                                             1.0 / self.world_size, 1 if self._first else 0, self.ctas_per_sm, _lib.stream()), "frcnn_dp_sgd_fused")
    _lib.count()
    self.hG.barrier(channel = 0)                                 # every shard delivered everywhere; every arena's gradients consumed
    self._first = False

  def remove_hooks(self):
    for h in self._hooks:
      h.remove()
    self._hooks = []


def optimizer_param_groups(model, weight_decay = 5e-4):
  """The reference's parameter groups (__main__.py:98-105): one per tensor that requires grad and has "weight" in its name."""
  return [{"params": [value], "weight_decay": weight_decay} for key, value in dict(model.named_parameters()).items() if value.requires_grad and "weight" in key]


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
