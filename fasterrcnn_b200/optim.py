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
  """eager (default on; FRCNN_EAGER_SGD=0 turns it off): tensors of at least ``eager_min_numel`` elements (VGG-16: fc1, fc2 = 87 % of the
  optimizer's bytes) are updated from a post-accumulate-grad hook on a side stream, as soon as their gradient exists, while the convolution
  backward still runs on the compute stream -- the GEMM CTAs leave register room for the update kernel's 256-thread CTAs (conv_tc.cu), two
  of which per SM measured best.  ``step()`` then covers the remaining tensors and joins the side stream.  Same arithmetic, same result
  (tests/test_zz_experiments_gpu.py: bit-identical weights); only the schedule differs.  Single-GPU only (under DataParallel the gradient
  is not final in the hook, and the update is skipped there).  Measured with the proposal chain on its own stream as well
  (FasterRCNNModel.train_step): 5.325 -> 5.247 ms per step, two runs each on one box; either change alone is within noise."""

  def __init__(self, params, lr = 1e-3, momentum = 0.9, weight_decay = 0.0, eager = None, eager_min_numel = 1 << 23, eager_ctas_per_sm = None):
    super().__init__(params, dict(lr = lr, momentum = momentum, weight_decay = weight_decay))
    self.grad_scale = 1.0
    if eager is None:
      eager = os.environ.get("FRCNN_EAGER_SGD", "1") not in ("", "0")
    self.eager = bool(eager)
    self.eager_ctas_per_sm = int(os.environ.get("FRCNN_EAGER_SGD_CTAS", "2")) if eager_ctas_per_sm is None else int(eager_ctas_per_sm)
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


def backward_order(named_params):
  """Optimizer tensors in the order train_step's backward produces their gradients (faster_rcnn.py: the RPN branch is back-propagated
  first, then the detector head, then the shared backbone -- each in reverse layer order).  Only overlap depends on this being right:
  buckets are launched strictly in order on every rank whatever order the hooks fire in."""
  groups = {"_stage2": [], "_stage3": [], "_stage1": []}
  other = []
  for name, p in named_params:
    for prefix, lst in groups.items():
      if name.startswith(prefix):
        lst.append(p)
        break
    else:
      other.append(p)
  return list(reversed(groups["_stage2"])) + list(reversed(groups["_stage3"])) + list(reversed(groups["_stage1"])) + list(reversed(other))


class GradArena:
  """One flat fp32 buffer holding the weight gradients of the optimizer's tensors back to back, in backward order, cut into a few
  buckets of consecutive tensors.  The filter-gradient kernels write straight into it (ops.register_grad_destination), so a bucket is
  one contiguous range: reduced in place by ONE collective, read in place by the optimizer kernel -- no flatten / unflatten copies.
  Every tensor starts on a 128-byte boundary; every bucket on a (128 x world)-byte one, so that bucket / world is a whole number of
  float4s (the fused kernel's shards)."""

  def __init__(self, params, world, bucket_bytes = 48 << 20, last_bucket_bytes = 8 << 20, allocate = None):
    self.params = list(params)
    assert self.params and all(p.dtype == t.float32 for p in self.params)
    self.world = int(world)
    align_t, align_b = 32, 32 * self.world
    # greedy buckets: close one as soon as it holds >= bucket_bytes; then carve a small final bucket off the tail (the last collective
    # cannot overlap anything: backward ends with it)
    cuts, acc = [], 0
    for i, p in enumerate(self.params):
      acc += p.numel() * 4
      if acc >= bucket_bytes:
        cuts.append(i + 1); acc = 0
    if not cuts or cuts[-1] != len(self.params):
      cuts.append(len(self.params))
    start = cuts[-2] if len(cuts) > 1 else 0
    if sum(p.numel() * 4 for p in self.params[start:]) > last_bucket_bytes:
      tail, j = 0, len(self.params)
      while j > start + 1 and tail + self.params[j - 1].numel() * 4 <= last_bucket_bytes:
        j -= 1; tail += self.params[j].numel() * 4
      if start < j < len(self.params):
        cuts.insert(len(cuts) - 1, j)
    self.offsets, self.buckets = [], []                      # element offset per tensor | (first tensor, end tensor, begin elem, end elem)
    at, first = 0, 0
    for end in cuts:
      begin = at
      for p in self.params[first:end]:
        self.offsets.append(at)
        at += (p.numel() + align_t - 1) // align_t * align_t
      at = (at + align_b - 1) // align_b * align_b
      self.buckets.append((first, end, begin, at))
      first = end
    self.total = at
    self.payload_bytes = [sum(p.numel() * 4 for p in self.params[f0:f1]) for f0, f1, _, _ in self.buckets]   # without the alignment padding
    self.bucket_of = {}
    for b, (f0, f1, _, _) in enumerate(self.buckets):
      for i in range(f0, f1):
        self.bucket_of[id(self.params[i])] = b
    dev = self.params[0].device
    self.flat = (allocate or (lambda n: t.zeros((n,), dtype = t.float32, device = dev)))(self.total)
    self.index = {id(p): i for i, p in enumerate(self.params)}

  def register(self):
    from . import ops
    for p, off in zip(self.params, self.offsets):
      ops.register_grad_destination(p, self.flat, off)

  def unregister(self):
    from . import ops
    ops.clear_grad_destinations(self.params)

  def view(self, p, buf = None):
    off = self.offsets[self.index[id(p)]]
    return (self.flat if buf is None else buf)[off:off + p.numel()].as_strided(p.shape, p.stride())

  def adopt(self, p):
    """Makes p.grad the arena view (copying only if the producing kernel did not write there already)."""
    v = self.view(p)
    if p.grad.data_ptr() != v.data_ptr() or p.grad.stride() != v.stride():
      v.copy_(p.grad)
      p.grad = v
    return v


class _BucketedHooks:
  """Shared machinery of the two data-parallel optimizers: post-accumulate hooks count a bucket's tensors; a bucket is handed to
  ``_launch_bucket`` as soon as it AND every earlier bucket is complete (same launch order on every rank, whatever fires when);
  ``_flush`` at step() zero-fills the tensors that received no gradient on this rank and launches what is left."""

  def _init_hooks(self, arena):
    self.arena = arena
    self._fired = [0] * len(arena.buckets)
    self._seen = set()
    self._next = 0
    self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad_ready) for p in arena.params if p.requires_grad]
    self.hook_order = []                                   # (debug) firing order of the last step, as arena indices

  @t.no_grad()
  def _on_grad_ready(self, p):
    if id(p) in self._seen:
      return
    self.arena.adopt(p)
    self._seen.add(id(p))
    self.hook_order.append(self.arena.index[id(p)])
    b = self.arena.bucket_of[id(p)]
    self._fired[b] += 1
    while self._next < len(self.arena.buckets) and self._fired[self._next] == self.arena.buckets[self._next][1] - self.arena.buckets[self._next][0]:
      self._launch_bucket(self._next)
      self._next += 1

  @t.no_grad()
  def _flush(self):
    for p in self.arena.params:
      if id(p) not in self._seen:
        v = self.arena.view(p)
        v.zero_()                                          # no gradient on this rank this step (e.g. an empty RoI sample): contribute zeros
        p.grad = v
    while self._next < len(self.arena.buckets):
      self._launch_bucket(self._next)
      self._next += 1

  def _reset_step(self):
    self._fired = [0] * len(self.arena.buckets)
    self._seen = set()
    self._next = 0
    self.hook_order = []

  def remove_hooks(self):
    for h in self._hooks:
      h.remove()
    self._hooks = []
    self.arena.unregister()


class DataParallel(_BucketedHooks):
  """Optimizer wrapper for one-process-per-GPU data parallelism: bucketed, overlapped gradient all-reduce + update.
  Quacks like the optimizer ``FasterRCNNModel.train_step`` expects (zero_grad / step / param_groups).

  The weight gradients live in a GradArena (written there directly by the filter-gradient kernels); each bucket -- VGG-16: {RPN, heads,
  fc2} 78 MB | fc1 411 MB | blocks 5-4 52 MB | block 3 6 MB -- is all-reduced in place (NCCL over NVLink, sum, fp32) the moment its last
  gradient exists, on NCCL's own stream, so fc1's reduction overlaps the convolution backward; ``step()`` waits for the handles and the
  fused SGD applies 1 / world.  Every rank reduces every bucket every step, in the same order (a tensor without a gradient on this
  rank contributes zeros -- the semantics of torch's DistributedDataParallel).  While reductions are in flight the persistent tcgen05
  GEMMs leave ``sm_reserve`` SMs to NCCL's CTAs (frcnn_set_sm_reserve; FRCNN_DP_SM_RESERVE): a one-CTA-per-SM grid that finds SMs
  taken would run a second, nearly empty wave.  With world_size 1 (or no process group) it is a transparent wrapper."""

  def __init__(self, optimizer, process_group = None, sm_reserve = None, named_params = None, bucket_bytes = None):
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
    self.arena = None
    if self.world_size > 1:
      owned = [p for group in optimizer.param_groups for p in group["params"] if p.requires_grad]
      if named_params is not None:
        ids = {id(p) for p in owned}
        ordered = [p for p in backward_order(named_params) if id(p) in ids]
        owned = ordered + [p for p in owned if id(p) not in {id(q) for q in ordered}]
      else:
        owned = list(reversed(owned))
      if bucket_bytes is None:
        bucket_bytes = int(os.environ.get("FRCNN_DP_BUCKET_MB", "48")) << 20
      cuda = all(p.is_cuda for p in owned)
      arena = GradArena(owned, self.world_size, bucket_bytes = bucket_bytes)
      if cuda:
        arena.register()                                     # CUDA parameters: the wgrad kernels write into the arena
      self._init_hooks(arena)
    if isinstance(optimizer, FusedSGD):
      optimizer.grad_scale = 1.0 / self.world_size

  @property
  def param_groups(self):
    return self.optimizer.param_groups

  def _launch_bucket(self, b):
    if self.sm_reserve > 0 and not self._reserved:
      from . import _lib
      _lib.set_sm_reserve(self.sm_reserve)                      # GEMMs launched from here on leave room for NCCL's CTAs
      self._reserved = True
    _, _, begin, end = self.arena.buckets[b]
    # the gradients were produced on the current (compute) stream; NCCL orders itself after it
    self._handles.append((self.arena.payload_bytes[b], dist.all_reduce(self.arena.flat[begin:end], op = dist.ReduceOp.SUM, group = self.group, async_op = True)))

  def zero_grad(self, set_to_none = True):
    self._handles = []
    if self.arena is not None:
      self._reset_step()
    self.optimizer.zero_grad(set_to_none = True if self.arena is not None else set_to_none)

  def step(self):
    nbytes = 0
    if self.arena is not None:
      self._flush()
    for n, h in self._handles:
      h.wait()                                                  # compute stream waits for the reduction
      nbytes += n
    self.bytes_reduced_last_step = nbytes
    if self._reserved:
      from . import _lib
      _lib.set_sm_reserve(0)                                    # the reductions are behind the compute stream now: all SMs again
      self._reserved = False
    if self.world_size > 1 and not isinstance(self.optimizer, FusedSGD):
      self.arena.flat.div_(self.world_size)
    self._handles = []
    self.optimizer.step()

  def remove_hooks(self):
    if self.arena is not None:
      _BucketedHooks.remove_hooks(self)


class NvlsShardedSGD(_BucketedHooks):
  """The data-parallel optimizer step as hand-written kernels over NVLink / NVSwitch instead of "NCCL all-reduce, then optimizer.step()"
  (csrc/dp_sgd.cu): per bucket ONE kernel per rank does the reduce-scatter of the weight gradients (multimem.ld_reduce: summed inside the
  switch), torch.optim.SGD on this rank's 1 / world shard of the bucket (the momentum buffer exists only for the shards: 1 / world of its
  memory and traffic) and the all-gather of the updated weights (multimem.st).  The optimizer's tensors live in two flat symmetric-memory
  arenas with one layout (GradArena): ``W`` -- the parameters become views of it -- and ``G`` -- the filter-gradient kernels write into it
  directly.  A bucket's kernel is launched on a side stream the moment its last gradient exists, bracketed by stream-ordered cross-rank
  barriers (torch.distributed._symmetric_memory), in 128-thread CTAs that fit beside a resident GEMM CTA, so reduce + update + broadcast of
  fc1's 411 MB run UNDER the convolution backward and no optimizer pass is left after it; ``step()`` only joins the side stream.
  Quacks like the optimizer FasterRCNNModel.train_step expects (zero_grad / step / param_groups).  Hyper-parameters must be uniform
  over the groups (they are in the reference's recipe, __main__.py:98-105).  A tensor that received no gradient on any rank still gets
  weight decay + momentum (zeros are reduced: DistributedDataParallel's semantics, not torch.optim.SGD's skip)."""

  def __init__(self, params, lr = 1e-3, momentum = 0.9, process_group = None, use_multicast = None, ctas_per_sm = 0, named_params = None, bucket_bytes = None,
               overlap = None, exchange = None):
    """exchange: "peer" (default; csrc/dp_sgd.cu through peer pointers), "multicast" (same kernel through multimem instructions), or
    "nccl": the same sharded step with the two transfers done by NCCL -- reduce-scatter of the bucket into this rank's shard, the
    library's SGD kernel on the shard, all-gather of the updated shard -- for fabrics without peer mappings and as the measured
    alternative at 8 ranks (profiles/r02_dp_n8.md).  FRCNN_DP_EXCHANGE overrides."""
    os.environ.setdefault("TORCH_SYMMMEM_IMPLICIT_POOL", "0")   # one allocation per arena: the mappings then start at the tensor (offset 0)
    assert dist.is_initialized(), "NvlsShardedSGD needs an initialised process group (one process per GPU)"
    if exchange is None:
      exchange = os.environ.get("FRCNN_DP_EXCHANGE", "multicast" if use_multicast else "peer")
    assert exchange in ("peer", "multicast", "nccl"), exchange
    self.exchange = exchange
    if exchange == "multicast":
      use_multicast = True
    if exchange != "nccl":
      import torch.distributed._symmetric_memory as symm
    group = process_group if process_group is not None else dist.group.WORLD
    self.group, self.world_size, self.rank = group, dist.get_world_size(group), dist.get_rank(group)
    assert 1 <= self.world_size <= 8, "one NVSwitch domain: at most 8 ranks"
    self.param_groups = [dict(g) for g in params]
    for g in self.param_groups:
      g.setdefault("lr", lr); g.setdefault("momentum", momentum); g.setdefault("weight_decay", 0.0)
    hp = {(g["lr"], g["momentum"], g["weight_decay"]) for g in self.param_groups}
    assert len(hp) == 1, "NvlsShardedSGD: lr / momentum / weight_decay must be the same for every group"
    self.lr, self.momentum, self.weight_decay = hp.pop()
    owned = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
    assert owned and all(p.is_cuda and p.dtype == t.float32 for p in owned)
    for p in owned:
      assert p.is_contiguous() or p.is_contiguous(memory_format = t.channels_last), "parameters must be dense"
    if named_params is not None:
      ids = {id(p) for p in owned}
      ordered = [p for p in backward_order(named_params) if id(p) in ids]
      owned = ordered + [p for p in owned if id(p) not in {id(q) for q in ordered}]
    else:
      owned = list(reversed(owned))
    self.params = owned
    dev = owned[0].device
    if bucket_bytes is None:
      bucket_bytes = int(os.environ.get("FRCNN_DP_BUCKET_MB", "48")) << 20
    # every fallible step first; the parameters are re-pointed into the arena only once all ranks agree that the set-up succeeded
    error = None
    peers_g = peers_w = [0] * self.world_size
    mc_g = mc_w = 0
    try:
      if exchange == "nccl":
        arena = GradArena(owned, self.world_size, bucket_bytes = bucket_bytes)
        self.W = t.zeros((arena.total,), dtype = t.float32, device = dev)
        self.G = arena.flat
        self.hW = self.hG = None
      else:
        symm.enable_symm_mem_for_group(group.group_name)
        arena = GradArena(owned, self.world_size, bucket_bytes = bucket_bytes, allocate = lambda n: symm.empty(n, dtype = t.float32, device = dev).zero_())
        self.W = symm.empty(arena.total, dtype = t.float32, device = dev).zero_()
        self.G = arena.flat
        self.hW, self.hG = symm.rendezvous(self.W, group), symm.rendezvous(self.G, group)
        peers_g, mc_g = self._mapping(self.hG, self.G)
        peers_w, mc_w = self._mapping(self.hW, self.W)
    except Exception as e:                                     # noqa: BLE001
      error = e
    ok = t.tensor([0 if error is not None else 1], dtype = t.int32, device = dev)
    dist.all_reduce(ok, op = dist.ReduceOp.MIN, group = group)
    if int(ok.item()) == 0:
      raise RuntimeError("NvlsShardedSGD: symmetric-memory set-up failed on at least one rank (%s)" % (error,))
    if use_multicast is None:
      # default = peer pointers: every element of the shard is read from each rank's arena and the result stored to each rank's, which
      # moves 547 MB x (W - 1) / W per direction; the multicast path (multimem.ld_reduce / multimem.st) also sends every rank's OWN share
      # through the switch and back: 547 MB x (1 + 1 / W) per direction -- 3x the bytes at W = 2, still 1.3x at W = 8 (measured:
      # profiles/r02_dp_sweep_n2.md)
      use_multicast = os.environ.get("FRCNN_DP_FUSED_MULTICAST", "0") not in ("", "0")
    self.use_multicast = bool(use_multicast) and mc_w != 0 and mc_g != 0
    self._mc = (mc_g, mc_w) if self.use_multicast else (None, None)
    import ctypes
    vp = ctypes.c_void_p * self.world_size
    self._peers = (vp(*peers_g), vp(*peers_w))
    with t.no_grad():
      for p in owned:
        view = arena.view(p, self.W)
        view.copy_(p)
        p.data = view                                            # same shape, same strides; storage = this rank's weight arena
    from . import ops
    ops.invalidate_weight_splits()                               # the parameters moved
    # momentum: one shard per bucket, back to back
    self._shards, at = [], 0
    for (_, _, begin, end) in arena.buckets:
      n = (end - begin) // self.world_size
      self._shards.append((begin + self.rank * n, n, at))
      at += n
    self.momentum_shard = t.zeros((at,), dtype = t.float32, device = dev)
    self._first = [True] * len(arena.buckets)
    if overlap is None:
      overlap = os.environ.get("FRCNN_DP_FUSED_OVERLAP", "1") not in ("", "0")
    self.overlap = bool(overlap)                                 # False: every bucket at step(), on the compute stream
    # launch shapes: under the backward TWO 256-thread CTAs per SM (measured best at 2 GPUs: 6.04 -> 5.86 ms / step against one; they fit
    # beside a resident GEMM CTA); alone on the GPU 8 per SM
    self.ctas_per_sm = int(ctas_per_sm) if ctas_per_sm else int(os.environ.get("FRCNN_DP_FUSED_CTAS", "2" if self.overlap else "8"))
    self.split_ctas_per_sm = int(os.environ.get("FRCNN_DP_SPLIT_CTAS", "2" if self.overlap else "8"))
    self._side = t.cuda.Stream(device = dev) if self.overlap else None
    self._used_side = False
    self._deferred = []
    self.sm_reserve = 0
    self.bytes_reduced_last_step = 0
    arena.register()
    self._init_hooks(arena)
    dist.barrier(group)

  def _mapping(self, handle, tensor):
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

  @staticmethod
  def _view(buf, off, p):
    return buf[off:off + p.numel()].as_strided(p.shape, p.stride())

  def _fused(self, b, pre_barrier = True, post_barrier = True):
    from . import _lib
    begin, n, mom_at = self._shards[b]
    if self.exchange == "nccl":
      # (blocking collectives: NCCL runs them on its own stream and makes the CURRENT stream -- the side stream when overlapped -- wait)
      _, _, b0, b1 = self.arena.buckets[b]
      dist.reduce_scatter_tensor(self.G[begin:begin + n], self.G[b0:b1], op = dist.ReduceOp.SUM, group = self.group)
      _lib.check(_lib.lib().frcnn_sgd_step_split(_lib.ptr(self.W[begin:begin + n]), _lib.ptr(self.G[begin:begin + n]), _lib.ptr(self.momentum_shard[mom_at:mom_at + n]), n,
                                                 float(self.lr), float(self.momentum), float(self.weight_decay), 1.0 / self.world_size, 1 if self._first[b] else 0, None,
                                                 _lib.stream()), "frcnn_sgd_step_split")
      _lib.count()
      dist.all_gather_into_tensor(self.W[b0:b1], self.W[begin:begin + n], group = self.group)
      self._first[b] = False
      self.bytes_reduced_last_step += self.arena.payload_bytes[b]
      if post_barrier:
        self._resplit(b)
      return
    if pre_barrier:
      self.hG.barrier(channel = 0)                               # every rank's gradients of this bucket are in its arena
    _lib.check(_lib.lib().frcnn_dp_sgd_fused(self._mc[0], self._mc[1], self._peers[0], self._peers[1], self.world_size, _lib.ptr(self.W),
                                             _lib.ptr(self.momentum_shard[mom_at:mom_at + n]), begin, n, float(self.lr), float(self.momentum), float(self.weight_decay),
                                             1.0 / self.world_size, 1 if self._first[b] else 0, self.ctas_per_sm, _lib.stream()), "frcnn_dp_sgd_fused")
    _lib.count()
    self._first[b] = False
    self.bytes_reduced_last_step += self.arena.payload_bytes[b]
    if post_barrier:
      self.hG.barrier(channel = 0)                               # every shard delivered everywhere; every arena's gradients consumed
      self._resplit(b)

  def _resplit(self, b):
    """The updated weights of bucket b arrived through the multicast mapping: bring the operand splits the tensor-core GEMMs read up to
    date right here, on the stream the update ran on (under the backward), ONE pass per tensor with the exponent its buffer carries
    (a full amax + split pass seeds it and refreshes it every 64 steps, as in FusedSGD)."""
    from . import _lib, ops
    if not ops._f16():
      return
    L = _lib.lib()
    f0, f1, _, _ = self.arena.buckets[b]
    for p in self.arena.params[f0:f1]:
      if not (p.dim() >= 2 and p.numel() >= 4096):
        continue                                               # (the narrow heads run on the CUDA cores and read fp32)
      e = ops.weight_split_buffer(p)
      age = e.get("age", 64)
      if age >= 64:
        _lib.check(L.frcnn_f16_split(_lib.ptr(p), p.numel(), _lib.ptr(e["buf"]), _lib.stream()), "frcnn_f16_split")
        _lib.count(2)
        age = 0
      else:
        _lib.check(L.frcnn_f16_split_carried(_lib.ptr(p), p.numel(), _lib.ptr(e["buf"]), self.split_ctas_per_sm, _lib.stream()), "frcnn_f16_split_carried")
        _lib.count()
      e["age"] = age + 1
      e["version"] = p._version

  def _launch_bucket(self, b):
    if self._side is None:
      # not overlapped: the buckets are only launched from step() (every gradient exists by then), back to back between ONE pair of
      # cross-rank barriers
      self._deferred.append(b)
      return
    ready = t.cuda.Event()
    ready.record()                                               # behind the bucket's last filter-gradient kernel (and every reader of its weights)
    with t.cuda.stream(self._side):
      self._side.wait_event(ready)
      self._fused(b)
    self._used_side = True

  def zero_grad(self, set_to_none = True):
    self.bytes_reduced_last_step = 0
    self._reset_step()
    for p in self.params:
      p.grad = None

  @t.no_grad()
  def step(self):
    from . import ops
    self._flush()
    if self._deferred:
      for i, b in enumerate(self._deferred):
        self._fused(b, pre_barrier = i == 0, post_barrier = False)
      if self.hG is not None:
        self.hG.barrier(channel = 0)
      for b in self._deferred:
        self._resplit(b)
      self._deferred = []
    if self._used_side:
      t.cuda.current_stream().wait_stream(self._side)            # the next forward reads the updated weights
      self._used_side = False


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
