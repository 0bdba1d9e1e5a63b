"""
Operator layer: torch.autograd.Function wrappers around the C ABI (include/frcnn_b200.h).

These are the drop-in replacements for the third-party ops the reference calls:
  conv2d_act     <- F.relu(nn.Conv2d(..., padding="same")) / t.sigmoid(conv) / conv, optionally
                    followed by nn.MaxPool2d(2,2)            (models/vgg16.py:76-96, rpn.py:88-90)
  linear_act     <- F.relu(nn.Linear) / nn.Linear            (models/vgg16.py:129-133, detector.py:76-78)
  roi_pool       <- torchvision.ops.RoIPool((7,7), 1/16)     (models/detector.py:27,72)
  softmax_rows   <- F.softmax(dim=1)                         (models/detector.py:77)
  nms / rpn_proposals / label_proposals / losses ...
Tensors keep the reference's logical shapes; 4-D activations and filters are held in
torch.channels_last memory format (physically NHWC / OHWI), which is what the kernels read.
"""
import numpy as np
import torch as t

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, check, lib, ptr, stream, workspace

import os as _os
import weakref

# "auto" (default) = the fp16 tensor-core engine wherever its shapes allow, CUDA cores elsewhere; "tf32" = the same policy with the 3xTF32
# tensor-core engine; "tc" = 3xTF32 strictly (unsupported shapes raise); "simt" = exact-fp32 CUDA cores only
_ENGINES = {"auto": _lib.ENGINE_TC_3XF16, "f16": _lib.ENGINE_TC_3XF16, "tf32": _lib.ENGINE_AUTO, "tc": _lib.ENGINE_TC_3XTF32, "simt": _lib.ENGINE_SIMT_FP32}
_engine = {"value": _ENGINES[_os.environ.get("FRCNN_ENGINE", "auto")]}


def set_engine(name):
  """'auto' = 'f16' (tcgen05 3xFP16 on scaled hi/lo splits; shapes it does not take run on the CUDA-core engine) | 'tf32' (tcgen05
  3xTF32 with the same fallback) | 'tc' (3xTF32, strict) | 'simt' (exact fp32 CUDA-core)."""
  _engine["value"] = _ENGINES[name]
  _split_cache.clear()                                             # cached operand splits are engine specific
  _weight_splits.clear()


def _f16():
  return _engine["value"] == _lib.ENGINE_TC_3XF16


def split_bytes(count):
  return lib().frcnn_f16_split_bytes(count) if _f16() else lib().frcnn_tf32_split_bytes(count)


def get_engine():
  return _engine["value"]


class _KernelTimer:
  """CUDA-event timing of the GEMM-shaped launches (bench.py's roofline leg): events are recorded on
  the launching stream around every conv/linear fwd / dgrad / wgrad call while enabled."""

  def __init__(self):
    self.enabled = False
    self.records = []

  def enable(self, on):
    self.enabled = bool(on)
    if on:
      self.records = []

  def begin(self):
    if not self.enabled:
      return None
    e = t.cuda.Event(enable_timing = True)
    e.record()
    return e

  def end(self, start, kind, gflop):
    if start is None:
      return
    e = t.cuda.Event(enable_timing = True)
    e.record()
    self.records.append((kind, gflop, start, e))

  def collect(self):
    t.cuda.synchronize()
    out = {}
    for kind, gflop, a, b in self.records:
      d = out.setdefault(kind, dict(ms = 0.0, gflop = 0.0, launches = 0))
      d["ms"] += a.elapsed_time(b); d["gflop"] += gflop; d["launches"] += 1
    self.records = []
    return out


kernel_timer = _KernelTimer()

# profiling aid: FRCNN_LAUNCH_LOG=<file> appends one line per tcgen05 GEMM launch (family, pass, GFLOP) in launch order, so an ncu
# capture filtered on tc_conv_kernel can be attributed to kernel families (tools/summarize_ncu.py)
_launch_log = open(_os.environ["FRCNN_LAUNCH_LOG"], "w") if _os.environ.get("FRCNN_LAUNCH_LOG") else None


def _engine_name(supported_tc):
  e = _engine["value"]
  if e == _lib.ENGINE_SIMT_FP32 or not supported_tc:
    return "simt_fp32"
  return "tcgen05_3xf16" if e == _lib.ENGINE_TC_3XF16 else "tcgen05_3xtf32"


def _require_cuda(*tensors):
  for x in tensors:
    if x is not None and not x.is_cuda:
      raise _lib.FrcnnError("fasterrcnn_b200 ops need CUDA tensors (no CPU path)")


def as_nhwc(x):
  """Logical (N,C,H,W) fp32 tensor -> same logical tensor, physically NHWC (own kernel)."""
  _require_cuda(x)
  assert x.dim() == 4 and x.dtype == t.float32
  n, c, h, w = x.shape
  nhwc_strides = (h * w * c, 1, w * c, c)
  if x.stride() == nhwc_strides:
    return x
  src = x.contiguous()
  if c == 1 or (h == 1 and w == 1):
    return src.as_strided((n, c, h, w), nhwc_strides)           # the two layouts coincide
  dst = t.empty((n, c, h, w), dtype = t.float32, device = x.device, memory_format = t.channels_last)
  check(lib().frcnn_nchw_to_nhwc(ptr(src), ptr(dst), n, c, h, w, stream()), "frcnn_nchw_to_nhwc")
  _lib.count()
  return dst


def as_nchw_contiguous(x):
  """Physically NHWC logical-NCHW tensor -> contiguous NCHW (own kernel)."""
  _require_cuda(x)
  if x.is_contiguous():
    return x
  n, c, h, w = x.shape
  src = as_nhwc(x)
  dst = t.empty((n, c, h, w), dtype = t.float32, device = x.device)
  check(lib().frcnn_nhwc_to_nchw(ptr(src), ptr(dst), n, c, h, w, stream()), "frcnn_nhwc_to_nchw")
  _lib.count()
  return dst


def _empty_nhwc(n, c, h, w, device):
  return t.empty((n, c, h, w), dtype = t.float32, device = device, memory_format = t.channels_last)


def _is_phys_nhwc(x):
  n, c, h, w = x.shape
  return x.stride() == (h * w * c, 1, w * c, c)


def _phys_nhwc(x):
  """Returns x with exact NHWC strides (copying through the layout kernel only if needed)."""
  return as_nhwc(x)


# ------------------------------------------------------------------------------------------------
# raw (non-autograd) launchers
# ------------------------------------------------------------------------------------------------

# ---- tf32 hi/lo operand splits shared between the passes of one step (tcgen05 engine) ---------------
_split_cache = {}
_uses_tc_cache = {}


def begin_step():
  """Drops the cached operand splits (called at the start of every forward / train_step)."""
  _split_cache.clear()
  _amax_hints.clear()


# fp16 engine: a GEMM launch can leave the per-CTA maxima of |output| behind (frcnn_conv2d_amax_slots); the later operand split of
# that output -- or of a tensor it bounds (its max-pooled map, its ReLU-masked gradient) -- then needs no pass for the maximum.
_amax_hints = {}               # (data_ptr, numel, version) -> (int32 buffer of 16 + slots words, slots, tensor kept alive)
_amax_slots_cache = {}


def _amax_slots(pass_, geom):
  r = _amax_slots_cache.get((pass_, geom))
  if r is None:
    r = int(lib().frcnn_conv2d_amax_slots(pass_, *geom))
    _amax_slots_cache[(pass_, geom)] = r
  return r


def _amax_buffer(pass_, geom, device):
  """int32 buffer for the output maxima of an fp16-engine fwd / dgrad launch of this shape, or None."""
  if not _f16() or not _uses_tc(pass_, geom):
    return None
  slots = _amax_slots(pass_, geom)
  return t.empty((16 + slots,), dtype = t.int32, device = device) if slots > 0 else None


def _set_amax_hint(x, buf):
  if buf is not None:
    _amax_hints[(x.data_ptr(), x.numel(), x._version)] = (buf, buf.numel() - 16, x)


def _amax_hint(x):
  return _amax_hints.get((x.data_ptr(), x.numel(), x._version))


def _copy_amax_hint(src, dst):
  """dst's values are bounded by src's (pooling, masking): src's maxima serve dst."""
  h = _amax_hint(src)
  if h is not None:
    _amax_hints[(dst.data_ptr(), dst.numel(), dst._version)] = (h[0], h[1], dst)


def _uses_tc(pass_, geom):
  key = (pass_, geom, _engine["value"])
  r = _uses_tc_cache.get(key)
  if r is None:
    r = bool(lib().frcnn_conv2d_uses_tensor_cores(pass_, *geom, _engine["value"]))
    _uses_tc_cache[key] = r
  return r


# Weights updated by the fused optimizer carry their operand split with them: frcnn_sgd_step_split writes the updated weights'
# [hi | lo] into a per-parameter buffer in the same pass, so the next step's GEMMs need no split pass over (for VGG-16) 547 MB of
# weights.  Entries are keyed by storage address and validated against the parameter's address / version, so a weight that was
# changed any other way (load_state_dict, .cuda(), a torch optimizer) simply falls back to the per-step split below.
_weight_splits = {}            # (data_ptr, numel) -> dict(ref = weakref(param), buf, version)


def weight_split_buffer(param):
  """Per-parameter [hi | lo] buffer for sgd_step(..., split_out =); (re)allocated when the parameter's storage moved."""
  key = (param.data_ptr(), param.numel())
  e = _weight_splits.get(key)
  if e is None or e["ref"]() is not param:
    for k in [k for k, v in _weight_splits.items() if v["ref"]() is None or v["ref"]() is param]:
      del _weight_splits[k]                                        # dead parameters and this parameter's old address
    e = dict(ref = weakref.ref(param), buf = t.empty((split_bytes(param.numel()),), dtype = t.uint8, device = param.device), version = -1)
    _weight_splits[key] = e
  return e


def pin_split(x):
  """Operand split of a long-lived tensor no optimizer updates (frozen filters: ResNet's conv1 / layer1, resnet.py:48-55): made once and kept until
  the tensor's version, its address or the engine changes -- the same table the fused optimizer's carried splits live in, so
  tf32_split() finds it.  In-place writes that bypass the version counter need invalidate_weight_splits(), as for any carried split."""
  e = weight_split_buffer(x)
  if e["version"] != x._version:
    if _f16():
      check(lib().frcnn_f16_split(ptr(x), x.numel(), ptr(e["buf"]), stream()), "frcnn_f16_split")
      _lib.count(2)
    else:
      check(lib().frcnn_tf32_split(ptr(x), x.numel(), ptr(e["buf"]), stream()), "frcnn_tf32_split")
      _lib.count()
    e["version"] = x._version
  return e["buf"]


def invalidate_weight_splits(param = None):
  """Drops the carried operand split of `param` (or of every parameter): the next GEMM that reads it re-splits from the fp32 values.
  REQUIRED after any write to a parameter that bypasses torch's version counter -- ``p.data.copy_()`` / ``.data.normal_()``,
  ``dist.broadcast(p.data)``, a raw-pointer kernel -- because the carried split is validated by address + ``_version`` only; in-place
  ops on the parameter itself (``p.copy_()``, ``load_state_dict``) bump the version and need nothing.  state.load, the sharded
  data-parallel optimizer and FusedSGD.load_state_dict call it."""
  if param is None:
    _weight_splits.clear()
    return
  for k in [k for k, v in _weight_splits.items() if v["ref"]() is param or v["ref"]() is None]:
    del _weight_splits[k]


# Gradient destinations (data parallel): the optimizer wrapper lays the weight gradients out in one flat arena, in the order backward
# produces them, and registers each parameter's slot here; the filter-gradient kernels then write straight into the arena, so a bucket of
# consecutive tensors is reduced in place (NCCL all-reduce or the fused NVLink kernel) with no gather / scatter copies.
_grad_dest = {}                # id(param) -> (weakref(param), arena (flat fp32), offset in elements)


def register_grad_destination(param, arena, offset):
  assert arena.dtype == t.float32 and arena.dim() == 1 and offset % 4 == 0 and offset + param.numel() <= arena.numel()
  _grad_dest[id(param)] = (weakref.ref(param), arena, int(offset))


def clear_grad_destinations(params = None):
  if params is None:
    _grad_dest.clear()
  else:
    for q in params:
      _grad_dest.pop(id(q), None)


def grad_destination(param):
  """The registered arena view for `param` with the given parameter's own shape / strides, or None."""
  e = _grad_dest.get(id(param))
  if e is None or e[0]() is not param:
    return None
  return e[1][e[2]:e[2] + param.numel()].as_strided(param.shape, param.stride())


def _new_grad(key, shape, device, channels_last = False):
  """Storage for a weight gradient in the layout the kernels write (row-major, or OHWI for k x k filters): a FRESH view of the
  registered arena slot (fresh, so that autograd's AccumulateGrad still sees a uniquely referenced tensor and adopts it without a copy)
  or a new tensor."""
  shape = tuple(shape)
  if channels_last:
    o, i, kh, kw = shape
    stride = (kh * kw * i, 1, kw * i, i)
  else:
    stride, acc = [], 1
    for d in reversed(shape):
      stride.append(acc); acc *= d
    stride = tuple(reversed(stride))
  e = _grad_dest.get(key) if key is not None else None
  if e is not None:
    q = e[0]()
    numel = 1
    for d in shape:
      numel *= d
    if q is not None and q.numel() == numel and e[1].device == device:
      return e[1][e[2]:e[2] + numel].as_strided(shape, stride)
  return t.empty_strided(shape, stride, dtype = t.float32, device = device)


def tf32_split(x, cache = True):
  """Returns the [hi | lo] split buffer of x (frcnn_tf32_split), computing it at most once per tensor version."""
  e = _weight_splits.get((x.data_ptr(), x.numel()))
  if e is not None:
    p = e["ref"]()
    if p is not None and p.data_ptr() == x.data_ptr() and p._version == e["version"]:
      return e["buf"]
  key = (x.data_ptr(), x.numel(), x._version)
  hit = _split_cache.get(key)
  if hit is not None:
    return hit[0]
  buf = t.empty((split_bytes(x.numel()),), dtype = t.uint8, device = x.device)
  if _f16():
    h = _amax_hint(x)
    if h is not None:
      check(lib().frcnn_f16_split_from_amax(ptr(x), x.numel(), ptr(h[0]), h[1], ptr(buf), stream()), "frcnn_f16_split_from_amax")
      _lib.count()
    else:
      check(lib().frcnn_f16_split(ptr(x), x.numel(), ptr(buf), stream()), "frcnn_f16_split")
      _lib.count(2)
  else:
    check(lib().frcnn_tf32_split(ptr(x), x.numel(), ptr(buf), stream()), "frcnn_tf32_split")
    _lib.count()
  if cache:
    _split_cache[key] = (buf, x)                 # holding x keeps its address from being recycled while cached
  return buf


def drop_split(x):
  _split_cache.pop((x.data_ptr(), x.numel(), x._version), None)


_gemm_workspace_cache = {}


def _gemm(pass_, a, b, out, geom, kind, gflop, a_split = None, b_split = None, scale = None, bias = None, residual = None, act = ACT_NONE, addend = None, amax_out = None):
  """One implicit-GEMM launch.  pass 0: a=x, b=w, out=y | pass 1: a=dy, b=w, out=dx | pass 2: a=dy, b=x, out=dw."""
  eng = _engine["value"]
  L = lib()
  f16 = _f16() and _uses_tc(pass_, geom)
  presplit = (a_split is not None or b_split is not None) or f16
  wkey = (pass_, geom, eng)
  need = _gemm_workspace_cache.get(wkey)
  if need is None:
    need = (L.frcnn_conv2d_fwd_workspace_bytes, L.frcnn_conv2d_dgrad_workspace_bytes, L.frcnn_conv2d_wgrad_workspace_bytes)[pass_](*geom, eng)
    _gemm_workspace_cache[wkey] = need
  if pass_ == 0:
    ws, ws_n = workspace(need)
    t0 = kernel_timer.begin()
    if presplit:
      if f16:
        check(L.frcnn_conv2d_fwd_f16(ptr(a), ptr(b), ptr(a_split), ptr(b_split), ptr(scale), ptr(bias), ptr(residual), ptr(out), *geom, act, ptr(amax_out), ws, ws_n, stream()), "frcnn_conv2d_fwd_f16")
      else:
        check(L.frcnn_conv2d_fwd_presplit(ptr(a), ptr(b), ptr(a_split), ptr(b_split), ptr(scale), ptr(bias), ptr(residual), ptr(out), *geom, act, ws, ws_n, stream()), "frcnn_conv2d_fwd_presplit")
    else:
      check(L.frcnn_conv2d_fwd(ptr(a), ptr(b), ptr(scale), ptr(bias), ptr(residual), ptr(out), *geom, act, eng, ws, ws_n, stream()), "frcnn_conv2d_fwd")
  elif pass_ == 1:
    ws, ws_n = workspace(need)
    t0 = kernel_timer.begin()
    if presplit:
      if f16:
        check(L.frcnn_conv2d_dgrad_f16(ptr(a), ptr(b), ptr(a_split), ptr(b_split), ptr(addend), ptr(out), *geom, ptr(amax_out), ws, ws_n, stream()), "frcnn_conv2d_dgrad_f16")
      else:
        check(L.frcnn_conv2d_dgrad_presplit(ptr(a), ptr(b), ptr(a_split), ptr(b_split), ptr(addend), ptr(out), *geom, ws, ws_n, stream()), "frcnn_conv2d_dgrad_presplit")
    else:
      check(L.frcnn_conv2d_dgrad(ptr(a), ptr(b), ptr(addend), ptr(out), *geom, eng, ws, ws_n, stream()), "frcnn_conv2d_dgrad")
  else:
    ws, ws_n = workspace(need)
    t0 = kernel_timer.begin()
    if presplit:
      check((L.frcnn_conv2d_wgrad_f16 if f16 else L.frcnn_conv2d_wgrad_presplit)(ptr(a), ptr(b), ptr(a_split), ptr(b_split), ptr(out), *geom, ws, ws_n, stream()), "frcnn_conv2d_wgrad_presplit")
    else:
      check(L.frcnn_conv2d_wgrad(ptr(a), ptr(b), ptr(out), *geom, eng, ws, ws_n, stream()), "frcnn_conv2d_wgrad")
  kernel_timer.end(t0, kind, gflop)
  if _launch_log is not None and presplit:                         # presplit <=> the tcgen05 engine took this launch
    _launch_log.write("%s %d %.6f\n" % (kind, pass_, gflop))
    _launch_log.flush()
  _lib.count()


def conv2d_fwd_raw(x, w, bias, stride, pad, act, scale = None, residual = None, reuse_x = False, reuse_w = True):
  """x logical (N,Cin,H,W) phys NHWC; w logical (Cout,Cin,KH,KW) phys OHWI -> y logical (N,Cout,Ho,Wo) phys NHWC.
  reuse_x / reuse_w: the operand will be used by another pass of this step -> its tf32 split is cached."""
  n, cin, h, wd = x.shape
  cout, cin2, kh, kw = w.shape
  assert cin == cin2
  ho = (h + 2 * pad - kh) // stride + 1
  wo = (wd + 2 * pad - kw) // stride + 1
  y = _empty_nhwc(n, cout, ho, wo, x.device)
  geom = (n, h, wd, cin, cout, kh, kw, stride, pad)
  xs = ws_ = None
  if _uses_tc(0, geom):
    # the activation's split is cached only when another pass of this step reads it again (the filter gradient); either way it is made
    # here, outside the GEMM call, so that the per-launch timing of the roofline leg covers the GEMM kernel alone
    xs = tf32_split(x, cache = reuse_x)
    if reuse_w:
      ws_ = tf32_split(w)
  amax = _amax_buffer(0, geom, x.device)
  _gemm(0, x, w, y, geom, "conv_fwd", 2e-9 * n * ho * wo * cout * kh * kw * cin, xs, ws_, scale = scale, bias = bias, residual = residual, act = act, amax_out = amax)
  _set_amax_hint(y, amax)
  return y


def conv2d_dgrad_raw(dy, w, x_shape, stride, pad, addend = None, reuse_dy = False, dy_split = None):
  n, cin, h, wd = x_shape
  cout, _, kh, kw = w.shape
  dx = _empty_nhwc(n, cin, h, wd, dy.device)
  geom = (n, h, wd, cin, cout, kh, kw, stride, pad)
  ds = ws_ = None
  if _uses_tc(1, geom):
    ws_ = tf32_split(w)
    if dy_split is not None:
      ds = dy_split
    elif reuse_dy:
      ds = tf32_split(dy)
  amax = _amax_buffer(1, geom, dy.device)
  _gemm(1, dy, w, dx, geom, "conv_dgrad", 2e-9 * dy.shape[0] * dy.shape[2] * dy.shape[3] * cout * kh * kw * cin, ds, ws_, addend = addend, amax_out = amax)
  _set_amax_hint(dx, amax)
  return dx


def conv2d_wgrad_raw(dy, x, w_shape, stride, pad, dy_split = None, w_key = None):
  n, cin, h, wd = x.shape
  cout, _, kh, kw = w_shape
  dw = _new_grad(w_key, (cout, cin, kh, kw), x.device, channels_last = not (kh == 1 and kw == 1))
  geom = (n, h, wd, cin, cout, kh, kw, stride, pad)
  ds = xs = None
  if _uses_tc(2, geom):
    key_dy = (dy.data_ptr(), dy.numel(), dy._version)
    key_x = (x.data_ptr(), x.numel(), x._version)
    ds = dy_split if dy_split is not None else (_split_cache[key_dy][0] if key_dy in _split_cache else None)
    xs = _split_cache[key_x][0] if key_x in _split_cache else None
  _gemm(2, dy, x, dw, geom, "conv_wgrad", 2e-9 * dy.shape[0] * dy.shape[2] * dy.shape[3] * cout * kh * kw * cin, ds, xs)
  drop_split(dy)
  drop_split(x)
  return dw


def bias_grad_raw(dz_rows, c):
  """dz_rows: any tensor whose memory is (rows, C) row-major."""
  rows = dz_rows.numel() // c
  db = t.empty((c,), dtype = t.float32, device = dz_rows.device)
  ws_bytes = lib().frcnn_bias_grad_workspace_bytes(rows, c)
  ws, ws_n = workspace(ws_bytes, slot = 1)
  check(lib().frcnn_bias_grad(ptr(dz_rows), ptr(db), rows, c, ws, ws_n, stream()), "frcnn_bias_grad")
  _lib.count(2)
  return db


def _act_bwd(dy, y, act, c, want_split, want_bias, need_fp32):
  """Backward of the fused epilogue y = act(z + b) over a (rows, c) row-major gradient.  Returns (dz, dz_split, db):
  dz fp32 tensor (when need_fp32 is False and the fused kernel runs it is only a placeholder carrying shape and a valid address:
  the tcgen05 GEMMs read dz_split), dz_split = [hi | lo] operand buffer or None, db = bias gradient or None.
  One frcnn_act_bwd_fused launch when the channel count allows it, else the separate kernels."""
  rows = dy.numel() // c
  L = lib()
  f16 = _f16()
  if act in (ACT_NONE, ACT_RELU) and rows > 0 and (want_split or want_bias) and L.frcnn_act_bwd_fused_supported(rows, c):
    write_dz = act == ACT_RELU and (need_fp32 or not want_split)
    dz = t.empty_like(y) if write_dz else None
    split = t.empty((split_bytes(dy.numel()),), dtype = t.uint8, device = dy.device) if want_split else None
    db = t.empty((c,), dtype = t.float32, device = dy.device) if want_bias else None
    ws, ws_n = workspace(L.frcnn_act_bwd_fused_workspace_bytes(rows, c), slot = 1) if want_bias else (None, 0)
    if f16:
      h = _amax_hint(dy) if want_split else None
      check(L.frcnn_act_bwd_fused_f16(ptr(dy), ptr(y) if act == ACT_RELU else None, act, ptr(dz), ptr(split), ptr(db), rows, c,
                                      ptr(h[0]) if h is not None else None, h[1] if h is not None else 0, ws, ws_n, stream()), "frcnn_act_bwd_fused_f16")
    else:
      check(L.frcnn_act_bwd_fused(ptr(dy), ptr(y) if act == ACT_RELU else None, act, ptr(dz), ptr(split), ptr(db), rows, c, ws, ws_n, stream()), "frcnn_act_bwd_fused")
    _lib.count((2 if want_bias else 1) + (1 if f16 and want_split and h is None else 0))
    if dz is None and act == ACT_NONE:
      dz = dy
    elif dz is None:                                             # placeholder over the hi half, in y's physical (channels-last) layout
      flat = (split[4096:4096 + dy.numel() * 4] if f16 else split[:dy.numel() * 4]).view(t.float32)     # (fp16 layout: hi + lo halves = 4 B / element)
      dz = flat.view(y.shape[0], y.shape[2], y.shape[3], y.shape[1]).permute(0, 3, 1, 2) if y.dim() == 4 else flat.view(y.shape)
    return dz, split, db
  if act == ACT_RELU:
    dz = t.empty_like(y)
    check(L.frcnn_relu_bwd(ptr(dy), ptr(y), ptr(dz), y.numel(), stream()), "frcnn_relu_bwd")
    _lib.count()
  elif act == ACT_SIGMOID:
    dz = t.empty_like(y)
    check(L.frcnn_sigmoid_bwd(ptr(dy), ptr(y), ptr(dz), y.numel(), stream()), "frcnn_sigmoid_bwd")
    _lib.count()
  else:
    dz = dy
  db = bias_grad_raw(dz, c) if want_bias and rows > 0 else None
  return dz, None, db


def _phys_filter(w):
  """Filter in OHWI physical layout (channels_last of the logical OIHW tensor)."""
  o, i, kh, kw = w.shape
  if kh == 1 and kw == 1:
    return w if w.is_contiguous() else w.contiguous()
  if w.stride() == (kh * kw * i, 1, kw * i, i):
    return w
  return as_nhwc(w.detach().contiguous())


# ------------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------------

def _fused_bwd_ok(geom, want_dx, want_dw, tc_dx, tc_dw, act, rows, c):
  """One-call layer backward (frcnn_conv2d_bwd_f16): fp16 engine, both gradients on the tensor cores (or dx not wanted), an activation the
  fused kernel handles; the per-launch timer of the roofline leg keeps the separate calls."""
  if not _f16() or kernel_timer.enabled or _launch_log is not None or not want_dw or not tc_dw or (want_dx and not tc_dx):
    return False
  if act not in (ACT_NONE, ACT_RELU) or rows == 0 or geom[7] != 1:
    return False
  return bool(lib().frcnn_act_bwd_fused_supported(rows, c))


def _conv_bwd_fused(dy, y, act, pool, xp, wp, geom, w_shape, want_dx, want_bias, w_key = None):
  """-> (dx, dw, db) of one conv / linear layer through frcnn_conv2d_bwd_f16.  dy / y / xp are NHWC-physical maps or (rows, C) matrices."""
  L = lib()
  dev = y.device
  c = geom[4]
  rows = y.numel() // c
  dz_full = t.empty_like(y) if pool else None
  split = t.empty((split_bytes(y.numel()),), dtype = t.uint8, device = dev)
  db = t.empty((c,), dtype = t.float32, device = dev) if want_bias else None
  if y.dim() == 4:
    dx = _empty_nhwc(*xp.shape, dev) if want_dx else None
    cout, cin, kh, kw = w_shape
    dw = _new_grad(w_key, (cout, cin, kh, kw), dev, channels_last = not (kh == 1 and kw == 1))
  else:
    dx = t.empty(tuple(xp.shape), dtype = t.float32, device = dev) if want_dx else None
    dw = _new_grad(w_key, tuple(w_shape), dev)
  h = _amax_hint(dy)
  dx_amax = _amax_buffer(1, geom, dev) if want_dx else None
  w_split = tf32_split(wp)
  kx = (xp.data_ptr(), xp.numel(), xp._version)
  x_split = _split_cache[kx][0] if kx in _split_cache else None
  eng = _engine["value"]
  wkey = ("bwd", geom, eng)
  need = _gemm_workspace_cache.get(wkey)
  if need is None:
    need = max(L.frcnn_conv2d_dgrad_workspace_bytes(*geom, eng), L.frcnn_conv2d_wgrad_workspace_bytes(*geom, eng))
    _gemm_workspace_cache[wkey] = need
  ws, ws_n = workspace(need)
  bws, bws_n = workspace(L.frcnn_act_bwd_fused_workspace_bytes(rows, c), slot = 1) if want_bias else (None, 0)
  check(L.frcnn_conv2d_bwd_f16(ptr(dy), ptr(y), act, int(bool(pool)), ptr(xp), ptr(x_split), ptr(wp), ptr(w_split), ptr(h[0]) if h is not None else None, h[1] if h is not None else 0,
                               ptr(dz_full), ptr(split), ptr(db), ptr(dx), ptr(dx_amax), ptr(dw), *geom, bws, bws_n, ws, ws_n, stream()), "frcnn_conv2d_bwd_f16")
  _lib.count(3 + (1 if pool else 0) + (1 if h is None else 0) + (1 if want_bias else 0) + (0 if want_dx else -1))
  if dx is not None:
    _set_amax_hint(dx, dx_amax)
  drop_split(xp)
  return dx, dw, db


class _ConvAct(t.autograd.Function):
  """y = [maxpool2x2] act(conv2d(x, w) + b).  Backward: fused (pool+)activation backward, then
  dgrad / wgrad / bias-grad kernels."""

  @staticmethod
  def forward(ctx, x, w, b, stride, pad, act, pool):
    _require_cuda(x, w, b)
    xp = _phys_nhwc(x.detach())
    wp = _phys_filter(w.detach())
    y = conv2d_fwd_raw(xp, wp, b.detach() if b is not None else None, stride, pad, act, reuse_x = bool(ctx.needs_input_grad[1]))
    ctx.stride, ctx.pad, ctx.act, ctx.pool = stride, pad, act, pool
    ctx.has_bias = b is not None
    ctx.w_shape = tuple(w.shape)
    ctx.w_key = id(w)                                                       # the parameter object: its gradient may have a registered destination
    if pool:
      assert act == ACT_RELU, "fused pooling is defined for the ReLU convs of VGG-16"
      n, c, h, wd = y.shape
      yp = _empty_nhwc(n, c, h // 2, wd // 2, y.device)
      check(lib().frcnn_maxpool2x2_fwd(ptr(y), ptr(yp), n, h, wd, c, stream()), "frcnn_maxpool2x2_fwd")
      _lib.count()
      _copy_amax_hint(y, yp)
      ctx.save_for_backward(xp, wp, y)
      return yp
    ctx.save_for_backward(xp, wp, y)
    return y

  @staticmethod
  def backward(ctx, dy):
    xp, wp, y = ctx.saved_tensors
    n, c, h, wd = y.shape
    dy = _phys_nhwc(dy)
    geom = (xp.shape[0], xp.shape[2], xp.shape[3], xp.shape[1], c, ctx.w_shape[2], ctx.w_shape[3], ctx.stride, ctx.pad)
    want_dx, want_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    tc_dx, tc_dw = want_dx and _uses_tc(1, geom), want_dw and _uses_tc(2, geom)
    need_fp32 = (want_dx and not tc_dx) or (want_dw and not tc_dw)          # a CUDA-core GEMM reads dz itself
    want_bias = ctx.has_bias and ctx.needs_input_grad[2]
    act = ctx.act
    if _fused_bwd_ok(geom, want_dx, want_dw, tc_dx, tc_dw, act, y.numel() // c, c):
      return _conv_bwd_fused(dy, y, act, ctx.pool, xp, wp, geom, ctx.w_shape, want_dx, want_bias, ctx.w_key) + (None, None, None, None)
    if ctx.pool:
      dz = t.empty_like(y)
      check(lib().frcnn_maxpool2x2_relu_bwd(ptr(dy), ptr(y), ptr(dz), n, h, wd, c, stream()), "frcnn_maxpool2x2_relu_bwd")
      _lib.count()
      _copy_amax_hint(dy, dz)
      dy, act = dz, ACT_NONE                                                # ReLU mask already applied by the pooling backward
    dz, dz_split, db = _act_bwd(dy, y, act, c, tc_dx or tc_dw, want_bias, need_fp32)
    dx = dw = None
    if want_dx:
      dx = conv2d_dgrad_raw(dz, wp, tuple(xp.shape), ctx.stride, ctx.pad, reuse_dy = want_dw, dy_split = dz_split if tc_dx else None)
    if want_dw:
      dw = conv2d_wgrad_raw(dz, xp, ctx.w_shape, ctx.stride, ctx.pad, dy_split = dz_split if tc_dw else None, w_key = ctx.w_key)
    return dx, dw, db, None, None, None, None


class _NoGradCtx:
  """Stand-in for the autograd context when nothing upstream needs a gradient (frozen VGG-16 blocks 1-2, inference)."""
  needs_input_grad = (False,) * 8
  w_key = None

  def save_for_backward(self, *tensors):
    pass


def conv2d_act(x, weight, bias, stride = 1, pad = 1, act = ACT_RELU, pool = False):
  if not t.is_grad_enabled() or not (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
    return _ConvAct.forward(_NoGradCtx(), x, weight, bias, stride, pad, act, pool)      # no autograd node, no saved activations
  return _ConvAct.apply(x, weight, bias, stride, pad, act, pool)


class _LinearAct(t.autograd.Function):
  """y = act(x @ w.T + b): the KH=KW=1, H=W=1 case of the implicit-GEMM kernels."""

  @staticmethod
  def forward(ctx, x, w, b, act):
    _require_cuda(x, w, b)
    x2 = x.detach().contiguous()
    w2 = w.detach().contiguous()
    m, k = x2.shape
    nout = w2.shape[0]
    y = t.empty((m, nout), dtype = t.float32, device = x.device)
    ctx.act = act
    ctx.has_bias = b is not None
    ctx.w_key = id(w)
    if m > 0:
      geom = (m, 1, 1, k, nout, 1, 1, 1, 0)
      xs = ws_ = None
      if _uses_tc(0, geom):
        ws_ = tf32_split(w2)
        if ctx.needs_input_grad[1]:
          xs = tf32_split(x2)
      amax = _amax_buffer(0, geom, x2.device)
      _gemm(0, x2, w2, y, geom, "linear_fwd", 2e-9 * m * k * nout, xs, ws_, bias = b.detach() if b is not None else None, act = act, amax_out = amax)
      _set_amax_hint(y, amax)
    ctx.save_for_backward(x2, w2, y)
    return y

  @staticmethod
  def backward(ctx, dy):
    x2, w2, y = ctx.saved_tensors
    m, k = x2.shape
    nout = w2.shape[0]
    dy = dy.contiguous()
    if m == 0:
      return t.zeros_like(x2), t.zeros_like(w2), (t.zeros((nout,), device = x2.device) if ctx.has_bias else None), None
    geom = (m, 1, 1, k, nout, 1, 1, 1, 0)
    want_dx, want_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    tc_dx, tc_dw = want_dx and _uses_tc(1, geom), want_dw and _uses_tc(2, geom)
    if _fused_bwd_ok(geom, want_dx, want_dw, tc_dx, tc_dw, ctx.act, m, nout):
      return _conv_bwd_fused(dy, y, ctx.act, False, x2, w2, geom, tuple(w2.shape), want_dx, ctx.has_bias and ctx.needs_input_grad[2], ctx.w_key) + (None,)
    need_fp32 = (want_dx and not tc_dx) or (want_dw and not tc_dw)
    dz, dz_split, db = _act_bwd(dy, y, ctx.act, nout, tc_dx or tc_dw, ctx.has_bias and ctx.needs_input_grad[2], need_fp32)
    dx = dw = None
    if want_dx:
      dx = t.empty((m, k), dtype = t.float32, device = x2.device)
      ds = ws_ = None
      if tc_dx:
        ws_ = tf32_split(w2)
        ds = dz_split if dz_split is not None else (tf32_split(dz) if want_dw else None)
      amax = _amax_buffer(1, geom, x2.device)
      _gemm(1, dz, w2, dx, geom, "linear_dgrad", 2e-9 * m * k * nout, ds, ws_, amax_out = amax)
      _set_amax_hint(dx, amax)
    if want_dw:
      dw = _new_grad(ctx.w_key, (nout, k), x2.device)
      ds = xs = None
      if tc_dw:
        kd, kx = (dz.data_ptr(), dz.numel(), dz._version), (x2.data_ptr(), x2.numel(), x2._version)
        ds = dz_split if dz_split is not None else (_split_cache[kd][0] if kd in _split_cache else None)
        xs = _split_cache[kx][0] if kx in _split_cache else None
      _gemm(2, dz, x2, dw, geom, "linear_wgrad", 2e-9 * m * k * nout, ds, xs)
      drop_split(dz)
      drop_split(x2)
    return dx, dw, db, None


def linear_act(x, weight, bias, act = ACT_NONE):
  return _LinearAct.apply(x, weight, bias, act)


class _TwoHeads(t.autograd.Function):
  """(y1, y2) = (act1(x w1^T + b1), act2(x w2^T + b2)) for two narrow heads sharing the input rows x (M, K): frcnn_heads_fwd / _bwd.
  x may be a 4-D NHWC-physical activation (rows = pixels) or a 2-D matrix; outputs are (M, N1) / (M, N2) row-major."""

  @staticmethod
  def forward(ctx, x, w1, b1, act1, w2, b2, act2):
    _require_cuda(x, w1, w2)
    xp = _phys_nhwc(x.detach()) if x.dim() == 4 else x.detach().contiguous()
    k = xp.shape[1]
    m = xp.numel() // k
    n1, n2 = w1.shape[0], w2.shape[0]
    y1 = t.empty((m, n1), dtype = t.float32, device = x.device)
    y2 = t.empty((m, n2), dtype = t.float32, device = x.device)
    L = lib()
    if m > 0:
      ws, ws_n = workspace(L.frcnn_heads_workspace_bytes(m, k, n1, n2), slot = 1)
      check(L.frcnn_heads_fwd(ptr(xp), m, k, ptr(w1.detach()), ptr(b1.detach()) if b1 is not None else None, n1, act1,
                              ptr(w2.detach()), ptr(b2.detach()) if b2 is not None else None, n2, act2, ptr(y1), ptr(y2), ws, ws_n, stream()), "frcnn_heads_fwd")
      _lib.count(2)
    ctx.acts = (act1, act2)
    ctx.w_keys = (id(w1), id(w2))
    ctx.has_bias = (b1 is not None, b2 is not None)
    ctx.x_is_map = x.dim() == 4
    ctx.save_for_backward(xp, w1.detach(), w2.detach(), y1, y2)
    return y1, y2

  @staticmethod
  def backward(ctx, dy1, dy2):
    xp, w1, w2, y1, y2 = ctx.saved_tensors
    k = xp.shape[1]
    m = xp.numel() // k
    n1, n2 = w1.shape[0], w2.shape[0]
    dev = xp.device
    dw1 = _new_grad(ctx.w_keys[0], w1.shape, dev) if w1.is_contiguous() else t.empty_strided(w1.shape, w1.stride(), dtype = t.float32, device = dev)
    dw2 = _new_grad(ctx.w_keys[1], w2.shape, dev) if w2.is_contiguous() else t.empty_strided(w2.shape, w2.stride(), dtype = t.float32, device = dev)
    db1 = t.empty((n1,), dtype = t.float32, device = dev) if ctx.has_bias[0] else None
    db2 = t.empty((n2,), dtype = t.float32, device = dev) if ctx.has_bias[1] else None
    dx = None
    if m == 0:
      return (t.zeros_like(xp) if ctx.needs_input_grad[0] else None), dw1.zero_(), (db1.zero_() if db1 is not None else None), None, dw2.zero_(), (db2.zero_() if db2 is not None else None), None
    if ctx.needs_input_grad[0]:
      dx = t.empty_like(xp)                                   # same physical layout as x (NHWC rows / matrix)
    dy1 = dy1.contiguous() if dy1 is not None else t.zeros_like(y1)
    dy2 = dy2.contiguous() if dy2 is not None else t.zeros_like(y2)
    L = lib()
    ws, ws_n = workspace(L.frcnn_heads_workspace_bytes(m, k, n1, n2), slot = 1)
    check(L.frcnn_heads_bwd(ptr(xp), m, k, ptr(w1), n1, ctx.acts[0], ptr(y1), ptr(dy1), ptr(w2), n2, ctx.acts[1], ptr(y2), ptr(dy2),
                            ptr(dx), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ws, ws_n, stream()), "frcnn_heads_bwd")
    _lib.count(4)
    return dx, dw1, db1, None, dw2, db2, None


def two_heads(x, w1, b1, act1, w2, b2, act2):
  return _TwoHeads.apply(x, w1, b1, act1, w2, b2, act2)


class _RoIPool(t.autograd.Function):
  @staticmethod
  def forward(ctx, feature_map, proposals, output_size, spatial_scale):
    _require_cuda(feature_map, proposals)
    assert feature_map.shape[0] == 1, "Batch size must be 1"
    fm = _phys_nhwc(feature_map.detach())
    _, c, h, w = fm.shape
    props = proposals.detach().contiguous().float()
    k = props.shape[0]
    ph, pw = output_size
    out = t.empty((k, c, ph, pw), dtype = t.float32, device = fm.device)
    arg = t.empty((k, ph * pw, c), dtype = t.int32, device = fm.device)      # bin-major: private to the forward / backward kernel pair
    if k > 0:
      check(lib().frcnn_roi_pool_fwd(ptr(fm), h, w, c, ptr(props), k, ph, pw, float(spatial_scale), ptr(out), ptr(arg), stream()), "frcnn_roi_pool_fwd")
      _lib.count()
    ctx.save_for_backward(arg, props)
    ctx.geom = (k, h, w, c, ph, pw, float(spatial_scale))
    return out

  @staticmethod
  def backward(ctx, dout):
    arg, props = ctx.saved_tensors
    k, h, w, c, ph, pw, scale = ctx.geom
    dfm = _empty_nhwc(1, c, h, w, dout.device)
    dout = dout.contiguous()
    if k == 0:
      return dfm.zero_(), None, None, None
    check(lib().frcnn_roi_pool_bwd(ptr(dout), ptr(arg), ptr(props), k, h, w, c, ph, pw, scale, None, ptr(dfm), stream()), "frcnn_roi_pool_bwd")
    _lib.count()
    return dfm, None, None, None


def roi_pool(feature_map, proposals, output_size = (7, 7), spatial_scale = 1.0 / 16.0):
  """feature_map logical (1,C,H,W); proposals (K,4) as (y1,x1,y2,x2) -> (K,C,7,7)."""
  return _RoIPool.apply(feature_map, proposals, output_size, spatial_scale)


class _RoIAlign(t.autograd.Function):
  """EXTENSION: torchvision.ops.roi_align semantics (fixed sampling_ratio, aligned flag); the reference has RoIPool only."""

  @staticmethod
  def forward(ctx, feature_map, proposals, output_size, spatial_scale, sampling_ratio, aligned):
    _require_cuda(feature_map, proposals)
    assert feature_map.shape[0] == 1, "Batch size must be 1"
    fm = as_nhwc(feature_map.detach())
    _, c, h, w = fm.shape
    props = proposals.detach().contiguous().float()
    k = props.shape[0]
    ph, pw = output_size
    out = t.empty((k, c, ph, pw), dtype = t.float32, device = fm.device)
    if k > 0:
      check(lib().frcnn_roi_align_fwd(ptr(fm), h, w, c, ptr(props), k, ph, pw, float(spatial_scale), int(sampling_ratio), int(bool(aligned)), ptr(out), stream()), "frcnn_roi_align_fwd")
      _lib.count()
    ctx.save_for_backward(props)
    ctx.geom = (k, h, w, c, ph, pw, float(spatial_scale), int(sampling_ratio), int(bool(aligned)))
    return out

  @staticmethod
  def backward(ctx, dout):
    (props,) = ctx.saved_tensors
    k, h, w, c, ph, pw, scale, s, al = ctx.geom
    dfm = _empty_nhwc(1, c, h, w, dout.device)
    if k == 0:
      return dfm.zero_(), None, None, None, None, None
    check(lib().frcnn_roi_align_bwd(ptr(dout.contiguous()), ptr(props), k, h, w, c, ph, pw, scale, s, al, None, ptr(dfm), stream()), "frcnn_roi_align_bwd")
    _lib.count()
    return dfm, None, None, None, None, None


def roi_align(feature_map, proposals, output_size = (7, 7), spatial_scale = 1.0 / 16.0, sampling_ratio = 2, aligned = False):
  """feature_map logical (1,C,H,W); proposals (K,4) as (y1,x1,y2,x2) -> (K,C,7,7).  Extension (no reference counterpart)."""
  return _RoIAlign.apply(feature_map, proposals, output_size, spatial_scale, sampling_ratio, aligned)


class _Softmax(t.autograd.Function):
  @staticmethod
  def forward(ctx, logits):
    _require_cuda(logits)
    x = logits.detach().contiguous()
    n, c = x.shape
    p = t.empty_like(x)
    if n > 0:
      check(lib().frcnn_softmax_rows(ptr(x), ptr(p), n, c, stream()), "frcnn_softmax_rows")
      _lib.count()
    ctx.save_for_backward(p)
    return p

  @staticmethod
  def backward(ctx, g):
    (p,) = ctx.saved_tensors
    n, c = p.shape
    g = g.contiguous()
    d = t.empty_like(p)
    if n > 0:
      check(lib().frcnn_softmax_rows_bwd(ptr(p), ptr(g), ptr(d), n, c, stream()), "frcnn_softmax_rows_bwd")
      _lib.count()
    return d


def softmax_rows(logits):
  return _Softmax.apply(logits)


class _RPNLosses(t.autograd.Function):
  """(class_loss, regression_loss) of models/rpn.py:176-272 in one kernel, gradients included."""

  @staticmethod
  def forward(ctx, scores, deltas, y_true):
    _require_cuda(scores, deltas, y_true)
    s = scores.detach().contiguous()
    d = deltas.detach().contiguous()
    y = y_true.detach().contiguous()
    a = s.numel()
    assert d.numel() == 4 * a and y.numel() == 6 * a
    out = t.empty((2,), dtype = t.float32, device = s.device)
    need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
    ds = t.empty_like(s) if need else None
    dd = t.empty_like(d) if need else None
    check(lib().frcnn_rpn_losses(ptr(s), ptr(d), ptr(y), a, ptr(out), ptr(ds), ptr(dd), stream()), "frcnn_rpn_losses")
    _lib.count()
    ctx.grads = (ds, dd)
    ctx.shapes = (scores.shape, deltas.shape)
    return out

  @staticmethod
  def backward(ctx, g):
    ds, dd = ctx.grads
    return (ds * g[0]).reshape(ctx.shapes[0]), (dd * g[1]).reshape(ctx.shapes[1]), None


def rpn_losses(scores, deltas, y_true):
  return _RPNLosses.apply(scores, deltas, y_true)


class _DetectorLosses(t.autograd.Function):
  """(class_loss, regression_loss) of models/detector.py:83-155 in one kernel."""

  @staticmethod
  def forward(ctx, probs, deltas, y_classes, y_deltas):
    _require_cuda(probs, deltas, y_classes, y_deltas)
    p = probs.detach().contiguous()
    d = deltas.detach().contiguous()
    yc = y_classes.detach().contiguous()
    yd = y_deltas.detach().contiguous()
    n, c = p.shape
    out = t.zeros((2,), dtype = t.float32, device = p.device)
    dp = t.empty_like(p)
    dd = t.empty_like(d)
    if n > 0:
      check(lib().frcnn_detector_losses(ptr(p), ptr(d), ptr(yc), ptr(yd), n, c, ptr(out), ptr(dp), ptr(dd), stream()), "frcnn_detector_losses")
      _lib.count()
    ctx.grads = (dp, dd)
    return out

  @staticmethod
  def backward(ctx, g):
    dp, dd = ctx.grads
    return dp * g[0], dd * g[1], None, None


def detector_losses(probs, deltas, y_classes, y_deltas):
  return _DetectorLosses.apply(probs, deltas, y_classes, y_deltas)


# ------------------------------------------------------------------------------------------------
# non-differentiable ops
# ------------------------------------------------------------------------------------------------

def nms(boxes, scores, iou_threshold):
  """torchvision.ops.nms replacement (models/rpn.py:147-151: fp32 boxes; models/faster_rcnn.py:216-220: float64 boxes with float32
  scores -- IoU then runs in IEEE double, frcnn_nms_sorted_f64): returns int64 indices of the kept boxes in descending score order
  (ties: lower index first)."""
  _require_cuda(boxes, scores)
  n = boxes.shape[0]
  if n == 0:
    return t.empty((0,), dtype = t.int64, device = boxes.device)
  f64 = boxes.dtype == t.float64
  b = boxes.detach().contiguous() if f64 else boxes.detach().contiguous().float()
  s = scores.detach().contiguous().float()
  dev = b.device
  # stable descending order == rank counting with ties -> LOWER index first: negate the tie rule
  # by ranking the reversed array (ties -> higher index first in reversed space).
  rev = t.flip(s, dims = (0,)).contiguous()
  order = t.empty((n,), dtype = t.int32, device = dev)
  cnt = t.empty((1 + n,), dtype = t.int32, device = dev)
  check(lib().frcnn_topk_order(ptr(rev), None, n, n, ptr(order), ptr(cnt), stream()), "frcnn_topk_order")
  _lib.count(3)
  order = (n - 1) - order                                        # back to original indices
  sorted_boxes = t.empty((n, 4), dtype = b.dtype, device = dev)
  # (a float64 row is gathered as eight 32-bit words)
  check(lib().frcnn_gather_rows_f32(ptr(b), 8 if f64 else 4, ptr(order.contiguous()), ptr(cnt), n, ptr(sorted_boxes), stream()), "frcnn_gather_rows_f32")
  keep = t.empty((n,), dtype = t.int32, device = dev)
  kept = t.empty((1,), dtype = t.int32, device = dev)
  ws, ws_n = workspace(lib().frcnn_nms_workspace_bytes(n), slot = 2)
  entry = lib().frcnn_nms_sorted_f64 if f64 else lib().frcnn_nms_sorted_f32
  check(entry(ptr(sorted_boxes), ptr(cnt), n, float(iou_threshold), n, ptr(keep), ptr(kept), ws, ws_n, stream()), "frcnn_nms_sorted_f64" if f64 else "frcnn_nms_sorted_f32")
  _lib.count(3)
  k = int(kept.item())
  return order[keep[:k].long()].long()


def nms_batched(boxes, scores, iou_threshold, max_keep = None):
  """B independent NMS problems in one stream-ordered sequence (per-class NMS at BASELINE config 5 sizes): boxes (B,n,4) fp32,
  scores (B,n) fp32, unsorted.  Returns (keep (B, max_keep) int32 original indices in score order, -1 padded; counts (B) int32),
  both on the device -- no host synchronisation."""
  _require_cuda(boxes, scores)
  bsz, n = int(scores.shape[0]), int(scores.shape[1])
  b = boxes.detach().contiguous().float()
  s = scores.detach().contiguous().float()
  cap = n if max_keep is None else min(int(max_keep), n)
  keep = t.empty((bsz, cap), dtype = t.int32, device = b.device)
  counts = t.empty((bsz,), dtype = t.int32, device = b.device)
  ws, ws_n = workspace(lib().frcnn_nms_batched_workspace_bytes(bsz, n, cap), slot = 2)
  check(lib().frcnn_nms_batched_f32(ptr(b), ptr(s), bsz, n, float(iou_threshold), cap, ptr(keep), ptr(counts), ws, ws_n, stream()), "frcnn_nms_batched_f32")
  _lib.count(6)
  return keep, counts


class ProposalBuffers:
  """Capacity-sized device buffers for the RPN proposal path (reused across steps)."""

  def __init__(self, num_anchors, pre_nms, device):
    self.num_anchors, self.pre_nms = num_anchors, pre_nms
    i32, f32, u8 = t.int32, t.float32, t.uint8
    self.boxes_all = t.empty((num_anchors, 4), dtype = f32, device = device)
    self.size_ok = t.empty((num_anchors,), dtype = u8, device = device)
    self.order = t.empty((pre_nms,), dtype = i32, device = device)
    self.count1 = t.empty((1 + num_anchors,), dtype = i32, device = device)
    self.boxes_sorted = t.empty((pre_nms, 4), dtype = f32, device = device)
    self.scores_sorted = t.empty((pre_nms,), dtype = f32, device = device)
    self.count2 = t.empty((1,), dtype = i32, device = device)
    self.keep = t.empty((pre_nms,), dtype = i32, device = device)
    self.count3 = t.empty((1,), dtype = i32, device = device)


_proposal_buffers = {}


def rpn_proposals(score_map, delta_map, image_shape, feature_pixels, pre_nms, post_nms, anchors = None, keep_mask = None, min_size = 16.0, iou_threshold = 0.7, return_debug = False,
                  defer_count = False, extra_rows = 0):
  """
  models/rpn.py:99-156 as five stream-ordered kernels with no host round trip until the final
  count: decode(+anchors, clip, size flag) -> rank/top-N -> ordered compaction -> NMS bit tiles +
  scan -> gather.  score_map (1,fh,fw,9), delta_map (1,fh,fw,36) contiguous fp32 (NHWC maps).
  Returns proposals (N,4) fp32 (y1,x1,y2,x2).  defer_count = True (training): no host sync at all -- returns the
  capacity-padded (post_nms + extra_rows, 4) buffer (rows past the count are zero) and the device-side count tensor, so the
  caller can append the GT boxes, label, and fetch count + labels in ONE device-to-host copy.
  """
  _require_cuda(score_map, delta_map)
  fh, fw = int(score_map.shape[1]), int(score_map.shape[2])
  a = fh * fw * 9
  dev = score_map.device
  scores = score_map.detach().contiguous()
  deltas = delta_map.detach().contiguous()
  key = (dev.index, a, pre_nms)
  buf = _proposal_buffers.get(key)
  if buf is None:
    buf = ProposalBuffers(a, pre_nms, dev)
    _proposal_buffers[key] = buf
  st = stream()
  L = lib()
  check(L.frcnn_rpn_decode(ptr(deltas), ptr(anchors), fh, fw, int(feature_pixels), int(image_shape[1]), int(image_shape[2]), float(min_size),
                           ptr(buf.boxes_all), ptr(buf.size_ok), None, None, st), "frcnn_rpn_decode")
  check(L.frcnn_topk_order(ptr(scores), ptr(keep_mask), a, pre_nms, ptr(buf.order), ptr(buf.count1), st), "frcnn_topk_order")
  check(L.frcnn_gather_filtered(ptr(buf.boxes_all), ptr(scores), ptr(buf.size_ok), ptr(buf.order), ptr(buf.count1), pre_nms,
                                ptr(buf.boxes_sorted), ptr(buf.scores_sorted), ptr(buf.count2), st), "frcnn_gather_filtered")
  ws, ws_n = workspace(L.frcnn_nms_workspace_bytes(pre_nms), slot = 2)
  check(L.frcnn_nms_sorted_f32(ptr(buf.boxes_sorted), ptr(buf.count2), pre_nms, float(iou_threshold), post_nms, ptr(buf.keep), ptr(buf.count3), ws, ws_n, st), "frcnn_nms_sorted_f32")
  out = (t.zeros if defer_count else t.empty)((post_nms + extra_rows, 4), dtype = t.float32, device = dev)
  check(L.frcnn_gather_rows_f32(ptr(buf.boxes_sorted), 4, ptr(buf.keep), ptr(buf.count3), post_nms, ptr(out), st), "frcnn_gather_rows_f32")
  _lib.count(8)
  if defer_count:
    return out, buf.count3
  n = int(buf.count3.item())                                     # the one host sync of the proposal path
  if return_debug:
    n1 = int(buf.count1[0].item()); n2 = int(buf.count2.item())
    return out[:n], dict(order = buf.order[:n1].clone(), boxes_all = buf.boxes_all.clone(), size_ok = buf.size_ok.clone(),
                         boxes_sorted = buf.boxes_sorted[:n2].clone(), scores_sorted = buf.scores_sorted[:n2].clone(), keep = buf.keep[:n].clone())
  return out[:n]


def rpn_decode(deltas, fh, fw, feature_pixels, img_h, img_w, min_size = 16.0):
  """K5 alone: anchors regenerated in-kernel + delta decode + clip + min-size flag for a (fh*fw*9, 4) delta tensor.
  Returns boxes (A,4) fp32 and size_ok (A) uint8 (rpn_proposals runs the same kernel as the first stage of the proposal path)."""
  _require_cuda(deltas)
  d = deltas.detach().contiguous()
  a = fh * fw * 9
  assert d.numel() == 4 * a
  boxes = t.empty((a, 4), dtype = t.float32, device = d.device)
  ok = t.empty((a,), dtype = t.uint8, device = d.device)
  check(lib().frcnn_rpn_decode(ptr(d), None, fh, fw, int(feature_pixels), int(img_h), int(img_w), float(min_size), ptr(boxes), ptr(ok), None, None, stream()), "frcnn_rpn_decode")
  _lib.count()
  return boxes, ok


def gather_rows(src, index_i32, count_i32, capacity):
  """dst[r] = src[index[r]] for r < *count (rows past the count are left untouched); index / count are device int32."""
  _require_cuda(src, index_i32, count_i32)
  src = src.contiguous()
  row_floats = src.numel() // src.shape[0]
  dst = t.empty((capacity,) + tuple(src.shape[1:]), dtype = t.float32, device = src.device)
  check(lib().frcnn_gather_rows_f32(ptr(src), row_floats, ptr(index_i32), ptr(count_i32), capacity, ptr(dst), stream()), "frcnn_gather_rows_f32")
  _lib.count()
  return dst


def append_rows(dst, dst_count, src):
  """dst[count + r] = src[r] on the device (count stays a device value): faster_rcnn.py:467 without a host round trip."""
  _require_cuda(dst, dst_count, src)
  src = src.contiguous()
  check(lib().frcnn_append_rows_f32(ptr(dst), ptr(dst_count), int(dst.shape[0]), int(dst.shape[1]), ptr(src), int(src.shape[0]), stream()), "frcnn_append_rows_f32")
  _lib.count()


def generate_anchors_device(image_shape, feature_map_hw, feature_pixels, device = "cuda"):
  """Anchor map / valid map as the decode kernel builds them (models/anchors.py:43-135): returns
  CUDA tensors anchors (fh,fw,36) fp32 and valid (fh,fw,9) fp32."""
  fh, fw = int(feature_map_hw[0]), int(feature_map_hw[1])
  a = fh * fw * 9
  zeros = t.zeros((a, 4), dtype = t.float32, device = device)
  boxes = t.empty((a, 4), dtype = t.float32, device = device)
  ok = t.empty((a,), dtype = t.uint8, device = device)
  anchors = t.empty((a, 4), dtype = t.float32, device = device)
  valid = t.empty((a,), dtype = t.float32, device = device)
  check(lib().frcnn_rpn_decode(ptr(zeros), None, fh, fw, int(feature_pixels), int(image_shape[1]), int(image_shape[2]), 16.0,
                               ptr(boxes), ptr(ok), ptr(anchors), ptr(valid), stream()), "frcnn_rpn_decode")
  _lib.count()
  return anchors.reshape(fh, fw, 36), valid.reshape(fh, fw, 9)


def label_proposals(proposals, gt_boxes, gt_classes, num_classes, min_object_iou = 0.5):
  """models/faster_rcnn.py:469-524 for an (n,4) proposal tensor that already contains the appended
  GT boxes.  Returns best_iou (n), class_idx (n) int32, onehot (n,C), packed targets (n,2,4(C-1))."""
  _require_cuda(proposals, gt_boxes, gt_classes)
  p = proposals.detach().contiguous()
  n = p.shape[0]
  m = gt_boxes.shape[0]
  dev = p.device
  best = t.empty((n,), dtype = t.float32, device = dev)
  cls = t.empty((n,), dtype = t.int32, device = dev)
  onehot = t.empty((n, num_classes), dtype = t.float32, device = dev)
  packed = t.empty((n, 2, 4 * (num_classes - 1)), dtype = t.float32, device = dev)
  check(lib().frcnn_label_proposals(ptr(p), n, ptr(gt_boxes.contiguous()), ptr(gt_classes.contiguous()), m, num_classes, float(min_object_iou),
                                    ptr(best), ptr(cls), ptr(onehot), ptr(packed), stream()), "frcnn_label_proposals")
  _lib.count()
  return best, cls, onehot, packed


def detect_postprocess(proposals, classes, deltas, image_hw, score_threshold, iou_threshold = 0.3):
  """models/faster_rcnn.py:179-226 in one launch.  Returns {class_idx: ndarray (k,5) float64}."""
  _require_cuda(proposals, classes, deltas)
  n, c = classes.shape
  dev = classes.device
  result = {}
  if n == 0:
    return {ci: np.zeros((0, 5), dtype = np.float64) for ci in range(1, c)}
  out = t.empty((c - 1, n, 5), dtype = t.float64, device = dev)
  counts = t.empty((c - 1,), dtype = t.int32, device = dev)
  check(lib().frcnn_detect_postprocess(ptr(proposals.detach().contiguous()), ptr(classes.detach().contiguous()), ptr(deltas.detach().contiguous()),
                                       n, c, int(image_hw[0]), int(image_hw[1]), float(np.float32(score_threshold)), float(iou_threshold),
                                       ptr(out), ptr(counts), stream()), "frcnn_detect_postprocess")
  _lib.count()
  out_h = out.cpu().numpy()
  counts_h = counts.cpu().numpy()
  for ci in range(1, c):
    result[ci] = out_h[ci - 1, :counts_h[ci - 1]].copy()
  return result


def sgd_step(param, grad, momentum_buf, lr, momentum, weight_decay, grad_scale = 1.0, first_step = False, carry_split = False):
  """torch.optim.SGD's update (momentum, L2 weight decay) in one kernel, in place.  carry_split: also write the updated weights'
  tf32 operand split into the parameter's persistent buffer (see _weight_splits)."""
  _require_cuda(param, grad, momentum_buf)
  assert param.is_contiguous() or param.is_contiguous(memory_format = t.channels_last)
  assert grad.stride() == param.stride() and momentum_buf.stride() == param.stride()
  e = weight_split_buffer(param) if carry_split else None
  if e is not None and _f16():
    # fp16 engine: the carried split keeps the exponent its buffer holds; a full split (fresh amax) seeds it and refreshes it every 64 steps
    age = e.get("age", 64)
    if age >= 64 or e["version"] != param._version:
      check(lib().frcnn_f16_split(ptr(param), param.numel(), ptr(e["buf"]), stream()), "frcnn_f16_split")
      _lib.count(2)
      age = 0
    e["age"] = age + 1
    check(lib().frcnn_sgd_step_split_f16(ptr(param), ptr(grad), ptr(momentum_buf), param.numel(), float(lr), float(momentum), float(weight_decay), float(grad_scale), int(first_step),
                                         ptr(e["buf"]), stream()), "frcnn_sgd_step_split_f16")
  else:
    check(lib().frcnn_sgd_step_split(ptr(param), ptr(grad), ptr(momentum_buf), param.numel(), float(lr), float(momentum), float(weight_decay), float(grad_scale), int(first_step),
                                     ptr(e["buf"]) if e is not None else None, stream()), "frcnn_sgd_step_split")
  if e is not None:
    e["version"] = param._version                                  # the raw-pointer update does not bump torch's version counter
  _lib.count()


def sgd_step_multi(entries, grad_scale = 1.0, ctas_per_sm = 0):
  """optimizer.step() for a list of (param, grad, momentum_buf, lr, momentum, weight_decay, first_step, carry_split) in one C call
  (frcnn_sgd_step_multi_ex): the per-tensor kernels are launched back to back from C instead of one ctypes round trip each.
  ctas_per_sm: launch shape (0 = default grid-stride launch; large = short-lived CTAs for a side-stream update, see the header)."""
  import ctypes
  n = len(entries)
  if n == 0:
    return
  f16 = _f16()
  L = lib()
  st = stream()
  vp, sz, fl, it = ctypes.c_void_p * n, ctypes.c_size_t * n, ctypes.c_float * n, ctypes.c_int * n
  params, grads, bufs, splits = [], [], [], []
  carried = []
  for param, grad, buf, lr, momentum, wd, first, carry in entries:
    _require_cuda(param, grad, buf)
    assert grad.stride() == param.stride() and buf.stride() == param.stride()
    e = weight_split_buffer(param) if carry else None
    if e is not None and f16:
      age = e.get("age", 64)
      if age >= 64 or e["version"] != param._version:           # (re)seed the carried split's exponent from a fresh amax
        check(L.frcnn_f16_split(ptr(param), param.numel(), ptr(e["buf"]), st), "frcnn_f16_split")
        _lib.count(2)
        age = 0
      e["age"] = age + 1
    params.append(param.data_ptr()); grads.append(grad.data_ptr()); bufs.append(buf.data_ptr())
    splits.append(e["buf"].data_ptr() if e is not None else None)
    carried.append((e, param))
  check(L.frcnn_sgd_step_multi_ex(n, vp(*params), vp(*grads), vp(*bufs), sz(*[x[0].numel() for x in entries]), fl(*[float(x[3]) for x in entries]),
                                  fl(*[float(x[4]) for x in entries]), fl(*[float(x[5]) for x in entries]), it(*[int(x[6]) for x in entries]), vp(*splits),
                                  2 if f16 else 1, float(grad_scale), int(ctas_per_sm), st), "frcnn_sgd_step_multi_ex")
  for e, param in carried:
    if e is not None:
      e["version"] = param._version                                # the raw-pointer update does not bump torch's version counter
  _lib.count(n)
