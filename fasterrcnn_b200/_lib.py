"""
ctypes binding of libfrcnn_sm100.so (C ABI in include/frcnn_b200.h).

PyTorch is used for device memory and streams only: every wrapper takes torch CUDA tensors,
passes ``data_ptr()`` + the current stream to the C entry point and raises on a non-zero
status.  There is NO fallback: if the shared library is missing or a tensor is not on a CUDA
device the call fails loudly.
"""
import ctypes
import os

import torch as t

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfrcnn_sm100.so")

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
ENGINE_AUTO, ENGINE_SIMT_FP32, ENGINE_TC_3XTF32, ENGINE_TC_3XF16 = 0, 1, 2, 3

_vp, _i, _f, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
_GEOM = [_i] * 9

_SIGNATURES = {
  "frcnn_version": (_i, []),
  "frcnn_last_error_string": (ctypes.c_char_p, []),
  "frcnn_set_pdl": (_i, [_i]),
  "frcnn_set_sm_reserve": (_i, [_i]),
  "frcnn_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_conv2d_fwd_workspace_bytes": (_sz, _GEOM + [_i]),
  "frcnn_conv2d_fwd": (_i, [_vp] * 6 + _GEOM + [_i, _i, _vp, _sz, _vp]),
  "frcnn_conv2d_dgrad_workspace_bytes": (_sz, _GEOM + [_i]),
  "frcnn_conv2d_dgrad": (_i, [_vp] * 4 + _GEOM + [_i, _vp, _sz, _vp]),
  "frcnn_conv2d_wgrad_workspace_bytes": (_sz, _GEOM + [_i]),
  "frcnn_conv2d_wgrad": (_i, [_vp] * 3 + _GEOM + [_i, _vp, _sz, _vp]),
  "frcnn_conv2d_uses_tensor_cores": (_i, [_i] + _GEOM + [_i]),
  "frcnn_debug_tc_trace": (None, [_vp]),
  "frcnn_debug_pair_max_active_clusters": (_i, []),
  "frcnn_tf32_split_bytes": (_sz, [_sz]),
  "frcnn_tf32_split": (_i, [_vp, _sz, _vp, _vp]),
  "frcnn_conv2d_fwd_presplit": (_i, [_vp] * 8 + _GEOM + [_i, _vp, _sz, _vp]),
  "frcnn_conv2d_dgrad_presplit": (_i, [_vp] * 6 + _GEOM + [_vp, _sz, _vp]),
  "frcnn_conv2d_wgrad_presplit": (_i, [_vp] * 5 + _GEOM + [_vp, _sz, _vp]),
  "frcnn_f16_split_bytes": (_sz, [_sz]),
  "frcnn_f16_split": (_i, [_vp, _sz, _vp, _vp]),
  "frcnn_conv2d_amax_slots": (_i, [_i] + _GEOM),
  "frcnn_f16_split_from_amax": (_i, [_vp, _sz, _vp, _i, _vp, _vp]),
  "frcnn_f16_split_carried": (_i, [_vp, _sz, _vp, _i, _vp]),
  "frcnn_conv2d_fwd_f16": (_i, [_vp] * 8 + _GEOM + [_i, _vp, _vp, _sz, _vp]),
  "frcnn_conv2d_dgrad_f16": (_i, [_vp] * 6 + _GEOM + [_vp, _vp, _sz, _vp]),
  "frcnn_conv2d_wgrad_f16": (_i, [_vp] * 5 + _GEOM + [_vp, _sz, _vp]),
  "frcnn_conv2d_bwd_f16": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp] + _GEOM + [_vp, _sz, _vp, _sz, _vp]),
  "frcnn_relu_bwd": (_i, [_vp, _vp, _vp, _sz, _vp]),
  "frcnn_act_bwd_scale": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
  "frcnn_sigmoid_bwd": (_i, [_vp, _vp, _vp, _sz, _vp]),
  "frcnn_bias_grad_workspace_bytes": (_sz, [_sz, _i]),
  "frcnn_bias_grad": (_i, [_vp, _vp, _sz, _i, _vp, _sz, _vp]),
  "frcnn_act_bwd_fused_supported": (_i, [_sz, _i]),
  "frcnn_act_bwd_fused_workspace_bytes": (_sz, [_sz, _i]),
  "frcnn_act_bwd_fused": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _i, _vp, _sz, _vp]),
  "frcnn_act_bwd_fused_f16": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _i, _vp, _i, _vp, _sz, _vp]),
  "frcnn_sgd_step_split_f16": (_i, [_vp, _vp, _vp, _sz, _f, _f, _f, _f, _i, _vp, _vp]),
  "frcnn_maxpool2x2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_maxpool2x2_relu_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_maxpool3x3s2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_subsample2": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_upsample2_zero": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_spatial_mean_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
  "frcnn_scale_rows": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
  "frcnn_spatial_mean_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
  "frcnn_add": (_i, [_vp, _vp, _vp, _sz, _vp]),
  "frcnn_rpn_decode": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
  "frcnn_rpn_targets": (_i, [_vp, _vp, _i, _vp, _i, _d, _d, _vp, _vp, _sz, _vp]),
  "frcnn_topk_order": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
  "frcnn_gather_filtered": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
  "frcnn_nms_workspace_bytes": (_sz, [_i]),
  "frcnn_nms_sorted_f32": (_i, [_vp, _vp, _i, _d, _i, _vp, _vp, _vp, _sz, _vp]),
  "frcnn_iou_matrix_f32": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
  "frcnn_decode_boxes_f32": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
  "frcnn_nms_sorted_f64": (_i, [_vp, _vp, _i, _d, _i, _vp, _vp, _vp, _sz, _vp]),
  "frcnn_nms_batched_workspace_bytes": (_sz, [_i, _i, _i]),
  "frcnn_nms_batched_f32": (_i, [_vp, _vp, _i, _i, _d, _i, _vp, _vp, _vp, _sz, _vp]),
  "frcnn_gather_rows_f32": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp]),
  "frcnn_append_rows_f32": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp]),
  "frcnn_roi_pool_fwd": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
  "frcnn_roi_pool_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
  "frcnn_roi_align_fwd": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _f, _i, _i, _vp, _vp]),
  "frcnn_roi_align_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _vp, _vp]),
  "frcnn_label_proposals": (_i, [_vp, _i, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
  "frcnn_rpn_losses": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
  "frcnn_softmax_rows": (_i, [_vp, _vp, _i, _i, _vp]),
  "frcnn_softmax_rows_bwd": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
  "frcnn_detector_losses": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
  "frcnn_heads_workspace_bytes": (_sz, [_i, _i, _i, _i]),
  "frcnn_heads_fwd": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
  "frcnn_heads_bwd": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
  "frcnn_sgd_step": (_i, [_vp, _vp, _vp, _sz, _f, _f, _f, _f, _i, _vp]),
  "frcnn_sgd_step_split": (_i, [_vp, _vp, _vp, _sz, _f, _f, _f, _f, _i, _vp, _vp]),
  "frcnn_sgd_step_multi": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp]),
  "frcnn_sgd_step_multi_ex": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _i, _vp]),
  "frcnn_dp_sgd_fused": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _sz, _f, _f, _f, _f, _i, _i, _vp]),
  "frcnn_detect_postprocess": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _d, _vp, _vp, _vp]),
}

_lib = None


class FrcnnError(RuntimeError):
  pass


def lib():
  """Loads the shared library (once).  Raises if it has not been built: there is no fallback."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise FrcnnError("libfrcnn_sm100.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C fasterrcnn_b200/csrc`" % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
      fn = getattr(handle, name)     # AttributeError if the symbol is not exported
      fn.restype = restype
      fn.argtypes = argtypes
    _lib = handle
  return _lib


def set_pdl(enabled):
  """Programmatic dependent launch for every kernel of the library (include/frcnn_b200.h: frcnn_set_pdl); returns the previous setting.
  Default = the FRCNN_PDL environment variable (off when unset)."""
  return bool(lib().frcnn_set_pdl(1 if enabled else 0))


def set_sm_reserve(sms):
  """SMs the persistent GEMM launches leave to a concurrent collective (include/frcnn_b200.h: frcnn_set_sm_reserve); returns the previous value."""
  return int(lib().frcnn_set_sm_reserve(int(sms)))


def exported_symbols():
  return sorted(_SIGNATURES.keys())


def check(status, what):
  if status != 0:
    msg = lib().frcnn_last_error_string()
    raise FrcnnError("%s failed with status %d: %s" % (what, status, msg.decode() if msg else "?"))


_raw_stream = getattr(t._C, "_cuda_getCurrentRawStream", None)
_device = {"index": None}


def device_index():
  """Index of this process's CUDA device.  One process drives one GPU, so the index is looked up once: asking torch on
  every launch (lazy-init checks + a driver query) was ~0.3 ms of host time per train step.  reset_device() re-reads it."""
  idx = _device["index"]
  if idx is None:
    idx = _device["index"] = t.cuda.current_device()
  return idx


def reset_device():
  _device["index"] = None
  _workspaces.clear()


def stream():
  """cudaStream_t of torch's current stream on this process's device (the C ABI launches on it)."""
  if _raw_stream is not None:
    return _raw_stream(device_index())                   # one C call instead of building a torch.cuda.Stream object per launch
  return t.cuda.current_stream().cuda_stream


def ptr(x):
  """Device pointer of a CUDA tensor (None -> NULL)."""
  if x is None:
    return None
  if not x.is_cuda:
    raise FrcnnError("expected a CUDA tensor: this package has no CPU path")
  return x.data_ptr()


# ---- stream-ordered scratch space (grown on demand, one per slot) ----------------------------
_workspaces = {}


def workspace(nbytes, slot = 0):
  """Returns (ptr, nbytes) of a cached scratch buffer on this process's device."""
  if nbytes == 0:
    return None, 0
  entry = _workspaces.get(slot)
  if entry is None or entry[1] < nbytes:
    buf = t.empty(int(nbytes * 1.25) + 256, dtype = t.uint8, device = t.device("cuda", device_index()))
    entry = _workspaces[slot] = (buf.data_ptr(), buf.numel(), buf)
  return entry[0], entry[1]


# kernels launched through this module since the last reset (bench.py's gpu_launches evidence)
launch_counter = {"calls": 0}


def count(n = 1):
  launch_counter["calls"] += n
