#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric: images/sec, VGG-16 Faster R-CNN train_step (forward +
backward + SGD) on a synthetic 3x600x1000 image, batch 1 per GPU, N in {1,2,4,8} B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--backbone vgg16|resnet50|resnet101] [--batch B] [--roi-op pool|align] [--rois R]     other BASELINE configs (3, 4)
                  [--micro]                                                                            BASELINE config 5 (HBM kernels)
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

One JSON line on rank 0.  `value` = whole-job images/s with the step's inputs resident in HBM;
`e2e` = the same metric through the public API (FasterRCNNModel.train_step) with HOST (pinned)
image + RPN ground-truth buffers copied in every step and the loss read back; `roofline` = the
dominant kernel (implicit-GEMM convolution) against the measured tensor peak; `cpu_baseline` =
the CPU oracle port of the reference step timed on this box's host cores (bounded sample).
--impl reference times that CPU port as the reference arm (the reference is pure Python and
/root/reference does not travel; oracle/ is its pinned restatement).
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import time

import numpy as np
import torch as t

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
T0 = time.perf_counter()

IMAGE_HW = (600, 1000)
WORKLOAD = "VGG-16 Faster R-CNN train_step (fwd+bwd+SGD), synthetic 3x600x1000 image, batch 1/GPU, 2 GT boxes, 128 RoIs"
METRIC = "images/sec fwd+bwd @ 1000x600, batch=1/GPU"
PDL_DEFAULT_MULTI_GPU = "1"                     # measured next to the overlapped exchange kernels: profiles/r02_dp_sweep_n2.md
BACKBONE_NAMES = {"vgg16": "VGG-16", "resnet50": "ResNet-50", "resnet101": "ResNet-101"}


def workload_name(args):
  """BASELINE.json's headline workload by default (config 2); the other configs through --backbone / --batch / --roi-op / --rois."""
  if args.backbone == "vgg16" and args.batch == 1 and args.roi_op == "pool" and args.rois == 128:
    return WORKLOAD
  return "%s Faster R-CNN train_step (fwd+bwd+SGD), synthetic 3x600x1000 image%s, batch %d/GPU, 2 GT boxes, %d RoIs per image, %s" % (
    BACKBONE_NAMES[args.backbone], "s" if args.batch > 1 else "", args.batch, args.rois, "RoIAlign (sampling_ratio 2)" if args.roi_op == "align" else "RoIPool")


def metric_name(args):
  return METRIC if args.batch == 1 else "images/sec fwd+bwd @ 1000x600, batch=%d/GPU" % args.batch


def cpu_model():
  try:
    with open("/proc/cpuinfo") as f:
      for line in f:
        if line.startswith("model name"):
          return line.split(":", 1)[1].strip()
  except OSError:
    pass
  return "unknown"
# algorithmic GEMM work of one step: conv fwd 366.32 + conv bwd 485.20 + RPN 32.8 + detector fc 92.1 GFLOP
GT = [((100.0, 150.0, 400.0, 600.0), 7), ((50.0, 650.0, 500.0, 850.0), 15)]


# engine -> (config.engine, roofline.note, dtype)
ENGINE_NOTES = {
  "f16": ("tcgen05 3xFP16 implicit GEMM for conv/linear fwd, dgrad and wgrad: operands split per tensor into power-of-two scaled fp16 hi + lo/2048 halves, three kind::f16 "
          "products per MAC accumulated in fp32 (TMEM, drained to registers every 256 k) -> fp32-grade results; exact-fp32 CUDA-core kernels for the RGB stem and the 9/21/36/80-wide heads",
          "algorithmic FLOPs (2*M*N*K) per launch / CUDA-event time; the tcgen05 kernels execute 3 fp16 products per algorithmic MAC, so the algorithmic rate is capped at 1/3 of the "
          "dense fp16/bf16 figure used as denominator", "f16x3->f32"),
  "tf32": ("tcgen05 3xTF32 (fp32-grade) implicit GEMM for conv/linear fwd, dgrad and wgrad; exact-fp32 CUDA-core kernels for the RGB stem and the 9/21/36/80-wide heads",
           "algorithmic FLOPs (2*M*N*K) per launch / CUDA-event time; tcgen05 kernels execute 3 TF32 products per algorithmic MAC (3xTF32, fp32-grade), TF32 dense peak is 1/2 of the bf16 "
           "figure used as denominator", "tf32x3->f32"),
  "simt": ("exact-fp32 CUDA-core implicit GEMM everywhere", "algorithmic FLOPs / CUDA-event time on the CUDA cores; the tensor peak is not the bound", "f32"),
}


class Box:
  def __init__(self, corners, class_index):
    self.corners, self.class_index, self.class_name = np.asarray(corners, dtype = np.float32), class_index, str(class_index)


def measured_peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return dict(hbm_gbs = d["hbm_gbs"], tflops = d.get("bf16_tflops_sustained", d["bf16_tflops"]), source = "MEASURED_PEAKS.json (bf16 sustained)")
  return dict(hbm_gbs = 6650.0, tflops = 1400.0, source = "fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# synthetic weights: Kaiming backbone (first layer /50 for the randn*50 image), reference head init
# ------------------------------------------------------------------------------------------------
def ncu_traffic(family):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel family, from the committed `ncu --set full` summary
  (profiles/ncu_traffic.json, written by tools/summarize_ncu.py); None when there is no capture for it."""
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")
  try:
    with open(path) as fp:
      entry = json.load(fp).get(family)
    return None if entry is None else (entry["dram_bytes_per_launch"], "%s (%d launches)" % (entry["source"], entry["launches"]))
  except (OSError, ValueError, KeyError):
    return None


def init_weights(model, seed):
  g = t.Generator(device = "cpu").manual_seed(seed)
  with t.no_grad():
    for key, p in model.named_parameters():
      if ".bn" in key or ".downsample.1." in key or key.endswith("_feature_extractor.1.weight") or key.endswith("_feature_extractor.1.bias"):
        # frozen BatchNorm affine (ResNet): gamma 1 / beta 0, the last BN of every bottleneck scaled down -- dozens of residual blocks
        # with unit-gain branches overflow a random init
        if key.endswith("bn3.weight"):
          p.fill_(0.25)
        continue
      if key.startswith("_stage1") or "_fc" in key or "_layer4" in key:
        if p.dim() > 1:
          fan_in = int(np.prod(p.shape[1:]))
          w = t.randn(p.shape, generator = g) * (2.0 / fan_in) ** 0.5
          if p.dim() == 4 and p.shape[1] == 3:
            w /= 50.0
          p.copy_(w)
        else:
          p.copy_(t.randn(p.shape, generator = g) * 0.01)
      elif key.endswith("_regressor.weight"):
        p.copy_(t.randn(p.shape, generator = g) * 0.001)
      elif p.dim() > 1:
        p.copy_(t.randn(p.shape, generator = g) * 0.01)
      else:
        p.zero_()


class ClockSampler:
  """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): an NVML polling thread (one sample every
  ~5 ms, so a 20-step region of ~150 ms still yields tens of samples; `nvidia-smi -lms` needs ~100 ms before its first line)."""
  REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

  def __init__(self, gpu_index):
    self.gpu_index = gpu_index
    self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
    self.thread, self.stop_flag, self.error = None, False, None

  def _physical_index(self):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
      ids = [v.strip() for v in vis.split(",") if v.strip()]
      if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
        return int(ids[self.gpu_index])
    return self.gpu_index

  def start(self):
    import threading
    try:
      import pynvml
      pynvml.nvmlInit()
      h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
    except Exception as e:                                         # noqa: BLE001
      self.error = "nvml unavailable: %s" % e
      return

    def poll():
      while not self.stop_flag:
        try:
          self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
          bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
          for name, bit in self.REASONS:
            if bits & bit:
              self.reasons.add(name)
          self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
        except Exception as e:                                     # noqa: BLE001
          self.error = str(e)
          return
        time.sleep(0.005)

    self.thread = threading.Thread(target = poll, daemon = True)
    self.thread.start()

  def stop(self):
    self.stop_flag = True
    if self.thread is not None:
      self.thread.join(timeout = 2)
    if not self.sm:
      return dict(sm_mhz = None, sm_max_mhz = self.max_mhz, reasons = [self.error or "no samples"], samples = 0)
    return dict(sm_mhz = statistics.median(self.sm), sm_min_mhz = min(self.sm), sm_max_mhz = self.max_mhz, reasons = sorted(self.reasons), samples = len(self.sm),
                power_w_max = max(self.power) if self.power else None, how = "NVML polled every ~5 ms during the timed region")


# ------------------------------------------------------------------------------------------------
# CPU oracle port (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def host_threads():
  """Threads the CPU arm may really use: min(logical CPUs, affinity mask, cgroup CPU quota)."""
  n = os.cpu_count() or 1
  try:
    n = min(n, len(os.sched_getaffinity(0)))
  except Exception:
    pass
  try:
    with open("/sys/fs/cgroup/cpu.max") as f:
      quota, period = f.read().split()
    if quota != "max":
      n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
  except Exception:
    pass
  env = os.environ.get("FRCNN_CPU_THREADS")
  return int(env) if env else n


def cpu_port_times(steps, warmup, args = None):
  """Per-step seconds of the CPU oracle port on this workload (every step processes args.batch images)."""
  from oracle import frcnn_oracle as orc
  cores = host_threads()
  t.set_num_threads(cores)
  backbone = args.backbone if args is not None else "vgg16"
  batch = args.batch if args is not None else 1
  rois = args.rois if args is not None else 128
  roi_op = args.roi_op if args is not None else "pool"
  if backbone == "vgg16":
    shapes = orc.vgg16_param_shapes()
  else:
    from oracle import resnet_oracle
    shapes = resnet_oracle.param_shapes(backbone)
  params = orc.synth_params(shapes, seed = 0, heads = "reference")
  for k in params:
    if k.endswith("bn3.weight"):
      params[k] = params[k] * 0.3
  model = orc.OracleModel(params, backbone = backbone, proposal_batch_size = rois)
  smps = [orc.synthetic_sample(IMAGE_HW, seed = b, backbone = backbone) for b in range(batch)]
  random.seed(0); np.random.seed(0); t.manual_seed(0)
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    if batch == 1 and roi_op == "pool":
      smp = smps[0]
      model.train_step(smp["image"], smp["anchor_map"], smp["anchor_valid_map"], smp["gt_rpn_map"], smp["gt_rpn_object_indices"],
                       smp["gt_rpn_background_indices"], smp["gt_corners"], smp["gt_class_idxs"])
    else:
      model.train_step_batch(t.cat([s["image"] for s in smps], dim = 0), smps, roi_op = roi_op)
    if i >= warmup:
      times.append(time.perf_counter() - t0)
  return times, cores


def run_reference(args, rank):
  if rank != 0:
    return
  times, cores = cpu_port_times(args.steps, args.warmup, args)
  total = sum(times)
  value = args.batch * len(times) / total
  sample = "%d train_steps of the CPU oracle port (torch-CPU conv/linear + C NMS/RoIPool restatement), %d threads, %s" % (len(times), cores, cpu_model())
  line = dict(impl = "reference", metric = metric_name(args), value = value, unit = "images/s", n_gpus = args.gpus, steps = args.steps, warmup = args.warmup,
              ms_per_step = 1e3 * total / len(times), higher_is_better = True, scaling = "weak", vs_baseline = None, dtype = "f32", data = "synthetic",
              config = dict(workload = workload_name(args), image = "%dx3x600x1000" % args.batch, backbone = args.backbone, parallelism = "host CPU, %d threads" % cores, cpu = cpu_model()),
              cpu_baseline = dict(value = value, unit = "images/s", cores = cores, kind = "port", sample = sample, cpu = cpu_model()),
              e2e = dict(value = value, unit = "images/s", h2d_bytes_per_step = 0, d2h_bytes_per_step = 0), gpu_launches = 0)
  print(json.dumps(line), flush = True)


def make_train_step(dev, args, rank = 0, world = 1):
  """Builds the benchmark workload on `dev`: model + optimizer + synthetic sample(s).  Returns step(from_host) -> Loss."""
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import anchors as fanchors, optim, resnet

  if args.backbone == "vgg16":
    backbone = f.vgg16.VGG16Backbone(dropout_probability = 0.0)
  else:
    backbone = resnet.ResNetBackbone({"resnet50": resnet.Architecture.ResNet50, "resnet101": resnet.Architecture.ResNet101}[args.backbone])
  model = f.FasterRCNNModel(num_classes = 21, backbone = backbone, allow_edge_proposals = True, proposal_batch_size = args.rois, roi_op = args.roi_op)
  init_weights(model, seed = 0)                                   # identical replicas
  model = model.cuda()
  named = list(model.named_parameters())
  # Default exchange by world size, from the round-2 measurements (profiles/r02_dp_sweep_n2.md, r02_dp_n8.md): at 2 ranks the fused
  # peer-memory kernel under the backward (5.87 ms / step against 6.06 for the bucketed NCCL all-reduce); at 8 ranks the NCCL
  # all-reduce on the gradient arena (6.57 against 6.66-6.70).  FRCNN_DP_FUSED=0 / 1 overrides.
  fused_dp = world > 1 and os.environ.get("FRCNN_DP_FUSED", "1" if world <= 2 else "0") not in ("", "0")
  if fused_dp:
    # reduce-scatter + SGD + all-gather as one kernel per bucket over NVLink / NVSwitch (csrc/dp_sgd.cu), overlapped with the backward.
    # The constructor agrees on success across ranks before it touches the parameters, so every rank takes the same branch here.
    try:
      optimizer = optim.NvlsShardedSGD(optim.optimizer_param_groups(model, 5e-4), lr = 1e-3, momentum = 0.9, named_params = named)
    except Exception as e:                                        # e.g. no symmetric-memory support on this box
      print("bench: fused data-parallel step unavailable (%s); NCCL all-reduce + SGD instead" % str(e)[:300], file = sys.stderr)
      fused_dp = False
  if not fused_dp:
    optimizer = optim.DataParallel(optim.create_optimizer(model, 1e-3, 0.9, 5e-4, fused = True), named_params = named)

  # per-rank synthetic samples (each rank its own images)
  h, w = IMAGE_HW
  boxes = [Box(b, c) for b, c in GT]
  anchor_map, anchor_valid_map = fanchors.generate_anchor_maps((3, h, w), model.backbone.compute_feature_map_shape((3, h, w)), 16)
  rpn_map, obj_idx, bg_idx = fanchors.generate_rpn_map(anchor_map, anchor_valid_map, boxes)
  g = t.Generator(device = "cpu").manual_seed(1000 + rank)
  image_host = (t.randn((args.batch, 3, h, w), generator = g) * 50.0).pin_memory()
  gt_map_host = t.from_numpy(rpn_map).unsqueeze(0).pin_memory()
  image_dev, gt_map_dev = image_host.cuda(), gt_map_host.cuda()
  random.seed(rank); t.manual_seed(rank)

  # end-to-end leg: every step's inputs come from page-locked HOST memory through the package's double-buffered feeder
  # (fasterrcnn_b200.datasets.feeder.DeviceFeeder): the copy of step i + 1 runs on the copy engine while step i computes
  from fasterrcnn_b200.datasets.feeder import DeviceFeeder
  feeder = DeviceFeeder(dev)

  def step(from_host = False):
    if from_host:
      if not feeder.pending:
        feeder.submit(image_host, gt_map_host)                    # first step of a region: nothing was prefetched
      img, gmap = feeder.take()
      feeder.submit(image_host, gt_map_host)                      # the next step's inputs (a loader would pass the next sample here)
    else:
      img, gmap = image_dev, gt_map_dev
    if args.batch == 1:
      return model.train_step(optimizer = optimizer, image_data = img, anchor_map = anchor_map, anchor_valid_map = anchor_valid_map, gt_rpn_map = gmap,
                              gt_rpn_object_indices = [obj_idx], gt_rpn_background_indices = [bg_idx], gt_boxes = [boxes])
    samples = [dict(anchor_map = anchor_map, anchor_valid_map = anchor_valid_map, gt_rpn_map = gmap, gt_rpn_object_indices = obj_idx,
                    gt_rpn_background_indices = bg_idx, gt_boxes = boxes) for _ in range(args.batch)]
    return model.train_step_batch(optimizer, img, samples)

  step.h2d_bytes = image_host.numel() * 4 + gt_map_host.numel() * 4
  step.model = model
  step.optimizer = optimizer
  return step


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------
def run_micro(args):
  """BASELINE config 5: the HBM-bound kernels at a scale where bandwidth, not launch latency, is measured (6000 RoIs, 6000 boxes x 20
  classes, 5.3 M anchors, fc1-sized SGD) -- ONE JSON line whose `micro` list holds a roofline record per kernel."""
  sys.path.insert(0, os.path.join(ROOT, "tools"))
  import microbench
  t.cuda.set_device(0)
  rows = microbench.run()
  peaks = measured_peaks()
  top = max((r for r in rows if "frac" in r), key = lambda r: r["ms"])
  line = dict(metric = "config 5 microbench: achieved HBM GB/s per kernel", value = top["achieved_GBs"], unit = "GB/s", n_gpus = 1, steps = 20, warmup = 3, ms_per_step = top["ms"],
              higher_is_better = True, scaling = "weak", vs_baseline = None, dtype = "f32", data = "synthetic",
              config = dict(workload = "NMS / RoIPool / RoIAlign / decode / SGD microbench: 6000 proposals x 20 classes, fm 512x37x62 and 1024x38x63", l2 = "160 MB buffer written between timed iterations (L2 flush)"),
              roofline = dict(bound = "hbm", kernel = top["kernel"] + " " + top["shape"], achieved = top["achieved_GBs"], peak = peaks["hbm_gbs"], unit = "GB/s", frac = top["achieved_GBs"] / peaks["hbm_gbs"],
                              traffic = None, peak_source = "MEASURED_PEAKS.json" if "MEASURED" in peaks["source"] else peaks["source"]),
              micro = rows)
  print(json.dumps(line), flush = True)


def gpu_eager_baseline():
  """The reference's step through eager PyTorch on this GPU (oracle/eager_gpu.py): cuDNN / cuBLAS / ATen / torchvision CUDA ops, fp32
  (TF32 off -- the CPU path's arithmetic) and with torch's default TF32 convolutions (how the reference would run out of the box)."""
  from oracle import eager_gpu
  out = {}
  for name, tf32, steps in (("fp32", False, 6), ("tf32_convs_torch_default", None, 6)):
    try:
      if tf32 is None:
        times, last = eager_gpu.time_train_steps(IMAGE_HW, steps, 3, tf32 = False, torch_default = True)
      else:
        times, last = eager_gpu.time_train_steps(IMAGE_HW, steps, 3, tf32 = tf32)
      out[name] = dict(value = len(times) / sum(times), unit = "images/s", ms_per_step = 1e3 * sum(times) / len(times), steps = len(times), last_total_loss = last[4])
    except Exception as e:                                        # noqa: BLE001  (torchvision CUDA ops missing, out of memory ...)
      out[name] = dict(unavailable = str(e)[:200])
    t.cuda.empty_cache()
  out["what"] = "reference train_step restated on eager PyTorch CUDA ops (cuDNN/cuBLAS convs and linears, ATen elementwise, torchvision nms/roi_pool, torch.optim.SGD, host-side sampling)"
  return out


def run_ours(args, rank, local_rank, world):
  import torch.distributed as dist
  import fasterrcnn_b200 as f
  from fasterrcnn_b200 import anchors as fanchors, ops, optim, _lib

  t.cuda.set_device(local_rank)
  dev = t.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id = dev)

  # programmatic dependent launch (frcnn_set_pdl): results are bit-identical either way (profiles/r01_pdl_ab.json: same losses over
  # 35 steps, same detections; 5.68 -> 5.43 ms/step on one GPU).  FRCNN_PDL=0 / 1 overrides.
  pdl = os.environ.get("FRCNN_PDL", "1" if world == 1 else PDL_DEFAULT_MULTI_GPU) not in ("", "0")
  _lib.set_pdl(pdl)
  step = make_train_step(dev, args, rank, world)
  verbose = os.environ.get("FRCNN_BENCH_VERBOSE", "0") not in ("", "0")

  def say(what):
    if verbose and rank == 0:
      print("bench[%.1f s]: %s" % (time.perf_counter() - T0, what), file = sys.stderr, flush = True)

  def barrier():
    if world > 1:
      dist.barrier()
    t.cuda.synchronize()

  def timed(n, from_host):
    barrier()
    e0, e1 = t.cuda.Event(enable_timing = True), t.cuda.Event(enable_timing = True)
    e0.record()
    for _ in range(n):
      loss = step(from_host)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
      tt = t.tensor([ms], device = dev)
      dist.all_reduce(tt, op = dist.ReduceOp.MAX)
      ms = float(tt.item())
    return ms, loss

  try:
    for i in range(max(args.warmup, 3)):
      step(False)
      if verbose:
        t.cuda.synchronize()
        say("warm-up step %d done" % i)
    t.cuda.synchronize()
  except Exception as e:                                          # insurance for the single-GPU default: a launch attribute this driver rejects
    if not (pdl and world == 1):
      raise
    print("bench: warm-up with programmatic dependent launch failed (%s); measuring with it off" % str(e)[:200], file = sys.stderr)
    pdl = False
    _lib.set_pdl(False)
    step = make_train_step(dev, args, rank, world)
    for _ in range(max(args.warmup, 3)):
      step(False)
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  # The headline: regions of EXACTLY K steps each (barrier + synchronize on both sides, max over ranks), repeated back to back until
  # --min-seconds of device time have been spent, so that the reported number is a sustained one (clocks and power settle); `value` is
  # the MEDIAN region, every region's ms/step is listed.
  _lib.launch_counter["calls"] = 0
  regions = []
  ms_dev, loss = timed(args.steps, False)
  say("first timed region done: %.3f ms/step" % (ms_dev / args.steps))
  launches = _lib.launch_counter["calls"]
  regions.append(ms_dev)
  want = int(min(60, max(0, (1e3 * args.min_seconds - ms_dev) / max(ms_dev, 1e-3))))
  if world > 1:
    wt = t.tensor([want], device = dev)
    dist.broadcast(wt, 0)
    want = int(wt.item())
  for _ in range(want):
    regions.append(timed(args.steps, False)[0])
  ms_dev = statistics.median(regions)
  clocks = sampler.stop() if rank == 0 else None
  rois = step.model.last_step_info.get("num_rois")
  # end-to-end leg: host buffers in, loss out (the loss read-back is part of train_step's return value)
  for _ in range(2):
    step(True)
  say("all %d device-resident regions done" % len(regions))
  e2e_regions = [timed(args.steps, True)[0] for _ in range(max(1, min(5, len(regions))))]
  ms_e2e = statistics.median(e2e_regions)
  say("end-to-end regions done")
  # roofline leg: the same K steps again with a CUDA-event pair around every conv / linear launch (on the launching stream).
  # Kept out of the headline region: two event records per launch cost host time the step is sensitive to.
  ops.kernel_timer.enable(True)
  timed(args.steps, False)
  gemm_stats = ops.kernel_timer.collect()
  ops.kernel_timer.enable(False)
  say("roofline leg done")

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  peaks = measured_peaks()
  engine_name = {_lib.ENGINE_TC_3XF16: "f16", _lib.ENGINE_AUTO: "tf32", _lib.ENGINE_TC_3XTF32: "tf32", _lib.ENGINE_SIMT_FP32: "simt"}[ops.get_engine()]
  images = world * args.batch * args.steps
  value = images / (ms_dev / 1e3)
  e2e_value = images / (ms_e2e / 1e3)
  h2d = step.h2d_bytes
  d2h = 5 * 4 + 4 + 4 * 2100                                     # losses (5 fp32) + proposal count + class indices for the sampler
  # dominant kernel family = implicit-GEMM convolution / linear
  top = max(gemm_stats.items(), key = lambda kv: kv[1]["ms"]) if gemm_stats else (None, None)
  roofline = None
  if top[0] is not None:
    st = top[1]
    achieved = st["gflop"] / st["ms"]                              # GFLOP / ms = TFLOP/s
    traffic, traffic_source = ncu_traffic(top[0]) or (None, None)
    roofline = dict(bound = "tensor", kernel = top[0], achieved = achieved, peak = peaks["tflops"], unit = "TFLOP/s", frac = achieved / peaks["tflops"], traffic = traffic, traffic_unit = "bytes per launch (dram read + write, ncu --set full)", traffic_source = traffic_source,
                    peak_source = peaks["source"], launches = st["launches"], ms_per_step = st["ms"] / args.steps,
                    note = ENGINE_NOTES[engine_name][1],
                    # what the tensor pipe actually executes: 3 MMA products per algorithmic MAC on the tcgen05 engines (1 on the CUDA-core engine)
                    mma_products_per_mac = 1 if engine_name == "simt" else 3,
                    tensor_pipe_frac = (1 if engine_name == "simt" else 3) * achieved / peaks["tflops"],
                    families = {k: dict(tflops = v["gflop"] / v["ms"], ms_per_step = v["ms"] / args.steps, launches = v["launches"]) for k, v in gemm_stats.items()})
  default_workload = args.backbone == "vgg16" and args.batch == 1 and args.roi_op == "pool" and args.rois == 128
  cpu = None
  if world == 1 and not args.no_cpu_baseline:
    n_cpu = 10 if default_workload else 4
    times, cores = cpu_port_times(n_cpu, 2, args)                  # ~10-20 s of CPU work on the box's cores
    cpu = dict(value = args.batch * len(times) / sum(times), unit = "images/s", cores = cores, kind = "port", cpu = cpu_model(),
               sample = "%d timed + 2 warm-up train_steps of the CPU oracle port on the same 600x1000 workload, %d threads" % (n_cpu, cores))
  eager = gpu_eager_baseline() if (world == 1 and default_workload and not args.no_gpu_eager) else None
  fused = isinstance(step.optimizer, optim.NvlsShardedSGD)
  line = dict(metric = metric_name(args), value = value, unit = "images/s", n_gpus = world, steps = args.steps, warmup = max(args.warmup, 3), ms_per_step = ms_dev / args.steps,
              higher_is_better = True, scaling = "weak", vs_baseline = None, dtype = ENGINE_NOTES[engine_name][2], data = "synthetic",
              config = dict(workload = workload_name(args), image = "%dx3x600x1000" % args.batch, backbone = args.backbone, global_batch = world * args.batch, rois_per_image = rois,
                            parallelism = ("dp%%d (one process per GPU; per bucket ONE fused reduce-scatter + SGD + all-gather kernel over NVLink, exchange = %s, overlapped with the backward)" % step.optimizer.exchange if fused
                                           else "dp%d (one process per GPU, bucketed NCCL gradient all-reduce on the gradient arena, overlapped with the backward)" if world > 1
                                           else "dp%d (one process, one GPU: no gradient exchange)") % world,
                            dp_buckets = (len(step.optimizer.arena.buckets) if getattr(step.optimizer, "arena", None) is not None else 0),
                            engine = ENGINE_NOTES[engine_name][0],
                            sm_reserve = step.optimizer.sm_reserve,     # SMs the GEMMs leave to NCCL while reductions are in flight (FRCNN_DP_SM_RESERVE)
                            pdl = "on (programmatic dependent launch between the library's kernels; FRCNN_PDL=0 turns it off)" if pdl else "off",
                            timed_regions = "%d regions of %d steps back to back (%.2f s of device time); value = median region" % (len(regions), args.steps, sum(regions) / 1e3),
                            l2 = "per-step working set (~1.7 GB of weights, activations, gradients) exceeds the 126 MB L2; no explicit flush"),
              regions = dict(count = len(regions), ms_per_step_min = min(regions) / args.steps, ms_per_step_median = ms_dev / args.steps, ms_per_step_max = max(regions) / args.steps,
                             first_region_ms_per_step = regions[0] / args.steps),
              e2e = dict(value = e2e_value, unit = "images/s", h2d_bytes_per_step = h2d, d2h_bytes_per_step = d2h, ms_per_step = ms_e2e / args.steps),
              gpu_launches = launches, clocks = clocks, roofline = roofline, cpu_baseline = cpu, gpu_eager_baseline = eager,
              last_loss = dict(total = loss.total, rpn_class = loss.rpn_class, detector_class = loss.detector_class))
  print(json.dumps(line), flush = True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type = int, default = 1)
  ap.add_argument("--steps", type = int, default = 20)
  ap.add_argument("--warmup", type = int, default = 5)
  ap.add_argument("--impl", default = "ours", choices = ["ours", "reference"])
  ap.add_argument("--no-cpu-baseline", action = "store_true")
  ap.add_argument("--no-gpu-eager", action = "store_true", help = "skip the eager-PyTorch-on-this-GPU comparator leg")
  ap.add_argument("--backbone", default = "vgg16", choices = ["vgg16", "resnet50", "resnet101"])
  ap.add_argument("--batch", type = int, default = 1, help = "images per GPU per step (> 1: train_step_batch, BASELINE config 3)")
  ap.add_argument("--roi-op", default = "pool", choices = ["pool", "align"])
  ap.add_argument("--rois", type = int, default = 128, help = "sampled RoIs per image (proposal_batch_size)")
  ap.add_argument("--min-seconds", type = float, default = 2.0, help = "keep repeating K-step regions until this much device time is spent")
  ap.add_argument("--micro", action = "store_true", help = "BASELINE config 5: HBM-kernel microbench instead of the train step")
  args = ap.parse_args()
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if args.impl == "reference":
    run_reference(args, rank)
    return
  if not t.cuda.is_available():
    raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU path")
  if args.micro:
    if rank == 0:
      run_micro(args)
    return
  if world != args.gpus:
    if args.gpus > 1 and world == 1:
      raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus %d ..." % (args.gpus, args.gpus))
  run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
  main()
