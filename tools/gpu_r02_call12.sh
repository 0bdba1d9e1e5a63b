#!/bin/bash
# Round 2, call 12 (one GPU): ResNet bottlenecks' stride-2 convolutions on the tensor cores + frozen BN as the GEMM epilogue.
# Kernel parity, the ResNet model tests, then bench A/B (FRCNN_RESNET_S2_TC=0 keeps the strided CUDA-core kernels) for configs 4 and 3.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "stride2 or subsample2" > gpurun_out/r02_c12_pytest_s2.log 2>&1
echo "stride-2 kernel tests: exit $?"; tail -n 12 gpurun_out/r02_c12_pytest_s2.log | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -k "resnet" > gpurun_out/r02_c12_pytest_resnet.log 2>&1
echo "ResNet model tests: exit $?"; tail -n 15 gpurun_out/r02_c12_pytest_resnet.log | cut -c1-300
bench() { # tag, env..., -- args
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 1 2> gpurun_out/r02_c12_$tag.err | grep "^{" > gpurun_out/r02_c12_$tag.json
  python - "$tag" <<'PY'
import json, sys
try:
  d = json.load(open("gpurun_out/r02_c12_%s.json" % sys.argv[1])); f = d["roofline"]["families"]
  print("%-22s %.1f images/s %.3f ms/step e2e %.1f launches/step %d | %s | loss %.6f" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"],
        " ".join("%s %.3f" % (k.replace("conv_", "c").replace("linear_", "l"), v["ms_per_step"]) for k, v in f.items()), d["last_loss"]["total"]))
except Exception as e:
  print(sys.argv[1], "no result:", e)
PY
  tail -n 2 gpurun_out/r02_c12_$tag.err | cut -c1-300
}
bench resnet101_s2simt FRCNN_RESNET_S2_TC=0 -- --backbone resnet101
bench resnet101 FRCNN_DUMMY=1 -- --backbone resnet101
bench resnet50_b2 FRCNN_DUMMY=1 -- --backbone resnet50 --batch 2 --roi-op align --rois 300
