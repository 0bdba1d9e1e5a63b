#!/bin/bash
# Round 2 (gpurun --gpus 2): the two-rank numerics on a ResNet backbone (config 4's data-parallel path), then ResNet-101 at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -q --timeout 600 -p no:cacheprovider -k resnet50 > gpurun_out/r02_pytest_dp2_resnet.log 2>&1
echo "two-rank ResNet-50 numerics: exit $?"; tail -n 4 gpurun_out/r02_pytest_dp2_resnet.log | cut -c1-300
grep "\[margins\]" gpurun_out/r02_pytest_dp2_resnet.log | cut -c1-500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --backbone resnet101 --steps 20 --warmup 5 --min-seconds 1 --no-cpu-baseline 2> gpurun_out/r02_resnet101_n2.err | grep "^{" > gpurun_out/r02_resnet101_n2.json
python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/r02_resnet101_n2.json"))
  print("ResNet-101 N=2: %.1f images/s %.3f ms/step e2e %.1f | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("parallelism", "")[:80]))
except Exception as e:
  print("no result:", e)
PY
tail -n 3 gpurun_out/r02_resnet101_n2.err | cut -c1-300
