#!/bin/bash
# Round 2 (gpurun --gpus 8): BASELINE config 4 -- ResNet-101, batch 1 per GPU, 8 x B200 data parallel -- one bench line.
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --backbone resnet101 --steps 20 --warmup 5 --min-seconds 1 --no-cpu-baseline 2> gpurun_out/r02_resnet101_n8.err | grep "^{" > gpurun_out/r02_resnet101_n8.json
python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/r02_resnet101_n8.json")); f = d["roofline"]["families"]
  print("ResNet-101 N=8: %.1f images/s %.3f ms/step e2e %.1f | %s | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
        " ".join("%s %.3f" % (k.replace("conv_", "c"), v["ms_per_step"]) for k, v in f.items()), d["config"].get("parallelism", "")[:90]))
except Exception as e:
  print("no result:", e)
PY
tail -n 3 gpurun_out/r02_resnet101_n8.err | cut -c1-300
