#!/bin/bash
# Round 2, call 11 (one GPU): conflict-free RoI staging (parity + config-5 microbench), and WHERE the ResNet-101 step spends its time
# (ncu launch list of bench.py --backbone resnet101; per-kernel shares only -- a number under ncu is never a bench value).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "roi" > gpurun_out/r02_c11_pytest_roi.log 2>&1
echo "RoI kernel tests: exit $?"; tail -n 2 gpurun_out/r02_c11_pytest_roi.log | cut -c1-200
timeout 300 python bench.py --micro 2> gpurun_out/r02_c11_micro.err | grep "^{" > gpurun_out/r02_c11_micro.json
python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/r02_c11_micro.json"))
  for m in d["micro"]:
    print("%-60s %-28s %8.3f ms  %s" % (m["kernel"][:60], m["shape"][:28], m["ms"], ("%.0f GB/s = %.3f" % (m["achieved_GBs"], m["frac"])) if "frac" in m else ""))
except Exception as e:
  print("no micro result:", e)
PY
FRCNN_PDL=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_resnet101_launches.csv \
  python bench.py --backbone resnet101 --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_c11_ncu_resnet.log 2>&1
echo "ncu resnet101 launch list: exit $?"; wc -l gpurun_out/r02_resnet101_launches.csv
python tools/summarize_launches.py gpurun_out/r02_resnet101_launches.csv 2>&1 | head -45 | cut -c1-200
