#!/bin/bash
# Round 2, call 13 (one GPU): host-side profile of the ResNet-101 train step (cProfile; the step is launch-bound: ~850 C calls per step).
mkdir -p gpurun_out
timeout 600 python -m cProfile -o /tmp/resnet101.prof bench.py --backbone resnet101 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_c13_bench.log 2>&1
echo "exit $?"; tail -n 1 gpurun_out/r02_c13_bench.log | cut -c1-200
python - <<'PY' > gpurun_out/r02_c13_host_profile.log 2>&1
import pstats
p = pstats.Stats("/tmp/resnet101.prof")
p.sort_stats("tottime").print_stats(45)
p.sort_stats("cumulative").print_stats(60)
PY
head -n 70 gpurun_out/r02_c13_host_profile.log | cut -c1-200
