#!/bin/bash
# first GPU round: parity tests, sanitizer on the small kernels, timing probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -k "anchor or roi_pool_fwd or label or degenerate or detect_post or losses" -p no:cacheprovider > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer.log
tail -15 gpurun_out/sanitizer.log
timeout 900 python tools/gpu_probe.py simt > gpurun_out/probe_simt.log 2>&1
echo "probe exit $?" >> gpurun_out/probe_simt.log
tail -70 gpurun_out/probe_simt.log
