#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/f16_debug.py > gpurun_out/f16_debug.log 2>&1; echo "f16 debug exit $?"; tail -60 gpurun_out/f16_debug.log | cut -c1-300
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "f16" > gpurun_out/pytest_f16.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_f16.log
grep -v "^    \|^$" gpurun_out/pytest_f16.log | tail -30 | cut -c1-300
