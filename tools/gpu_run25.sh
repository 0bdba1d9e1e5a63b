#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_k1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_k1.log
grep -v "^    \|^$" gpurun_out/pytest_k1.log | tail -12 | cut -c1-400
for i in 1 2 3; do timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "rpn_proposal" 2>&1 | tail -2; done
