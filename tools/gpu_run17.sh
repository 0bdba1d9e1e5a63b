#!/bin/bash
# batch > 1 extension (config 3) parity tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -k "batch2" > gpurun_out/pytest_batch.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_batch.log
grep -v "^    \|^$" gpurun_out/pytest_batch.log | tail -30 | cut -c1-400
