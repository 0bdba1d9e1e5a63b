#!/bin/bash
# second GPU round: tcgen05 engine bring-up (isolated process), then the full suite, then timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "tcgen05" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
TC_RC=$?
echo "pytest tc exit $TC_RC" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
if [ $TC_RC -ne 0 ]; then export FRCNN_ENGINE=simt; echo "TC engine failed -> running the rest on FRCNN_ENGINE=simt"; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --deselect tests/test_kernels_gpu.py::test_tcgen05_conv_fwd_matches_fp32_engine --deselect tests/test_kernels_gpu.py::test_tcgen05_linear_fwd_splitk_matches_fp32_engine > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python tools/gpu_probe.py ${FRCNN_ENGINE:-auto} > gpurun_out/probe_auto.log 2>&1
echo "probe exit $?" >> gpurun_out/probe_auto.log
tail -75 gpurun_out/probe_auto.log
