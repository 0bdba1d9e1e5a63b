#!/bin/bash
# Round 2 evidence call (one GPU): whole suite on the default configuration, smoke, the bench lines, ncu launch list + full capture of the
# dominant kernels, compute-sanitizer on the kernel tests.
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "whole GPU suite (defaults): exit $?"; tail -n 4 gpurun_out/r02_pytest_gpu_final.log | cut -c1-300; grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_gpu_final.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke: exit $?"; tail -n 3 gpurun_out/r02_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_bench_n1.err | grep "^{" > gpurun_out/r02_bench_n1.json; echo "bench: exit ${PIPESTATUS[0]}"; cut -c1-330 gpurun_out/r02_bench_n1.json
timeout 300 python bench.py --micro 2> gpurun_out/r02_bench_micro.err | grep "^{" > gpurun_out/r02_bench_micro.json; echo "bench --micro: exit ${PIPESTATUS[0]}"
timeout 400 python bench.py --backbone resnet101 --steps 10 --warmup 3 --min-seconds 1 --no-gpu-eager 2> gpurun_out/r02_bench_resnet101.err | grep "^{" > gpurun_out/r02_bench_resnet101.json; echo "bench resnet101: exit ${PIPESTATUS[0]}"; cut -c1-200 gpurun_out/r02_bench_resnet101.json
timeout 400 python bench.py --backbone resnet50 --batch 2 --roi-op align --rois 300 --steps 10 --warmup 3 --min-seconds 1 --no-gpu-eager 2> gpurun_out/r02_bench_resnet50_b2.err | grep "^{" > gpurun_out/r02_bench_resnet50_b2.json; echo "bench resnet50 b2: exit ${PIPESTATUS[0]}"; cut -c1-200 gpurun_out/r02_bench_resnet50_b2.json
# launch list (cold-cache, serialised): shares, not absolutes.  PDL off under the profiler (kernels are serialised anyway).
FRCNN_PDL=0 FRCNN_LAUNCH_LOG=gpurun_out/r02_launch_log.txt timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launch list: exit $?"; wc -l gpurun_out/r02_launches.csv
FRCNN_PDL=0 FRCNN_LAUNCH_LOG=gpurun_out/r02_launch_log_full.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 76 -c 38 -f -o gpurun_out/r02_prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full (one step of tcgen05 launches): exit $?"; tail -n 2 gpurun_out/r02_ncu_full.log | cut -c1-200
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "nms or roi or decode or rpn or label or loss or softmax or sgd" > gpurun_out/r02_sanitizer_memcheck_small.log 2>&1
echo "memcheck (proposal / RoI / loss kernels): exit $?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_sanitizer_memcheck_small.log | tail -3
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "vgg_64 or f16_engine_exact or tcgen05_linear" > gpurun_out/r02_sanitizer_memcheck_tc.log 2>&1
echo "memcheck (tcgen05 GEMM tests): exit $?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_sanitizer_memcheck_tc.log | tail -3
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "nms_bit_exact or roi_pool or rpn_proposal_stage" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck (NMS / RoIPool / proposal stage): exit $?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r02_sanitizer_racecheck.log | tail -3
