#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^E  \|^    \|^$" gpurun_out/pytest_gpu.log | tail -25
timeout 300 python tools/tc_trace.py > gpurun_out/trace_persistent.log 2>&1; echo "trace exit $?"
FRCNN_TC_PERSISTENT=0 FRCNN_TC_SPLITS=-1 timeout 300 python tools/tc_trace.py > gpurun_out/trace_oneshot.log 2>&1
grep "^conv" gpurun_out/trace_persistent.log; echo ---; grep "^conv" gpurun_out/trace_oneshot.log
timeout 300 python tools/timeline.py 3 > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"; head -40 gpurun_out/timeline.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
FRCNN_TC_PERSISTENT=0 FRCNN_TC_SPLITS=-1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_oneshot.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_n1_oneshot.json'));print('oneshot', d['value'], d['roofline']['families'])"
FRCNN_LAUNCH_LOG=gpurun_out/launch_log.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 80 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
