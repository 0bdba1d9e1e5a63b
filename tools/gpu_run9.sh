#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^E  \|^    \|^$" gpurun_out/pytest_gpu.log | tail -25
timeout 300 python tools/tc_trace.py > gpurun_out/trace_persistent.log 2>&1; echo "trace exit $?"
grep "^conv" gpurun_out/trace_persistent.log
timeout 300 python tools/timeline.py 3 > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"; head -12 gpurun_out/timeline.log | cut -c1-200
timeout 300 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; grep -i "nms\|stem" gpurun_out/microbench.jsonl | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 420 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
FRCNN_LAUNCH_LOG=gpurun_out/launch_log.txt timeout 900 ncu --set full --clock-control none -k regex:tc_conv_kernel -c 76 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2>/dev/null
ls -la gpurun_out/prof_tc.ncu-rep; sz=$(stat -c %s gpurun_out/prof_tc.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/prof_tc.ncu-rep; fi
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 20 -c 3 -o gpurun_out/prof_tc_src -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_src.log 2>&1
du -sh gpurun_out
