#!/bin/bash
# Evidence run for the fp16 engine: bench (ours + reference arm), ncu launch list, ncu --set full of the tcgen05 kernels, timeline, microbench
mkdir -p gpurun_out
timeout 900 python bench.py --steps 40 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
timeout 300 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench exit $?"
timeout 300 python tools/timeline.py 3 > gpurun_out/timeline.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
FRCNN_LAUNCH_LOG=gpurun_out/launch_log.txt timeout 900 ncu --set full --clock-control none -k regex:tc_conv_kernel -c 76 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2>/dev/null
rm -f gpurun_out/prof_tc.ncu-rep
du -sh gpurun_out
