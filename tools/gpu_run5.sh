#!/bin/bash
# tests after the kernel changes, microbench, CPU-thread scan for the reference arm, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^E  \|^    \|^$" gpurun_out/pytest_gpu.log | tail -25
timeout 600 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench exit $?"; cat gpurun_out/microbench.jsonl; tail -3 gpurun_out/microbench.err
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA node\(s\)" > gpurun_out/lscpu.txt; cat gpurun_out/lscpu.txt
for n in 8 16 32 64; do FRCNN_CPU_THREADS=$n timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cpu threads', d['cpu_baseline']['cores'], 'images/s', round(d['value'],4), 'ms/step', round(d['ms_per_step'],1))"; done | tee gpurun_out/cpu_thread_scan.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 300 python tools/gpu_probe.py auto 2>&1 | grep -v "^-\|Self C" | tail -45 > gpurun_out/probe_auto.log; cat gpurun_out/probe_auto.log
