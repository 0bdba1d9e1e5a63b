#!/bin/bash
# full GPU round on the default (auto) engine: tests, probe, bench, ncu launch list + full capture of the TC kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^E  \|^    \|^$" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python tools/gpu_probe.py auto > gpurun_out/probe_auto.log 2>&1
echo "probe exit $?" >> gpurun_out/probe_auto.log
grep -v "^-\|Self C" gpurun_out/probe_auto.log | tail -60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 30 -c 4 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/
