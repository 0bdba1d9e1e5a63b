#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^E  \|^    \|^$" gpurun_out/pytest_gpu.log | tail -15
timeout 300 python tools/tc_trace.py > gpurun_out/trace_2prod.log 2>&1; echo "trace exit $?"
grep "^conv" gpurun_out/trace_2prod.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print('bench', d['value'], d['ms_per_step'], {k:(round(v['tflops'],1), round(v['ms_per_step'],3)) for k,v in d['roofline']['families'].items()})"
