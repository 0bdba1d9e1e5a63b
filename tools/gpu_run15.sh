#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^    \|^$" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-600
timeout 300 python tools/cpu_profile.py 20 > gpurun_out/cpu_profile.log 2>&1; echo "profile exit $?"; head -2 gpurun_out/cpu_profile.log
timeout 300 python tools/timeline.py 3 > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"; head -16 gpurun_out/timeline.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], {k:(round(v['tflops'],1), round(v['ms_per_step'],3)) for k,v in d['roofline']['families'].items()})"
