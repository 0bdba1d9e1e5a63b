#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "f16 or conv or linear" > gpurun_out/pytest_kernels.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_kernels.log
grep -v "^    \|^$" gpurun_out/pytest_kernels.log | tail -25 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
echo "bench exit $?"; tail -3 gpurun_out/bench_f16.err
python -c "import json;d=json.load(open('gpurun_out/bench_f16.json'));print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['dtype'], {k:(round(v['tflops'],1), round(v['ms_per_step'],3)) for k,v in d['roofline']['families'].items()}, d['last_loss'])"
timeout 1500 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_model_f16.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_model_f16.log
grep -v "^    \|^$" gpurun_out/pytest_model_f16.log | tail -20 | cut -c1-300
timeout 300 python tools/timeline.py 3 > gpurun_out/timeline_f16.log 2>&1; sed -n 2,4p gpurun_out/timeline_f16.log; grep -A 12 "per-kernel totals" gpurun_out/timeline_f16.log | cut -c1-150
