#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "nms or rpn or f16_engine" > gpurun_out/pytest_kernels.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_kernels.log
grep -v "^    \|^$" gpurun_out/pytest_kernels.log | tail -25 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
echo "bench exit $?"; tail -3 gpurun_out/bench_f16.err
python -c "import json;d=json.load(open('gpurun_out/bench_f16.json'));print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['dtype'], {k:(round(v['tflops'],1), round(v['ms_per_step'],3)) for k,v in d['roofline']['families'].items()}, d['last_loss'])"
timeout 300 python tools/microbench.py 2>/dev/null | grep -i "nms" | cut -c1-250
timeout 300 python tools/cpu_profile.py 20 > gpurun_out/cpu_profile.log 2>&1; head -50 gpurun_out/cpu_profile.log | cut -c1-160
