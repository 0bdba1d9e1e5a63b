#!/bin/bash
mkdir -p gpurun_out
lscpu | grep -i "model name\|flags" | cut -c1-400 > gpurun_out/lscpu_full.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^    \|^$" gpurun_out/pytest_gpu.log | tail -25 | cut -c1-600
for i in 1 2 3; do timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "rpn_proposal_stage" 2>&1 | tail -2; done
timeout 300 python tools/timeline.py 3 > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"; head -6 gpurun_out/timeline.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print('bench', d['value'], d['ms_per_step'], {k:(round(v['tflops'],1), round(v['ms_per_step'],3)) for k,v in d['roofline']['families'].items()})"
