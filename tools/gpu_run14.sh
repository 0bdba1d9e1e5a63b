#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cpu_profile.py 20 > gpurun_out/cpu_profile.log 2>&1; echo "profile exit $?"; head -3 gpurun_out/cpu_profile.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
