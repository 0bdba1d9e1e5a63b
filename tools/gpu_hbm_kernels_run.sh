#!/bin/bash
# HBM-kernel round: focused parity tests of the changed kernels, microbench, ncu --set full of the RoI / NMS / decode / SGD kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "nms or roi or rpn_proposal or anchor" > gpurun_out/pytest_hbm.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_hbm.log
grep -v "^    \|^$" gpurun_out/pytest_hbm.log | tail -8 | cut -c1-600
timeout 300 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench exit $?"; tail -3 gpurun_out/microbench.err
cut -c1-260 gpurun_out/microbench.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'roi_|nms_|rpn_decode|rank_|sgd_' -o gpurun_out/prof_hbm -f python tools/hbm_kernels_once.py > gpurun_out/ncu_hbm.log 2>&1
echo "ncu hbm exit $?"; tail -2 gpurun_out/ncu_hbm.log | cut -c1-600
ncu -i gpurun_out/prof_hbm.ncu-rep --page raw --csv > gpurun_out/prof_hbm_raw.csv 2>/dev/null
du -sh gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['traffic'])"
