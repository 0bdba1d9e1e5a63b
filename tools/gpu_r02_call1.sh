#!/bin/bash
# Round 2, first GPU call (one GPU): whole suite with PDL on + margins, schedule A/B, bench lines for every single-GPU BASELINE config.
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.jsonl
FRCNN_PDL=1 timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest_gpu_pdl.log 2>&1
echo "suite with PDL on: exit $?"; tail -n 5 gpurun_out/r02_pytest_gpu_pdl.log | cut -c1-300
grep -n "FAILED\|Error" gpurun_out/r02_pytest_gpu_pdl.log | head -20
for cfg in "FRCNN_EAGER_SGD=0" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=1" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=100000"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 1 > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err
  echo "$cfg: $(python -c "import json,sys; d=json.load(open('gpurun_out/r02_bench_$tag.json')); print(round(d['value'],1), 'images/s', round(d['ms_per_step'],3), 'ms', d['regions'], d['last_loss']['total'])" 2>&1 | tail -n 1)"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench default: exit $?"; cut -c1-600 gpurun_out/r02_bench_n1.json; tail -n 3 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --micro > gpurun_out/r02_bench_micro.json 2> gpurun_out/r02_bench_micro.err
echo "bench --micro: exit $?"; cut -c1-300 gpurun_out/r02_bench_micro.json; tail -n 3 gpurun_out/r02_bench_micro.err
timeout 400 python bench.py --backbone resnet101 --steps 10 --warmup 3 --min-seconds 1 --no-gpu-eager > gpurun_out/r02_bench_resnet101.json 2> gpurun_out/r02_bench_resnet101.err
echo "bench resnet101: exit $?"; cut -c1-400 gpurun_out/r02_bench_resnet101.json; tail -n 3 gpurun_out/r02_bench_resnet101.err
timeout 400 python bench.py --backbone resnet50 --batch 2 --roi-op align --rois 300 --steps 10 --warmup 3 --min-seconds 1 --no-gpu-eager > gpurun_out/r02_bench_resnet50_b2.json 2> gpurun_out/r02_bench_resnet50_b2.err
echo "bench resnet50 batch 2 align: exit $?"; cut -c1-400 gpurun_out/r02_bench_resnet50_b2.json; tail -n 3 gpurun_out/r02_bench_resnet50_b2.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
echo "smoke: exit $?"; tail -n 8 gpurun_out/r02_smoke.log | cut -c1-300
