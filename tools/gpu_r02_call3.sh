#!/bin/bash
# Round 2, third GPU call (one GPU): CTA-pair (cta_group::2) GEMM kernels -- parity first, then A/B against the single-CTA kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "conv or linear or f16 or tcgen05 or engine or split or heads" > gpurun_out/r02_pytest_pair_kernels.log 2>&1
echo "GEMM kernel tests on the pair kernels: exit $?"; tail -n 6 gpurun_out/r02_pytest_pair_kernels.log | cut -c1-300
grep -n "Error\|FAILED\|trap\|illegal" gpurun_out/r02_pytest_pair_kernels.log | head -10 | cut -c1-300
FRCNN_PDL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest_gpu_pair.log 2>&1
echo "whole suite (pair kernels, PDL on): exit $?"; tail -n 5 gpurun_out/r02_pytest_gpu_pair.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_gpu_pair.log | head -20 | cut -c1-300
for cfg in "FRCNN_TC_PAIR=0" "FRCNN_TC_PAIR=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 1 2> gpurun_out/r02_bench_$tag.err | grep "^{" > gpurun_out/r02_bench_$tag.json
  echo "$cfg: $(python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$tag.json')); f=d['roofline']['families']
print(round(d['value'],1),'images/s',round(d['ms_per_step'],3),'ms |',' '.join('%s %.3f ms %.0f TF'%(k.replace('conv_','c').replace('linear_','l'),v['ms_per_step'],v['tflops']) for k,v in f.items()),'| loss',d['last_loss']['total'])" 2>&1 | tail -n 1)"
  tail -n 2 gpurun_out/r02_bench_$tag.err | cut -c1-300
done
