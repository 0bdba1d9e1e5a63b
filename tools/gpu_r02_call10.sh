#!/bin/bash
# Round 2, call 10 (one GPU): paired full-sector epilogue stores A/B (probe + bench), parity of the GEMM kernels with them, new defaults.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "conv or linear or f16 or tcgen05 or engine or split or heads" > gpurun_out/r02_c10_pytest_paired.log 2>&1
echo "GEMM kernel tests (paired stores where exposed): exit $?"; tail -n 2 gpurun_out/r02_c10_pytest_paired.log | cut -c1-200
FRCNN_TC_PAIRED_STORES=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "conv or linear or f16 or tcgen05 or engine or split or heads" > gpurun_out/r02_c10_pytest_paired_all.log 2>&1
echo "GEMM kernel tests (paired stores everywhere): exit $?"; tail -n 2 gpurun_out/r02_c10_pytest_paired_all.log | cut -c1-200
for ps in 0 1; do
  FRCNN_TC_PAIRED_STORES=$ps timeout 200 python tools/pair_probe.py > gpurun_out/r02_c10_probe_paired$ps.log 2>&1; echo "probe paired_stores=$ps: exit $?"; grep -v "^{" gpurun_out/r02_c10_probe_paired$ps.log | cut -c1-130
done
for cfg in "FRCNN_TC_PAIRED_STORES=0" "FRCNN_TC_PAIRED_STORES=1" "FRCNN_TC_PAIRED_STORES=-1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  e="$cfg"; if [ "$cfg" = "FRCNN_TC_PAIRED_STORES=-1" ]; then e="FRCNN_DUMMY=1"; fi
  env $e timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 2 2> gpurun_out/r02_c10_$tag.err | grep "^{" > gpurun_out/r02_c10_$tag.json
  echo "$cfg: $(python -c "
import json; d=json.load(open('gpurun_out/r02_c10_$tag.json')); f=d['roofline']['families']
print(round(d['value'],1),'images/s',round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],1),'|',' '.join('%s %.3f'%(k.replace('conv_','c').replace('linear_','l'),v['ms_per_step']) for k,v in f.items()),'| loss',d['last_loss']['total'])" 2>&1 | tail -n 1)"
  tail -n 2 gpurun_out/r02_c10_$tag.err | cut -c1-300
done
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "feeder or config2 or small" > gpurun_out/r02_c10_pytest_model.log 2>&1
echo "model tests (new defaults): exit $?"; tail -n 2 gpurun_out/r02_c10_pytest_model.log | cut -c1-200
