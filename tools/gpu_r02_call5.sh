#!/bin/bash
# Round 2, call 5 (one GPU): warp-uniform MMA issue loop (uniform-register descriptors) for both kernels, relaxed remote arrive for the pair.
mkdir -p gpurun_out
for pv in 0 1; do
  FRCNN_TC_PAIR=$pv timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "conv or linear or f16 or tcgen05 or engine or split or heads" > gpurun_out/r02_pytest_c5_pair$pv.log 2>&1
  echo "GEMM kernel tests pair=$pv: exit $?"; tail -n 2 gpurun_out/r02_pytest_c5_pair$pv.log | cut -c1-300
  FRCNN_TC_PAIR=$pv timeout 200 python tools/pair_probe.py > gpurun_out/r02_c5_probe_pair$pv.log 2>&1; echo "probe pair=$pv: exit $?"; grep -v "^{" gpurun_out/r02_c5_probe_pair$pv.log | cut -c1-120
done
for cfg in "FRCNN_TC_PAIR=0" "FRCNN_TC_PAIR=1" "FRCNN_TC_PAIR=0 FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2" "FRCNN_TC_PAIR=0 FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=3"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 1 2> gpurun_out/r02_c5_bench_$tag.err | grep "^{" > gpurun_out/r02_c5_bench_$tag.json
  echo "$cfg: $(python -c "
import json; d=json.load(open('gpurun_out/r02_c5_bench_$tag.json')); f=d['roofline']['families']
print(round(d['value'],1),'images/s',round(d['ms_per_step'],3),'ms |',' '.join('%s %.3f ms %.0f TF'%(k.replace('conv_','c').replace('linear_','l'),v['ms_per_step'],v['tflops']) for k,v in f.items()),'| loss',d['last_loss']['total'])" 2>&1 | tail -n 1)"
  tail -n 2 gpurun_out/r02_c5_bench_$tag.err | cut -c1-300
done
FRCNN_PDL=1 FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2 TIMELINE_DUMP=1 timeout 300 python tools/timeline.py 3 > gpurun_out/r02_c5_timeline_eager.log 2>&1
echo "timeline eager: exit $?"; grep "steps\|overlap:" gpurun_out/r02_c5_timeline_eager.log | head -12 | cut -c1-330
FRCNN_PDL=1 timeout 300 python tools/timeline.py 3 > gpurun_out/r02_c5_timeline.log 2>&1
echo "timeline default: exit $?"; grep "steps" gpurun_out/r02_c5_timeline.log | head -3; sed -n '/per-kernel totals/,$p' gpurun_out/r02_c5_timeline.log | head -32 | cut -c1-160
