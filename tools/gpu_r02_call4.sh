#!/bin/bash
# Round 2, call 4 (one GPU): (a) the GEMM kernel in its two-warpgroup / setmaxnreg layout: parity, then whether a streaming kernel now
# really runs UNDER it (eager SGD A/B); (b) single-CTA vs CTA-pair per layer + one ncu --set full capture of each on the 150x250 layer.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "conv or linear or f16 or tcgen05 or engine or split or heads" > gpurun_out/r02_pytest_wg_kernels.log 2>&1
echo "GEMM kernel tests (warpgroup layout): exit $?"; tail -n 3 gpurun_out/r02_pytest_wg_kernels.log | cut -c1-300
FRCNN_TC_PAIR=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "conv or linear or f16 or tcgen05 or engine or split or heads" > gpurun_out/r02_pytest_wg_pair_kernels.log 2>&1
echo "GEMM kernel tests (warpgroup layout, pair): exit $?"; tail -n 3 gpurun_out/r02_pytest_wg_pair_kernels.log | cut -c1-300
for cfg in "FRCNN_EAGER_SGD=0" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=1" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=4" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=100000"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 1 2> gpurun_out/r02_wg_bench_$tag.err | grep "^{" > gpurun_out/r02_wg_bench_$tag.json
  echo "$cfg: $(python -c "
import json; d=json.load(open('gpurun_out/r02_wg_bench_$tag.json')); f=d['roofline']['families']
print(round(d['value'],1),'images/s',round(d['ms_per_step'],3),'ms |',' '.join('%s %.3f'%(k.replace('conv_','c').replace('linear_','l'),v['ms_per_step']) for k,v in f.items()),'| loss',d['last_loss']['total'])" 2>&1 | tail -n 1)"
  tail -n 2 gpurun_out/r02_wg_bench_$tag.err | cut -c1-300
done
FRCNN_TC_PAIR=0 timeout 200 python tools/pair_probe.py > gpurun_out/r02_pair_probe_single.log 2>&1; echo "probe single: exit $?"; grep -v "^{" gpurun_out/r02_pair_probe_single.log | cut -c1-120
FRCNN_TC_PAIR=1 timeout 200 python tools/pair_probe.py > gpurun_out/r02_pair_probe_pair.log 2>&1; echo "probe pair: exit $?"; grep -v "^{" gpurun_out/r02_pair_probe_pair.log | cut -c1-120
for pv in 0 1; do
  FRCNN_TC_PAIR=$pv PROBE_ITERS=1 PROBE_ONLY=150x250 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 3 -f -o gpurun_out/r02_prof_pair$pv python tools/pair_probe.py > gpurun_out/r02_ncu_pair$pv.log 2>&1
  echo "ncu pair=$pv: exit $?"; tail -n 2 gpurun_out/r02_ncu_pair$pv.log | cut -c1-200
done
