#!/bin/bash
# Round 2, call 9 (one GPU): racecheck after the nms_scan fix, the schedule A/B (proposal stream x eager SGD) on one box, whole suite, bench.
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins_*.jsonl
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "nms_bit_exact or nms_fp64 or roi_pool or rpn_proposal_stage" > gpurun_out/r02_sanitizer_racecheck_fixed.log 2>&1
echo "racecheck after the fix: exit $?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r02_sanitizer_racecheck_fixed.log | tail -3
for cfg in "FRCNN_PROPOSAL_STREAM=0 FRCNN_EAGER_SGD=0" "FRCNN_PROPOSAL_STREAM=1 FRCNN_EAGER_SGD=0" "FRCNN_PROPOSAL_STREAM=0 FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2" "FRCNN_PROPOSAL_STREAM=1 FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2" "FRCNN_PROPOSAL_STREAM=0 FRCNN_EAGER_SGD=0" "FRCNN_PROPOSAL_STREAM=1 FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 2 2> gpurun_out/r02_c9_$tag.err | grep "^{" > gpurun_out/r02_c9_$tag.json
  echo "$cfg: $(python -c "
import json; d=json.load(open('gpurun_out/r02_c9_$tag.json')); print(round(d['value'],1),'images/s',round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],1),'(%.3f ms)'%d['e2e']['ms_per_step'],'| loss',d['last_loss']['total'])" 2>&1 | tail -n 1)"
  tail -n 2 gpurun_out/r02_c9_$tag.err | cut -c1-300
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "whole GPU suite (defaults): exit $?"; tail -n 4 gpurun_out/r02_pytest_gpu_final.log | cut -c1-300; grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_gpu_final.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke: exit $?"; tail -n 3 gpurun_out/r02_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_bench_n1.err | grep "^{" > gpurun_out/r02_bench_n1.json; echo "bench: exit ${PIPESTATUS[0]}"; cut -c1-330 gpurun_out/r02_bench_n1.json
timeout 300 python bench.py --micro 2> gpurun_out/r02_bench_micro.err | grep "^{" > gpurun_out/r02_bench_micro.json; echo "bench --micro: exit ${PIPESTATUS[0]}"
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 2> gpurun_out/r02_bench_reference.err | grep "^{" > gpurun_out/r02_bench_reference.json; echo "bench --impl reference: exit ${PIPESTATUS[0]}"; cut -c1-200 gpurun_out/r02_bench_reference.json
du -sh gpurun_out
