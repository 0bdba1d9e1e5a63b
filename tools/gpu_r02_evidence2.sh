#!/bin/bash
# Round 2 evidence call 2 (one GPU): proposal chain on a side stream (parity + A/B), racecheck details, ncu launch list + full capture
# (converted to CSV on the box: the .ncu-rep files are too large to bring back).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_zz_pdl_gpu.py tests/test_zz_experiments_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest_propstream.log 2>&1
echo "model tests with the proposal chain on a side stream: exit $?"; tail -n 3 gpurun_out/r02_pytest_propstream.log | cut -c1-300; grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_propstream.log | head
for cfg in "FRCNN_PROPOSAL_STREAM=0" "FRCNN_PROPOSAL_STREAM=1" "FRCNN_PROPOSAL_STREAM=1 FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --min-seconds 2 2> gpurun_out/r02_ps_$tag.err | grep "^{" > gpurun_out/r02_ps_$tag.json
  echo "$cfg: $(python -c "
import json; d=json.load(open('gpurun_out/r02_ps_$tag.json')); print(round(d['value'],1),'images/s',round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],1),'| loss',d['last_loss']['total'])" 2>&1 | tail -n 1)"
  tail -n 2 gpurun_out/r02_ps_$tag.err | cut -c1-300
done
timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "nms_bit_exact or roi_pool or rpn_proposal_stage" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck: exit $?"; grep -A6 "Error: Race\|Warning: Race\|hazard" gpurun_out/r02_sanitizer_racecheck.log | cut -c1-260 | head -70
for tool in memcheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "nms or roi or decode or rpn or label or loss or softmax or sgd or vgg_64 or f16_engine_exact or tcgen05_linear" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool: exit $?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_sanitizer_$tool.log | tail -2
done
FRCNN_PDL=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launch list: exit $?"; wc -l gpurun_out/r02_launches.csv
FRCNN_PDL=0 FRCNN_LAUNCH_LOG=gpurun_out/r02_launch_log.txt timeout 900 ncu --set full --clock-control none -k regex:tc_conv_kernel -c 76 -f -o /tmp/r02_prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full: exit $?"; ncu -i /tmp/r02_prof_tc.ncu-rep --page raw --csv > gpurun_out/r02_prof_tc_raw.csv 2>/dev/null; ls -la gpurun_out/r02_prof_tc_raw.csv /tmp/r02_prof_tc.ncu-rep | cut -c1-120
du -sh gpurun_out
