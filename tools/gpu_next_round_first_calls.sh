#!/bin/bash
# First GPU calls of the next round: validate what was written after round 1's GPU budget had ended.
#   bash tools/gpu_next_round_first_calls.sh 1      # one GPU  (~8 min): PDL on for the whole suite, schedule experiments, bench A/B
#   bash tools/gpu_next_round_first_calls.sh 2      # two GPUs (~3 min, gpurun --gpus 2): SM reserve / PDL next to the NCCL all-reduce
mkdir -p gpurun_out
if [ "${1:-1}" = "1" ]; then
  FRCNN_PDL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_pdl.log 2>&1
  echo "suite with PDL on: exit $?"; tail -n 3 gpurun_out/pytest_gpu_pdl.log | cut -c1-200
  FRCNN_TEST_EXPERIMENTS=1 timeout 600 python -m pytest tests/test_zz_experiments_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_experiments.log 2>&1
  echo "experiments: exit $?"; tail -n 3 gpurun_out/pytest_experiments.log | cut -c1-200
  for cfg in "FRCNN_EAGER_SGD=0" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=1" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=2" "FRCNN_EAGER_SGD=1 FRCNN_EAGER_SGD_CTAS=100000"; do
    tag=$(echo "$cfg" | tr ' =' '__')
    env $cfg timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
    echo "$cfg: $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_$tag.json')); print(round(d['value'],1), 'images/s', round(d['ms_per_step'],3), 'ms', d['last_loss']['total'])" 2>&1 | tail -n 1)"
  done
else
  N=${1}
  run() { env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 8; }
  FRCNN_TEST_EXPERIMENTS=1 timeout 900 python -m pytest tests/test_zz_experiments_gpu.py -q -k fused_dp -p no:cacheprovider > gpurun_out/pytest_fused_dp.log 2>&1
  echo "fused DP step vs NCCL + SGD: exit $?"; tail -n 3 gpurun_out/pytest_fused_dp.log | cut -c1-200
  for cfg in "FRCNN_PDL=0" "FRCNN_DP_FUSED=1" "FRCNN_DP_FUSED=1 FRCNN_DP_FUSED_MULTICAST=0" "FRCNN_DP_FUSED=1 FRCNN_PDL=1" "FRCNN_PDL=1" "FRCNN_DP_SM_RESERVE=16 NCCL_MAX_CTAS=16" "FRCNN_DP_SM_RESERVE=8 NCCL_MAX_CTAS=8" "FRCNN_DP_SM_RESERVE=32 NCCL_MAX_CTAS=32" "FRCNN_DP_SM_RESERVE=16 NCCL_MAX_CTAS=16 FRCNN_PDL=1"; do
    tag=$(echo "$cfg" | tr ' =' '__')
    run $cfg > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
    echo "N=$N $cfg: $(python -c "import json; d=json.load(open('gpurun_out/bench_n${N}_$tag.json')); print(round(d['value'],1), 'images/s', round(d['ms_per_step'],3), 'ms')" 2>&1 | tail -n 1)"
  done
fi
