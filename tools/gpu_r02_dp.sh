#!/bin/bash
# Round 2 data-parallel call (gpurun --gpus N): numerics of the N-rank step vs the oracle (N = 2 only), then the schedule sweep.
#   bash tools/gpu_r02_dp.sh 2 [quick]
N=${1:-2}
mkdir -p gpurun_out
run() { env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --min-seconds 1 --no-cpu-baseline; }
summ() { python - "$1" <<'PY'
import json, sys
try:
  d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
  fam = d["roofline"]["families"]
  print(round(d["value"], 1), "images/s", round(d["ms_per_step"], 3), "ms |", " ".join("%s %.3f" % (k.replace("conv_", "c").replace("linear_", "l"), v["ms_per_step"]) for k, v in fam.items()), "| loss", round(d["last_loss"]["total"], 5), "|", d["config"]["parallelism"][:40])
except Exception as e:
  print("no result:", e)
PY
}
if [ "$N" = "2" ]; then
  timeout 1200 python -m pytest tests/test_dp_gpu.py -q -x --timeout 900 -p no:cacheprovider -k "nccl_sharded" > gpurun_out/r02_pytest_dp2.log 2>&1
  echo "two-rank numerics: exit $?"; tail -n 6 gpurun_out/r02_pytest_dp2.log | cut -c1-400
  grep "\[margins\]" gpurun_out/r02_pytest_dp2.log | cut -c1-600
fi
FRCNN_TC_PAIR=1 timeout 300 python bench.py --steps 20 --warmup 5 --min-seconds 1 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_dp_n1.json 2> gpurun_out/r02_dp_n1.err
echo "N=1 on this box: $(summ gpurun_out/r02_dp_n1.json)"
CONFIGS=("FRCNN_DP_FUSED=1" "FRCNN_DP_FUSED=1 FRCNN_DP_EXCHANGE=nccl" "FRCNN_DP_FUSED=1 FRCNN_DP_EXCHANGE=nccl FRCNN_DP_FUSED_OVERLAP=0" "FRCNN_DP_FUSED=0")
if [ "$2" = "quick" ]; then CONFIGS=("FRCNN_DP_FUSED=1 FRCNN_DP_EXCHANGE=nccl" "FRCNN_DP_FUSED=0"); fi
for cfg in "${CONFIGS[@]}"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  run $cfg > gpurun_out/r02_dp_n${N}_$tag.json 2> gpurun_out/r02_dp_n${N}_$tag.err
  echo "N=$N $cfg: $(summ gpurun_out/r02_dp_n${N}_$tag.json)"
  grep -i "error\|unavailable\|Traceback" gpurun_out/r02_dp_n${N}_$tag.err | head -3 | cut -c1-300
done
