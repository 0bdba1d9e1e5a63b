#!/bin/bash
# Round 2, 8-GPU call (gpurun --gpus 8): the default data-parallel configuration and two alternatives, plus N = 1 on the same box.
N=${1:-8}
mkdir -p gpurun_out
run() { env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --min-seconds 1 --no-cpu-baseline; }
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
timeout 200 python bench.py --steps 20 --warmup 5 --min-seconds 1 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_n8_n1.json 2> gpurun_out/r02_n8_n1.err
echo "N=1 on this box: $(summ gpurun_out/r02_n8_n1.json)"
for cfg in "FRCNN_DP_FUSED=1" "FRCNN_DP_FUSED=1 FRCNN_DP_FUSED_CTAS=3" "FRCNN_DP_FUSED=0" "FRCNN_DP_FUSED=1 FRCNN_DP_FUSED_OVERLAP=0"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  run $cfg > gpurun_out/r02_n${N}_$tag.json 2> gpurun_out/r02_n${N}_$tag.err
  echo "N=$N $cfg: $(summ gpurun_out/r02_n${N}_$tag.json)"
  grep -i "error\|unavailable\|Traceback" gpurun_out/r02_n${N}_$tag.err | head -3 | cut -c1-300
done
