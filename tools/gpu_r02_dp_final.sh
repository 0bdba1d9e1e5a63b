#!/bin/bash
# Round 2, final two-GPU call (gpurun --gpus 2): the four two-rank numerics variants (margin records come back in
# gpurun_out/parity_margins_test_dp_gpu.jsonl), then the driver's own N=1 and N=2 bench commands on this box with the defaults.
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins_test_dp_gpu.jsonl
timeout 1500 python -m pytest tests/test_dp_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest_dp2_final.log 2>&1
echo "two-rank numerics (4 variants): exit $?"; tail -n 4 gpurun_out/r02_pytest_dp2_final.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager 2> gpurun_out/r02_final_n1.err | grep "^{" > gpurun_out/r02_final_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r02_final_n2.err | grep "^{" > gpurun_out/r02_final_n2.json
python - <<'PY'
import json
for n in (1, 2):
  try:
    d = json.load(open("gpurun_out/r02_final_n%d.json" % n))
    print("N=%d: %.1f images/s %.3f ms/step e2e %.1f | %s | loss %.6f" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("parallelism", "")[:60], d["last_loss"]["total"]))
  except Exception as e:
    print("N=%d: no result (%s)" % (n, e))
PY
tail -n 3 gpurun_out/r02_final_n2.err | cut -c1-300
