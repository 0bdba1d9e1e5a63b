#!/bin/bash
# Round 2, final one-GPU call: the whole -m gpu suite, smoke(), the driver's bench command, the ncu launch list of the same command.
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins_test_model_gpu.jsonl gpurun_out/parity_margins_test_kernels_gpu.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r02_final_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?"; tail -n 3 gpurun_out/r02_final_pytest_gpu.log | cut -c1-300; grep -n "^FAILED\|^ERROR" gpurun_out/r02_final_pytest_gpu.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
echo "smoke: exit $?"; tail -n 4 gpurun_out/r02_final_smoke.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02_bench_n1.err | grep "^{" > gpurun_out/r02_bench_n1.json
python - <<'PY'
import json
try:
  d = json.load(open("gpurun_out/r02_bench_n1.json")); f = d["roofline"]["families"]
  print("N=1: %.1f images/s %.3f ms/step e2e %.1f | %s | roofline frac %.3f | cpu %.3f images/s | eager %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
        " ".join("%s %.3f" % (k.replace("conv_", "c").replace("linear_", "l"), v["ms_per_step"]) for k, v in f.items()), d["roofline"]["frac"], d["cpu_baseline"]["value"],
        json.dumps(d.get("gpu_eager_baseline"))[:300]))
except Exception as e:
  print("no result:", e)
PY
tail -n 2 gpurun_out/r02_bench_n1.err | cut -c1-300
FRCNN_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launch list: exit $?"; wc -l gpurun_out/r02_launches.csv
du -sh gpurun_out
