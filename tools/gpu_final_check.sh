#!/bin/bash
# Round-end check: the whole GPU suite, smoke(), the bench line (ours + reference arm)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^    \|^$" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py --steps 40 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; cut -c1-250 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
