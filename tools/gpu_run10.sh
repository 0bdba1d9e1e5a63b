#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^E  \|^    \|^$" gpurun_out/pytest_gpu.log | tail -25
timeout 300 python tools/tc_trace.py > gpurun_out/trace_streamk.log 2>&1; echo "trace exit $?"
grep "^conv" gpurun_out/trace_streamk.log
timeout 300 python tools/timeline.py 3 > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"; head -12 gpurun_out/timeline.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
FRCNN_TC_STREAMK=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_splitk.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_n1_splitk.json'));print('splitk', d['value'], {k:(round(v['tflops'],1), round(v['ms_per_step'],3)) for k,v in d['roofline']['families'].items()})"
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print('streamk', d['value'], {k:(round(v['tflops'],1), round(v['ms_per_step'],3)) for k,v in d['roofline']['families'].items()})"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "tcgen05 or act_bwd or nms_bit_exact or rgb_stem" > gpurun_out/sanitizer_tc.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer_tc.log; tail -4 gpurun_out/sanitizer_tc.log
