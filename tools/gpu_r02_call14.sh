#!/bin/bash
# Round 2, call 14 (one GPU): who calls torch.cuda.is_available() 187 times per ResNet-101 train step?  (Answer: torch._utils.
# _get_available_device_type under autograd.Function.apply, ~5 us each -- the 0.8 s in the profile is the FIRST call, the driver's cuInit.)
mkdir -p gpurun_out
timeout 600 python -m cProfile -o /tmp/resnet101.prof bench.py --backbone resnet101 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager --min-seconds 0 > gpurun_out/r02_c14_bench.log 2>&1
echo "exit $?"
python - <<'PY' > gpurun_out/r02_c14_callers.log 2>&1
import pstats
p = pstats.Stats("/tmp/resnet101.prof")
p.print_callers("is_available")
p.print_callers("_cuda_getDeviceCount")
p.print_callers("is_bf16_supported")
p.print_callers("_lazy_init")
PY
cut -c1-220 gpurun_out/r02_c14_callers.log | head -90
