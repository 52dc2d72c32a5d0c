#!/bin/bash
# Developer GPU session: smoke, memcheck on a tiny frame, then every GPU test file in its own process.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/gpu_smoke.py 64 128 > gpurun_out/smoke_small.log 2>&1; echo "smoke_small rc=$?"
timeout 300 python tools/gpu_smoke.py 256 512 > gpurun_out/smoke_mid.log 2>&1; echo "smoke_mid rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_smoke.py 64 128 > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
for f in tests/test_gpu_*.py; do
  timeout 900 python -m pytest $f -q -m gpu -x --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1; echo "$f rc=$?"
done
tail -n 30 gpurun_out/smoke_small.log
