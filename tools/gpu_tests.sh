#!/bin/bash
# Every GPU test file in its own process (a CUDA fault in one file must not poison the others).
mkdir -p gpurun_out
for f in tests/test_gpu_*.py; do
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1; echo "$f rc=$?"
  tail -n 3 gpurun_out/$(basename $f .py).log
done
