#!/bin/bash
# banner tests + frame processor + ABI smoke on the GPU box
mkdir -p gpurun_out
for f in tests/test_gpu_banner.py tests/test_gpu_frame_processor.py tests/test_gpu_overlay.py; do
  timeout 600 python -m pytest $f -q -m gpu -x --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1; rc=$?
  echo "$f rc=$rc $(tail -n 1 gpurun_out/$(basename $f .py).log)"
  if [ $rc -ne 0 ]; then tail -n 40 gpurun_out/$(basename $f .py).log; fi
done
