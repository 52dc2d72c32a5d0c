#!/bin/bash
# throughput of bench.py for a few pipeline settings (developer experiment)
mkdir -p gpurun_out
for args in "--slots 3" "--slots 3 --no-kernel-timing" "--slots 2" "--slots 4" "--slots 6" "--slots 4 --no-kernel-timing"; do
  python bench.py --steps 100 --warmup 10 --skip-cpu-baseline --skip-e2e $args > gpurun_out/sweep.json 2>gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
  python -c "import json;d=json.load(open('gpurun_out/sweep.json'));print('$args', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'knn_ms', d['roofline']['kernel_ms'], 'pix_ms', d['roofline_pixel']['kernel_ms'])"
done
