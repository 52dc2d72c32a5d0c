#!/bin/bash
# quick iteration: selected tests, bench, and launch lists for a few tunables
mkdir -p gpurun_out
for f in ${TESTS:-tests/test_gpu_knn.py tests/test_gpu_fuse.py}; do
  timeout 900 python -m pytest $f -q -m gpu -x --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/$(basename $f .py).log
done
python bench.py --steps ${STEPS:-100} --warmup 10 --skip-cpu-baseline --skip-e2e ${BENCH_ARGS} > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_dev.json'));print('value',d['value'],'ms/step',d['ms_per_step'],'pixel_ms',d['roofline']['kernel_ms'],'mismatch',d['result_mismatches_vs_first_pass'],'lat',d['batch_latency_ms'])"
tail -n 5 gpurun_out/bench_dev.err
for cs in ${CELL_SCALES}; do
  SD_KNN_CELL_SCALE=$cs ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cs$cs.csv \
     python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-e2e --no-graph --slots 1 --batches 1 > gpurun_out/ncu_list_cs$cs.log 2>&1
  echo "== cell scale $cs"; python tools/ncu_summary.py gpurun_out/launches_cs$cs.csv | head -8
done
