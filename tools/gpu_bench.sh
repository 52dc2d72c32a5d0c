#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps ${STEPS:-100} --warmup 10 --skip-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_dev.json; tail -n 20 gpurun_out/bench_dev.err
if [ -n "$NCU_LIST" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-e2e --no-graph --slots 1 --batches 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
fi
