#!/bin/bash
# Everything that goes into profiles/ for a round: tests, default bench (+ reference arm), launch list, ncu --set full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed|error" > gpurun_out/tests_summary.txt; cat gpurun_out/tests_summary.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-e2e --no-graph --slots 1 --batches 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
KERNELS="knn_kernel knn_heavy_kernel radius_kernel sor_mark_kernel compact_kernel pixel_label_kernel pixel_scatter_kernel select_pass_kernel grid_scatter_kernel" SKIP=1 bash tools/gpu_prof5.sh 2>&1 | grep "rc="
python __graft_entry__.py --smoke 2>&1 | tail -2
python tools/bench_configs.py > gpurun_out/configs.jsonl 2>/dev/null; cat gpurun_out/configs.jsonl
