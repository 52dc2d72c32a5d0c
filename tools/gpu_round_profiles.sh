#!/bin/bash
# Everything that goes into profiles/ for a round: tests, default bench (+ reference arm), launch list, ncu --set full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_*.py; do
  timeout 1200 python -m pytest $f -q -m gpu --timeout 1000 > gpurun_out/$(basename $f .py).log 2>&1; echo "$f rc=$?"
  tail -n 1 gpurun_out/$(basename $f .py).log
done 2>&1 | tee gpurun_out/tests_summary.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-e2e --skip-configs --no-kernel-timing --no-graph --slots 1 --batches 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -30 gpurun_out/launches_summary.txt
for k in ${KERNELS:-knn_kernel knn_heavy_kernel radius_kernel sor_mark_kernel compact_kernel pixel_label_kernel pixel_scatter_kernel select_pass_kernel grid_scatter_kernel mean_leaf_kernel}; do
  SD_FUSE_SINGLE_STREAM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof5_$k python tools/profile_once.py 5 > gpurun_out/ncu5_$k.log 2>&1; echo "$k rc=$?"
done
python __graft_entry__.py --smoke 2>&1 | tail -2
