#!/bin/bash
# developer counters of the k-NN kernel for a few cell scales (library must be built with -DSD_KNN_STATS)
mkdir -p gpurun_out
for cs in ${CELL_SCALES:-0.5}; do
  echo "== cell scale $cs"; SD_KNN_CELL_SCALE=$cs python tools/profile_once.py ${B:-1} 2>&1 | tail -2
done
if [ -n "$NCU" ]; then
  SD_FUSE_SINGLE_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 1 -c 1 -f -o gpurun_out/prof5_knn_kernel python tools/profile_once.py 5 > gpurun_out/ncu5_knn_kernel.log 2>&1; echo "ncu rc=$?"
fi
