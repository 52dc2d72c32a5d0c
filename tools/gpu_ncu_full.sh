#!/bin/bash
# ncu --set full captures of the top kernels (second occurrence = warm), one report per kernel.
mkdir -p gpurun_out
for k in ${KERNELS:-knn_kernel radius_kernel pixel_fuse_kernel compact_kernel}; do
  SD_FUSE_SINGLE_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-1} -c 1 -f -o gpurun_out/prof_$k \
     python tools/profile_once.py 1 > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
