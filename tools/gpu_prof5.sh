#!/bin/bash
# ncu --set full of the named kernels on a 5-frame batch (second call = warm), one report per kernel
mkdir -p gpurun_out
for k in ${KERNELS}; do
  SD_FUSE_SINGLE_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-1} -c 1 -f -o gpurun_out/prof5_$k python tools/profile_once.py ${B:-5} > gpurun_out/ncu5_$k.log 2>&1; echo "$k rc=$?"
done
ls -la gpurun_out/prof5_*.ncu-rep
