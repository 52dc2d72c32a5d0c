#!/bin/bash
mkdir -p gpurun_out
python tools/profile_once.py 1 2>&1 | grep -E "debug|Error|error" 
for k in ${KERNELS}; do
  SD_FUSE_SINGLE_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python tools/profile_once.py 1 > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
done
