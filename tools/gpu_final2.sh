#!/bin/bash
# last run of the round on the final tree: what the driver runs (one-process GPU suite, smoke, reference arm, default bench)
# + the extended randomised parity tool
bash tools/gpu_final.sh
timeout 600 python tools/gpu_stress_parity.py ${STRESS_CASES:-300} 4242 > gpurun_out/stress_parity.txt 2>&1; echo "stress rc=$?"; tail -n 3 gpurun_out/stress_parity.txt
