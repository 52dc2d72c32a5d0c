#!/bin/bash
# developer experiment: batches in flight x hardware queue count (CUDA_DEVICE_MAX_CONNECTIONS) x graph / eager submission
mkdir -p gpurun_out
for mc in ${MCS:-8 32}; do
  for sl in ${SLOTS:-3 5 8}; do
    for gr in ${GRAPHS:-graph nograph}; do
      extra=""; [ "$gr" = "nograph" ] && extra="--no-graph"
      CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python bench.py --steps ${STEPS:-120} --warmup 10 --slots $sl --batches ${BATCHES:-6} --skip-cpu-baseline --skip-e2e --skip-configs --no-kernel-timing $extra ${EXTRA_ARGS} > gpurun_out/ovl_${mc}_${sl}_${gr}.json 2> gpurun_out/ovl_${mc}_${sl}_${gr}.err
      python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ovl_${mc}_${sl}_${gr}.json').read().strip().splitlines()[-1])
    print('== maxconn=$mc slots=$sl $gr: %.1f frames/s  %.4f ms/step  latency %.2f ms  mismatches %d' % (d['value'], d['ms_per_step'], d['batch_latency_ms']['mean'], d['result_mismatches_vs_first_pass']))
except Exception as e:
    print('== maxconn=$mc slots=$sl $gr failed', e)
PY
    done
  done
done
