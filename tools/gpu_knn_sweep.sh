#!/bin/bash
# k-NN tuning sweep (developer experiment): library variants under semantic_depth_b200/variants x cell scales;
# prints frames/s and the isolated k-NN segment time of each; the k-NN parity tests run once per variant
mkdir -p gpurun_out
cp semantic_depth_b200/libsd_fusion.so /tmp/libsd_orig.so
for spec in ${SWEEP}; do
  name="${spec%%:*}"; scales="${spec#*:}"
  if [ "$name" != "base" ]; then cp semantic_depth_b200/variants/libsd_fusion_$name.so semantic_depth_b200/libsd_fusion.so; else cp /tmp/libsd_orig.so semantic_depth_b200/libsd_fusion.so; fi
  for cs in ${scales//,/ }; do
    if [ -n "$SWEEP_TESTS" ]; then
      SD_KNN_CELL_SCALE=$cs timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_fullsize.py -q -m gpu -x --timeout 600 > gpurun_out/sweep_test_${name}_$cs.log 2>&1
      echo "tests $name cs=$cs rc=$? $(tail -n 1 gpurun_out/sweep_test_${name}_$cs.log)"
    fi
    SD_KNN_CELL_SCALE=$cs timeout 300 python bench.py --steps ${STEPS:-60} --warmup 6 --batches 3 --skip-cpu-baseline --skip-e2e --skip-configs ${SWEEP_ARGS} > gpurun_out/sweep_${name}_$cs.json 2> gpurun_out/sweep_${name}_$cs.err
    python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/sweep_${name}_$cs.json').read().strip().splitlines()[-1])
    print('== $name cs=$cs: %.1f frames/s  knn alone %.4f ms  ror %.4f ms  plane %.4f fences %.4f mad %.4f  stages sum %.4f ms  mismatches %d golden %s' % (d['value'], d['stage_ms']['road_knn'], d['stage_ms']['road_ror'], d['stage_ms']['road_plane'], d['stage_ms']['fences'], d['stage_ms']['road_mad'], d['stage_ms']['sum'], d['result_mismatches_vs_first_pass'], (d.get('golden_check_batch0') or {}).get('all_stage_counts_equal')))
except Exception as e:
    print('== $name cs=$cs failed', e)
PY
  done
done
cp /tmp/libsd_orig.so semantic_depth_b200/libsd_fusion.so
